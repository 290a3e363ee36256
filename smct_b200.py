"""Import alias: ``import smct_b200`` loads the package directory ``smc-t-v2_b200/`` (whose mandated
name is not a Python identifier) and registers it as the package ``smct_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "smc-t-v2_b200")
_spec = importlib.util.spec_from_file_location("smct_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["smct_b200"] = _mod
_spec.loader.exec_module(_mod)
