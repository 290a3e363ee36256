"""CPU-only checks of the drop-in boundary: the library builds/loads, exports every symbol that
include/smct.h declares, and the ctypes structs have the C layout.  No compute call is made."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "smct.h")


@pytest.fixture(scope="module")
def lib():
    import smct_b200
    return smct_b200.load_library()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smct_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for n in ("smct_forward", "smct_cell_step", "smct_attention_step", "smct_logweights", "smct_resample",
              "smct_gather", "smct_workspace_bytes", "smct_last_error", "smct_abi_version"):
        assert n in names


def test_library_exports_every_declared_symbol(lib):
    for n in declared_functions():
        assert hasattr(lib, n), "libsmct_b200.so does not export %s" % n
    assert set(lib._smct_signatures) == set(declared_functions())


def test_abi_version_and_build_info(lib):
    from smct_b200 import _lib
    assert lib.smct_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in lib.smct_build_info()


def test_ctypes_struct_layout_matches_c():
    from smct_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "smct.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(smct_shape), sizeof(smct_params), sizeof(smct_io),
         offsetof(smct_shape, ln_eps), offsetof(smct_shape, b_global0), offsetof(smct_io, seed), offsetof(smct_io, loss_sums));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        vals = [int(v) for v in subprocess.check_output([exe]).split()]
    assert vals == [ctypes.sizeof(_lib.SmctShape), ctypes.sizeof(_lib.SmctParams), ctypes.sizeof(_lib.SmctIO),
                    _lib.SmctShape.ln_eps.offset, _lib.SmctShape.b_global0.offset, _lib.SmctIO.seed.offset,
                    _lib.SmctIO.loss_sums.offset]


def test_shape_validation_without_gpu(lib):
    import smct_b200.functional as F
    s = F.make_shape(2, 4, 3, 8, H=3, dff=8)  # D % H != 0
    assert lib.smct_workspace_bytes(ctypes.byref(s)) == 0
    assert b"divisible" in lib.smct_last_error()
    s = F.make_shape(2, 4, 3, 8, dff=8)
    assert lib.smct_workspace_bytes(ctypes.byref(s)) > 0


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/."""
    pkg = os.path.join(ROOT, "smc-t-v2_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)


def test_no_cpu_fallback():
    import torch
    import smct_b200.functional as F
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        F.fill_uniform(0, 0, 0, 1, 1)


def test_binding_enums_match_the_header():
    """The ctypes binding's name -> value tables must be the header's enums (SMCT_ALGO_*, SMCT_WEIGHTS_*, ...)."""
    from smct_b200 import _lib
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    enums = {k: int(v) for k, v in re.findall(r"\b(SMCT_[A-Z_0-9]+)\s*=\s*(-?\d+)", src)}
    for table, prefix in ((_lib.ALGO, "SMCT_ALGO_"), (_lib.WEIGHTS, "SMCT_WEIGHTS_"), (_lib.NOISE_Z, "SMCT_NOISE_Z_"),
                          (_lib.INPUT, "SMCT_INPUT_"), (_lib.LOSS, "SMCT_LOSS_")):
        declared = {k[len(prefix):].lower(): v for k, v in enums.items() if k.startswith(prefix)}
        assert table == declared, (prefix, table, declared)
    assert (enums["SMCT_OK"], enums["SMCT_ERR_INVALID"], enums["SMCT_ERR_CUDA"], enums["SMCT_ERR_WORKSPACE"],
            enums["SMCT_ERR_UNSUPPORTED"]) == (_lib.OK, _lib.ERR_INVALID, _lib.ERR_CUDA, _lib.ERR_WORKSPACE, _lib.ERR_UNSUPPORTED)


def test_library_matches_the_tree(lib):
    """Build provenance: the .so carries the hash of the sources it was compiled from, and it is this tree's."""
    from smct_b200 import _build_mod as B
    assert lib.smct_source_hash().decode() == B.source_hash() == B.built_hash()
    assert B.source_hash().encode() in lib.smct_build_info()
    assert not B.needs_build()
