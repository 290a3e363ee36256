"""ctypes binding of libsmct_b200.so (include/smct.h).  Fails loudly: no library -> no compute."""
import ctypes
import os

from . import build as _build

c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_i32p = ctypes.POINTER(ctypes.c_int32)
VP = ctypes.c_void_p

ABI_VERSION = 3
OK, ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
WEIGHTS = {"gaussian": 0, "classification": 1}
NOISE_Z = {"checkout": 0, "pyc": 1}
INPUT = {"linear": 0, "embedding": 1, "encoded": 2}
ALGO = {"auto": 0, "staged": 1, "fused": 2, "fused_ffma": 3, "staged_seq": 4, "fused_v1": 5, "small": 6}


class SmctShape(ctypes.Structure):
    _fields_ = [("struct_bytes", ctypes.c_int32)] + [(n, ctypes.c_int32) for n in (
        "B", "P", "S", "D", "H", "dff", "F_x", "F_y", "full_model", "attn_window", "len_resampling",
        "weights_mode", "noise_z_mode", "input_mode", "noise_on", "logvar_dim", "algo")] + [
        ("ln_eps", ctypes.c_float), ("sigma_obs", ctypes.c_float), ("b_global0", ctypes.c_int64)]


PARAM_FIELDS = ("Wq", "bq", "Wk", "bk", "Wv", "bv", "Wz", "bz", "ln1_g", "ln1_b", "ln2_g", "ln2_b",
                "W1", "b1", "W2", "b2", "Wo", "W_in", "b_in", "logvar_k", "logvar_q", "logvar_v", "logvar_z")


class SmctParams(ctypes.Structure):
    _fields_ = [(n, VP) for n in PARAM_FIELDS]


GRAD_FIELDS = PARAM_FIELDS


class SmctGrads(ctypes.Structure):
    _fields_ = [(n, VP) for n in GRAD_FIELDS]


LOSS = {"pyc": 0, "checkout": 1}


class SmctIO(ctypes.Structure):
    _fields_ = [("x_in", VP), ("y", VP), ("eps", VP), ("u", VP), ("i_forced", VP), ("seed", ctypes.c_uint64),
                ("pred", VP), ("pred_resampl", VP), ("K", VP), ("V", VP), ("R", VP), ("w", VP), ("i_out", VP),
                ("logw", VP), ("noise_q", VP), ("noise_z", VP), ("noise_K", VP), ("noise_V", VP), ("attn", VP),
                ("loss_sums", VP), ("classic_w", VP)]


class SmctError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsmct_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def _declare(lib):
    i32, i64, u64, f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float
    SP, PP, IP = ctypes.POINTER(SmctShape), ctypes.POINTER(SmctParams), ctypes.POINTER(SmctIO)
    GP = ctypes.POINTER(SmctGrads)
    sig = {
        "smct_abi_version": (ctypes.c_int, []),
        "smct_last_error": (ctypes.c_char_p, []),
        "smct_build_info": (ctypes.c_char_p, []),
        "smct_source_hash": (ctypes.c_char_p, []),
        "smct_launch_count": (ctypes.c_longlong, [ctypes.c_int]),
        "smct_workspace_bytes": (ctypes.c_size_t, [SP]),
        "smct_forward": (ctypes.c_int, [SP, PP, IP, VP, ctypes.c_size_t, VP]),
        "smct_backward_workspace_bytes": (ctypes.c_size_t, [SP]),
        "smct_backward": (ctypes.c_int, [SP, PP, IP, GP, f32, i32, VP, ctypes.c_size_t, VP]),
        "smct_adam_step": (ctypes.c_int, [i64, VP, VP, VP, VP, f32, f32, f32, f32, VP]),
        "smct_cell_step": (ctypes.c_int, [SP, PP, i32, VP, i32, VP, VP, VP, VP, u64, VP, VP, VP, VP, VP, VP, VP, VP,
                                          VP, VP, VP, ctypes.c_size_t, VP]),
        "smct_attention_step": (ctypes.c_int, [SP, PP, i32, VP, VP, u64, VP, VP, VP, VP, VP, VP, VP, VP, VP,
                                               ctypes.c_size_t, VP]),
        "smct_logweights": (ctypes.c_int, [i32, i32, i32, i32, f32, VP, VP, VP, VP]),
        "smct_resample": (ctypes.c_int, [i32, i32, i32, VP, VP, VP, VP, VP, VP]),
        "smct_gather": (ctypes.c_int, [i32, i32, i64, i64, VP, VP, VP, VP]),
        "smct_fill_noise": (ctypes.c_int, [u64, i32, i32, i64, i32, i32, i32, VP, VP]),
        "smct_fill_uniform": (ctypes.c_int, [u64, i32, i64, i32, i32, VP, VP]),
        "smct_sumsq_scaled": (ctypes.c_int, [i64, i32, VP, VP, i32, VP, VP]),
        "smct_tc_selftest": (ctypes.c_int, [VP, VP, VP, VP, VP]),
        "smct_encode": (ctypes.c_int, [SP, PP, VP, VP, VP, VP]),
        "smct_classic_loss": (ctypes.c_int, [i32, i32, i32, i32, i32, VP, VP, VP, VP, VP]),
        "smct_metric": (ctypes.c_int, [i32, i32, i32, i32, i32, VP, VP, VP, VP, VP]),
        "smct_em_update": (ctypes.c_int, [i32, i64, VP, VP, VP, VP, VP, VP, VP, VP, f32, f32, VP, VP]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: loud by design
        fn.restype = res
        fn.argtypes = args
    return sig


def lib_path():
    return _build.LIB


def load(build_if_needed=True):
    """Load libsmct_b200.so; raises (never falls back) when it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_needed:
        _build.build()   # no-op when the embedded source hash matches the tree; a compile error or a stale binary without nvcc raises
    if not os.path.exists(path):
        raise RuntimeError("libsmct_b200.so is not built (%s); run `python smc-t-v2_b200/build.py`. "
                           "There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    lib._smct_signatures = _declare(lib)
    if lib.smct_abi_version() != ABI_VERSION:
        raise RuntimeError("libsmct_b200.so ABI %d != binding ABI %d" % (lib.smct_abi_version(), ABI_VERSION))
    have, want = lib.smct_source_hash().decode(), _build.source_hash()
    if have != want:
        raise RuntimeError("libsmct_b200.so was built from other sources (embedded hash %s, tree %s); run "
                           "`python smc-t-v2_b200/build.py --force`" % (have, want))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SmctError(rc, load().smct_last_error().decode("utf-8", "replace"))
