"""Thin Python calls over the C-ABI (include/smct.h).  torch is the tensor container only:
device memory, streams and dtype bookkeeping — every computation happens in libsmct_b200.so.
There is no CPU path: a missing library or a non-CUDA tensor raises."""
import ctypes
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import ALGO, INPUT, LOSS, NOISE_Z, WEIGHTS, SmctGrads, SmctIO, SmctParams, SmctShape, check

_ws_cache: Dict[int, torch.Tensor] = {}

OPTIONAL_OUTPUTS = ("i_out", "logw", "noise_q", "noise_z", "noise_K", "noise_V", "attn", "loss_sums")


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("smct_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")


def make_shape(B, P, S, D, H=1, dff=0, F_x=1, F_y=1, full_model=True, attn_window=None, len_resampling=None,
               weights="gaussian", noise_z_mode="checkout", input_mode="linear", noise=True, logvar_dim=1,
               algo="auto", ln_eps=1e-6, sigma_obs=0.5, b_global0=0) -> SmctShape:
    s = SmctShape()
    s.struct_bytes = ctypes.sizeof(SmctShape)
    s.B, s.P, s.S, s.D, s.H, s.dff, s.F_x, s.F_y = int(B), int(P), int(S), int(D), int(H), int(dff), int(F_x), int(F_y)
    s.full_model = int(bool(full_model))
    s.attn_window = -1 if attn_window is None else int(attn_window)
    s.len_resampling = -1 if len_resampling is None else int(len_resampling)
    s.weights_mode = WEIGHTS[weights]
    s.noise_z_mode = NOISE_Z[noise_z_mode]
    s.input_mode = INPUT[input_mode]
    s.noise_on = int(bool(noise))
    s.logvar_dim = int(logvar_dim)
    s.algo = ALGO[algo]
    s.ln_eps = float(ln_eps)
    s.sigma_obs = float(sigma_obs)
    s.b_global0 = int(b_global0)
    return s


def shape_from_config(cfg, **kw) -> SmctShape:
    """Build a shape from an object with the oracle's Config field names (duck-typed; no import of oracle/)."""
    base = dict(B=cfg.B, P=cfg.P, S=cfg.S, D=cfg.D, H=cfg.H, dff=cfg.dff, F_x=cfg.F_x, F_y=cfg.F_y,
                full_model=cfg.full_model, attn_window=cfg.attn_window, len_resampling=cfg.len_resampling,
                weights=cfg.weights, noise_z_mode=cfg.noise_z_mode, input_mode=cfg.input_mode, noise=cfg.noise,
                ln_eps=cfg.ln_eps, sigma_obs=cfg.sigma_obs)
    base.update(kw)
    return make_shape(**base)


def _ptr(t: Optional[torch.Tensor], dtype=None, name="tensor"):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA torch tensor (no CPU fallback)" % name)
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return ctypes.c_void_p(t.data_ptr())


def _check_index_range(t: Optional[torch.Tensor], n: int, name: str):
    """Integer inputs that index device memory (token ids, class ids, forced ancestor indices) must lie in [0, n):
    TensorFlow raises InvalidArgument for them on CPU; the kernels clamp, this raises.  One small reduction + sync,
    only on the integer-input paths (the regression hot path has none)."""
    if t is None or t.numel() == 0:
        return
    lo, hi = (int(v) for v in torch.aminmax(t))
    if lo < 0 or hi >= n:
        raise ValueError("%s has values in [%d, %d], outside [0, %d)" % (name, lo, hi, n))


def _check_inputs(shape: SmctShape, x_in, y, i_forced=None):
    B, S = shape.B, shape.S
    if shape.input_mode == INPUT["embedding"]:
        if tuple(x_in.shape) != (B, S):
            raise ValueError("x_in must have shape (B,S) = %s for token inputs, got %s" % ((B, S), tuple(x_in.shape)))
        _check_index_range(x_in, shape.F_x, "x_in (token ids)")
    else:
        want = (B, S, shape.D if shape.input_mode == INPUT["encoded"] else shape.F_x)
        if tuple(x_in.shape) != want:
            raise ValueError("x_in must have shape %s, got %s" % (want, tuple(x_in.shape)))
    if shape.weights_mode == WEIGHTS["classification"]:
        if tuple(y.shape) != (B, S):
            raise ValueError("y must have shape (B,S) = %s for class targets, got %s" % ((B, S), tuple(y.shape)))
        _check_index_range(y, shape.F_y, "y (class ids)")
    elif tuple(y.shape) != (B, S, shape.F_y):
        raise ValueError("y must have shape %s, got %s" % ((B, S, shape.F_y), tuple(y.shape)))
    if i_forced is not None:
        if tuple(i_forced.shape) != (S, B, shape.P):
            raise ValueError("i_forced must have shape (S,B,P) = %s, got %s" % ((S, B, shape.P), tuple(i_forced.shape)))
        _check_index_range(i_forced, shape.P, "i_forced (ancestor indices)")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def params_struct(params: Dict[str, torch.Tensor], shape: SmctShape) -> SmctParams:
    p = SmctParams()
    for n in _lib.PARAM_FIELDS:
        t = params.get(n)
        if t is None:
            setattr(p, n, None)
            continue
        setattr(p, n, _ptr(t, torch.float32, n))
    for n in ("Wq", "bq", "Wk", "bk", "Wv", "bv", "Wz", "bz", "ln1_g", "ln1_b", "Wo",
              "logvar_k", "logvar_q", "logvar_v", "logvar_z"):
        if params.get(n) is None:
            raise ValueError("parameter %s is required" % n)
    if shape.full_model:
        for n in ("ln2_g", "ln2_b", "W1", "b1", "W2", "b2"):
            if params.get(n) is None:
                raise ValueError("parameter %s is required when full_model" % n)
    return p


def workspace(shape: SmctShape, device=None) -> torch.Tensor:
    lib = _lib.load()
    need = lib.smct_workspace_bytes(ctypes.byref(shape))
    if need == 0:
        raise _lib.SmctError(-1, lib.smct_last_error().decode())
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        _ws_cache.pop(key, None)
        ws = None
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


def release_workspace():
    _ws_cache.clear()


def forward(shape: SmctShape, params: Dict[str, torch.Tensor], x_in: torch.Tensor, y: torch.Tensor,
            eps: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None,
            i_forced: Optional[torch.Tensor] = None, seed: int = 0, want=(), out: Optional[dict] = None) -> dict:
    """SMC_Transformer.call through smct_forward.  Returns a dict with pred, pred_resampl, K, V, R, w and the
    optional outputs named in `want` (subset of OPTIONAL_OUTPUTS).  `out` may hold preallocated tensors."""
    require_cuda()
    lib = _lib.load()
    B, P, S, D, H, Fy = shape.B, shape.P, shape.S, shape.D, shape.H, shape.F_y
    dev = x_in.device
    out = {} if out is None else dict(out)

    def alloc(name, shp, dtype=torch.float32):
        t = out.get(name)
        if t is None:
            t = torch.empty(shp, dtype=dtype, device=dev)
            out[name] = t
        elif tuple(t.shape) != tuple(shp) or t.dtype != dtype:
            raise ValueError("preallocated %s has shape %s/%s, expected %s/%s" % (name, tuple(t.shape), t.dtype, shp, dtype))
        return t

    io = SmctIO()
    x_dtype = torch.int32 if shape.input_mode == INPUT["embedding"] else torch.float32
    y_dtype = torch.int32 if shape.weights_mode == WEIGHTS["classification"] else torch.float32
    io.x_in = _ptr(x_in, x_dtype, "x_in")
    io.y = _ptr(y, y_dtype, "y")
    _check_inputs(shape, x_in, y, i_forced)
    if eps is not None and tuple(eps.shape) != (4, S, B, P, D):
        raise ValueError("eps must have shape (4,S,B,P,D)")
    if eps is not None and shape.D == 64 and shape.algo in (ALGO["auto"], ALGO["fused"], ALGO["fused_v1"]):
        # the fp16-split tensor-core operands of the fused forward are scaled from analytic bounds |LN1(x) W + b| + 16 sigma
        # (smct_tc.cuh); the device generator cannot exceed 6.8 sigma, injected draws are checked here, so the saturating
        # fp16 conversions of the operands never saturate
        if float(eps.abs().max()) > 16.0:
            raise ValueError("injected eps exceeds 16 standard deviations: outside the operand scaling of the fused forward")
    if u is not None and tuple(u.shape) != (S, B, P):
        raise ValueError("u must have shape (S,B,P)")
    io.eps = _ptr(eps, torch.float32, "eps")
    io.u = _ptr(u, torch.float64, "u")
    io.i_forced = _ptr(i_forced, torch.int32, "i_forced")
    io.seed = int(seed)
    io.pred = _ptr(alloc("pred", (B, P, S, Fy)))
    io.pred_resampl = _ptr(alloc("pred_resampl", (B, P, S, Fy)))
    io.K = _ptr(alloc("K", (B, P, S, D)))
    io.V = _ptr(alloc("V", (B, P, S, D)))
    io.R = _ptr(alloc("R", (B, P, S, D)))
    io.w = _ptr(alloc("w", (B, P, S)))
    for name in want:
        if name not in OPTIONAL_OUTPUTS:
            raise ValueError("unknown optional output %r" % name)
    if "i_out" in want:
        io.i_out = _ptr(alloc("i_out", (S, B, P), torch.int32))
    if "logw" in want:
        io.logw = _ptr(alloc("logw", (S, B, P)))
    for n in ("noise_q", "noise_z", "noise_K", "noise_V"):
        if n in want:
            setattr(io, n, _ptr(alloc(n, (B, P, S, D))))
    if "attn" in want:
        io.attn = _ptr(alloc("attn", (B, P, H, S, S)))
    if "loss_sums" in want:
        io.loss_sums = _ptr(alloc("loss_sums", (8,), torch.float64))
    ps = params_struct(params, shape)
    ws = workspace(shape, dev)
    check(lib.smct_forward(ctypes.byref(shape), ctypes.byref(ps), ctypes.byref(io), ctypes.c_void_p(ws.data_ptr()),
                           ws.numel(), _stream()))
    return out


def cell_step(shape: SmctShape, params, t: int, x: torch.Tensor, y, K, V, R, eps4=None, u=None, i_forced=None,
              seed: int = 0, want_attn=True):
    """SMC_Transf_Cell.call through smct_cell_step; K,V,R (B,P,S,D) are updated in place."""
    require_cuda()
    lib = _lib.load()
    B, P, S, D, H, Fy = shape.B, shape.P, shape.S, shape.D, shape.H, shape.F_y
    dev = x.device
    per_particle = 1 if x.dim() == 3 else 0
    if per_particle and tuple(x.shape) != (B, P, D):
        raise ValueError("x must be (B,P,D) or (B,D)")
    if not per_particle and tuple(x.shape) != (B, D):
        raise ValueError("x must be (B,P,D) or (B,D)")
    f32 = torch.float32
    r = torch.empty((B, P, D), dtype=f32, device=dev)
    pred = torch.empty((B, P, Fy), dtype=f32, device=dev)
    attn = torch.empty((B, P, H, S), dtype=f32, device=dev) if want_attn else None
    nq = torch.empty((B, P, D), dtype=f32, device=dev)
    nz = torch.empty((B, P, D), dtype=f32, device=dev)
    w = torch.empty((B, P), dtype=f32, device=dev)
    i_out = torch.empty((B, P), dtype=torch.int32, device=dev)
    y_dtype = torch.int32 if shape.weights_mode == WEIGHTS["classification"] else f32
    ps = params_struct(params, shape)
    ws = workspace(shape, dev)
    check(lib.smct_cell_step(ctypes.byref(shape), ctypes.byref(ps), int(t), _ptr(x, f32, "x"), per_particle,
                             _ptr(y, y_dtype, "y"), _ptr(eps4, f32, "eps4"), _ptr(u, torch.float64, "u"),
                             _ptr(i_forced, torch.int32, "i_forced"), int(seed),
                             _ptr(K, f32, "K"), _ptr(V, f32, "V"), _ptr(R, f32, "R"),
                             _ptr(r), _ptr(pred), _ptr(attn), _ptr(nq), _ptr(nz), _ptr(w), _ptr(i_out),
                             ctypes.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
    return dict(r=r, pred=pred, attn=attn, noise_q=nq, noise_z=nz, w=w, i_t=i_out)


def attention_step(shape: SmctShape, params, t: int, inputs: torch.Tensor, K, V, eps4=None, seed: int = 0):
    """Self_Attention_SMC.call through smct_attention_step; K,V (B,P,S,D) get slot t written in place."""
    require_cuda()
    lib = _lib.load()
    B, P, S, D, H = shape.B, shape.P, shape.S, shape.D, shape.H
    dev = inputs.device
    f32 = torch.float32
    z = torch.empty((B, P, D), dtype=f32, device=dev)
    attn = torch.empty((B, P, H, S), dtype=f32, device=dev)
    noises = [torch.zeros((B, P, D), dtype=f32, device=dev) for _ in range(4)]
    ps = params_struct(params, shape)
    ws = workspace(shape, dev)
    check(lib.smct_attention_step(ctypes.byref(shape), ctypes.byref(ps), int(t), _ptr(inputs, f32, "inputs"),
                                  _ptr(eps4, f32, "eps4"), int(seed), _ptr(K, f32, "K"), _ptr(V, f32, "V"), _ptr(z),
                                  _ptr(attn), _ptr(noises[0]), _ptr(noises[1]), _ptr(noises[2]), _ptr(noises[3]),
                                  ctypes.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
    return dict(z=z, attn=attn, noise_k=noises[0], noise_q=noises[1], noise_v=noises[2], noise_z=noises[3])


def logweights(pred: torch.Tensor, y: torch.Tensor, weights="gaussian", sigma_obs=0.5) -> torch.Tensor:
    """compute_w_regression (log form) / compute_w_classification.  pred (B,P,F), y (B,F) f32 | (B) i32."""
    require_cuda()
    lib = _lib.load()
    B, P, F = pred.shape
    out = torch.empty((B, P), dtype=torch.float32, device=pred.device)
    y_dtype = torch.int32 if weights == "classification" else torch.float32
    check(lib.smct_logweights(B, P, F, WEIGHTS[weights], float(sigma_obs), _ptr(pred, torch.float32, "pred"),
                              _ptr(y, y_dtype, "y"), _ptr(out), _stream()))
    return out


def resample_indices(logw: torch.Tensor, u: torch.Tensor, normalise=True):
    """tf.random.categorical(logits, P) with injected uniforms.  Returns (i_t int32 (B,P), w (B,P), logits (B,P))."""
    require_cuda()
    lib = _lib.load()
    B, P = logw.shape
    i_out = torch.empty((B, P), dtype=torch.int32, device=logw.device)
    w = torch.empty((B, P), dtype=torch.float32, device=logw.device)
    logits = torch.empty((B, P), dtype=torch.float32, device=logw.device)
    check(lib.smct_resample(B, P, int(bool(normalise)), _ptr(logw, torch.float32, "logw"), _ptr(u, torch.float64, "u"),
                            _ptr(w), _ptr(logits), _ptr(i_out), _stream()))
    return i_out, w, logits


def gather(params: torch.Tensor, i_t: torch.Tensor, n_elems: Optional[int] = None) -> torch.Tensor:
    """resample(params, i_t): out[b,m,...] = params[b, i_t[b,m], ...] (first n_elems of each particle row)."""
    require_cuda()
    lib = _lib.load()
    B, P = params.shape[:2]
    row = params[0, 0].numel()
    n = row if n_elems is None else int(n_elems)
    out = torch.empty_like(params) if n == row else params.clone()
    check(lib.smct_gather(B, P, row, n, _ptr(params, torch.float32, "params"), _ptr(i_t, torch.int32, "i_t"), _ptr(out),
                          _stream()))
    return out


def fill_noise(seed, which, t, b_global0, B, P, D, device="cuda") -> torch.Tensor:
    require_cuda()
    lib = _lib.load()
    out = torch.empty((B, P, D), dtype=torch.float32, device=device)
    check(lib.smct_fill_noise(int(seed), int(which), int(t), int(b_global0), B, P, D, _ptr(out), _stream()))
    return out


def fill_uniform(seed, t, b_global0, B, P, device="cuda") -> torch.Tensor:
    require_cuda()
    lib = _lib.load()
    out = torch.empty((B, P), dtype=torch.float64, device=device)
    check(lib.smct_fill_uniform(int(seed), int(t), int(b_global0), B, P, _ptr(out), _stream()))
    return out


def sumsq_scaled(noise: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
    """sum_{rows,d} n^2 * exp(-logvar[d]) accumulated in fp64 (one term of compute_SMC_loss)."""
    require_cuda()
    lib = _lib.load()
    D = noise.shape[-1]
    rows = noise.numel() // D
    out = torch.zeros((1,), dtype=torch.float64, device=noise.device)
    lv = logvar.reshape(-1)
    check(lib.smct_sumsq_scaled(rows, D, _ptr(noise, torch.float32, "noise"), _ptr(lv, torch.float32, "logvar"),
                                int(lv.numel()), _ptr(out), _stream()))
    return out


def classic_loss_sum(pred: torch.Tensor, y: torch.Tensor, weights="gaussian", mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum over (b,p,s) of mask[b,s] * ||y-pred||^2 (gaussian) or of mask[b,s] * sparse cross-entropy (classification),
    fp64.  mask (B,S) f32 (the reference's attention_mask, SMC_Transformer.py:158-163) or None."""
    require_cuda()
    lib = _lib.load()
    B, P, S, Fy = pred.shape
    out = torch.zeros((1,), dtype=torch.float64, device=pred.device)
    y_dtype = torch.int32 if weights == "classification" else torch.float32
    if mask is not None and tuple(mask.shape) != (B, S):
        raise ValueError("mask must have shape (B,S)")
    check(lib.smct_classic_loss(B, P, S, Fy, WEIGHTS[weights], _ptr(pred, torch.float32, "pred"), _ptr(y, y_dtype, "y"),
                                _ptr(mask, torch.float32, "mask"), _ptr(out), _stream()))
    return out


def em_update(noises, logvars, it: float, em_param: float):
    """EM variance update (src/train/utils.py:22-51) on the device: noises = (noise_K, noise_q, noise_V, noise_z), each
    (B,P,S,D); logvars = the four scalar log-variance tensors (1 element each), updated IN PLACE.  No host round trip."""
    require_cuda()
    lib = _lib.load()
    B = noises[0].shape[0]
    per_b = noises[0][0].numel()
    for n in noises:
        if n.shape[0] != B or n[0].numel() != per_b:
            raise ValueError("the four noises must have the same shape")
    for lv in logvars:
        if lv.numel() != 1:
            raise ValueError("EM mode works on scalar log-variances (like the reference)")
    scratch = torch.empty(4 * B, dtype=torch.float64, device=noises[0].device)
    check(lib.smct_em_update(B, per_b, *[_ptr(n, torch.float32, "noise") for n in noises],
                             *[_ptr(lv, torch.float32, "logvar") for lv in logvars], float(it), float(em_param),
                             _ptr(scratch), _stream()))


def metric_sums(pred: torch.Tensor, y: torch.Tensor, weights="gaussian", mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(sum_{b,s} m*metric, sum_{b,s} m) of the training metric on the unresampled predictions (B,P,S,Fy): sparse CE of
    the particle-mean softmax (classification) or squared error of the particle-mean prediction (gaussian); fp64 (2,)."""
    require_cuda()
    lib = _lib.load()
    B, P, S, Fy = pred.shape
    out = torch.zeros((2,), dtype=torch.float64, device=pred.device)
    y_dtype = torch.int32 if weights == "classification" else torch.float32
    if mask is not None and tuple(mask.shape) != (B, S):
        raise ValueError("mask must have shape (B,S)")
    check(lib.smct_metric(B, P, S, Fy, WEIGHTS[weights], _ptr(pred, torch.float32, "pred"), _ptr(y, y_dtype, "y"),
                          _ptr(mask, torch.float32, "mask"), _ptr(out), _stream()))
    return out


def encode(shape: SmctShape, params, x_in: torch.Tensor, layer_norm=False) -> torch.Tensor:
    """get_encoded_input: (B,S,D) encoded rows (optionally after layernorm1), shared by the particles."""
    require_cuda()
    lib = _lib.load()
    X = torch.empty((shape.B, shape.S, shape.D), dtype=torch.float32, device=x_in.device)
    x_dtype = torch.int32 if shape.input_mode == INPUT["embedding"] else torch.float32
    ps = params_struct(params, shape)
    check(lib.smct_encode(ctypes.byref(shape), ctypes.byref(ps), _ptr(x_in, x_dtype, "x_in"),
                          None if layer_norm else _ptr(X), _ptr(X) if layer_norm else None, _stream()))
    return X


def tc_selftest(A: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """(128,64) @ (64,64) on tcgen05 with the three-term bf16 split (building block of the fused dense part)."""
    require_cuda()
    lib = _lib.load()
    D = torch.empty((128, 64), dtype=torch.float32, device=A.device)
    scratch = torch.empty(16384 + 256, dtype=torch.uint8, device=A.device)
    check(lib.smct_tc_selftest(_ptr(A, torch.float32, "A"), _ptr(W, torch.float32, "W"), _ptr(D), _ptr(scratch), _stream()))
    return D


_bws_cache: Dict[int, torch.Tensor] = {}


def backward(shape: SmctShape, params: Dict[str, torch.Tensor], x_in: torch.Tensor, y: torch.Tensor, fwd: dict,
             eps: Optional[torch.Tensor] = None, seed: int = 0, inv_n: Optional[float] = None, loss_variant="pyc",
             grads: Optional[Dict[str, torch.Tensor]] = None, classic_w: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """tape.gradient(loss, trainable_variables) through smct_backward.  `fwd` is the dict returned by forward()
    on the same shape/params/inputs/randomness with want=("i_out","loss_sums").  Gradients are accumulated into
    `grads` (created zero-filled when None).  inv_n defaults to 1/(B*P*S) of this call."""
    require_cuda()
    lib = _lib.load()
    dev = x_in.device
    if grads is None:
        grads = {n: torch.zeros_like(t) for n, t in params.items() if n in _lib.PARAM_FIELDS}
    g = SmctGrads()
    for n in _lib.GRAD_FIELDS:
        setattr(g, n, _ptr(grads.get(n), torch.float32, "grad " + n))
    io = SmctIO()
    x_dtype = torch.int32 if shape.input_mode == INPUT["embedding"] else torch.float32
    io.x_in = _ptr(x_in, x_dtype, "x_in")
    io.y = _ptr(y, torch.int32 if shape.weights_mode == WEIGHTS["classification"] else torch.float32, "y")
    _check_inputs(shape, x_in, y)
    if tuple(fwd["i_out"].shape) != (shape.S, shape.B, shape.P):
        raise ValueError("fwd['i_out'] must have shape (S,B,P)")
    io.eps = _ptr(eps, torch.float32, "eps")
    io.seed = int(seed)
    io.K = _ptr(fwd["K"], torch.float32, "K")
    io.V = _ptr(fwd["V"], torch.float32, "V")
    io.i_out = _ptr(fwd["i_out"], torch.int32, "i_out")
    io.loss_sums = _ptr(fwd.get("loss_sums"), torch.float64, "loss_sums")
    if classic_w is not None and tuple(classic_w.shape) != (shape.B, shape.S):
        raise ValueError("classic_w must have shape (B,S)")
    io.classic_w = _ptr(classic_w, torch.float32, "classic_w")
    for n in ("noise_K", "noise_q", "noise_V", "noise_z"):   # recorded noises: per-dimension log-variance gradients
        setattr(io, n, _ptr(fwd.get(n), torch.float32, n))
    need = lib.smct_backward_workspace_bytes(ctypes.byref(shape))
    if need == 0:
        raise _lib.SmctError(-4, lib.smct_last_error().decode())
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _bws_cache.get(key)
    if ws is None or ws.numel() < need:
        _bws_cache.pop(key, None)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _bws_cache[key] = ws
    if inv_n is None:
        inv_n = 1.0 / (float(shape.B) * shape.P * shape.S)
    ps = params_struct(params, shape)
    check(lib.smct_backward(ctypes.byref(shape), ctypes.byref(ps), ctypes.byref(io), ctypes.byref(g), float(inv_n),
                            LOSS[loss_variant], ctypes.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
    return grads


def adam_step(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr_t: float, beta1=0.9,
              beta2=0.98, epsilon=1e-9):
    """One tf.keras Adam update of a parameter tensor, in place (lr_t already bias-corrected)."""
    require_cuda()
    lib = _lib.load()
    check(lib.smct_adam_step(param.numel(), _ptr(param, torch.float32, "param"), _ptr(grad, torch.float32, "grad"),
                             _ptr(m, torch.float32, "m"), _ptr(v, torch.float32, "v"), float(lr_t), float(beta1),
                             float(beta2), float(epsilon), _stream()))
