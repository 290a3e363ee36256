"""Mirror of src/models/SMC_Transformer/transformer_utils.py (hot-path part: `resample`)."""
import numpy as np
import torch

from ... import functional as F


def _to_cuda(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(device="cuda", dtype=dtype).contiguous()
    return torch.as_tensor(np.asarray(x), device="cuda").to(dtype).contiguous()


def resample(params, i_t):
    """resample(params, i_t) — transformer_utils.py:39-47: tf.gather(params, i_t, axis=1, batch_dims=1).
    params (B,P,...) float, i_t (B,P) integer ancestor indices -> out[b,m] = params[b, i_t[b,m]]."""
    params = _to_cuda(params, torch.float32)
    i_t = _to_cuda(i_t, torch.int32)
    if params.dim() < 2 or tuple(i_t.shape) != tuple(params.shape[:2]):
        raise ValueError("resample: params (B,P,...) and i_t (B,P) expected, got %s and %s" % (tuple(params.shape), tuple(i_t.shape)))
    if int(i_t.min()) < 0 or int(i_t.max()) >= params.shape[1]:
        raise IndexError("resample: indices out of range [0, %d)" % params.shape[1])  # tf.gather raises on CPU
    return F.gather(params, i_t)


def create_look_ahead_mask(size):
    """transformer_utils.py:34-36 (only used by the multi-layer Decoder, which is out of scope)."""
    return 1.0 - np.tril(np.ones((size, size), np.float32))
