"""Mirror of src/train/utils.py: training metric and the EM variance update (host-side bookkeeping on
tensors the forward already produced; small reductions, not the hot path)."""
import math

import torch

from .. import functional as F
from ..models.SMC_Transformer.self_attention_SMC import logvar_vector


def _mask_bs(attention_mask, B, S, device):
    if attention_mask is None:
        return None
    return torch.as_tensor(attention_mask).to(device, torch.float32).reshape(B, S).contiguous()


def compute_categorical_cross_entropy(targets, preds, num_particles, attention_mask=None):
    """train/utils.py:3-20: sparse CE of the particle-mean softmax (classification era), masked mean with an
    attention_mask (:15-19; the mask is tiled over the particles there but the loss has one row per (b,s), so numerator
    and denominator both carry the factor P).  Computed by smct_metric (C-ABI)."""
    preds = preds.to(torch.float32).contiguous()
    B, P, S, V = preds.shape
    y = torch.as_tensor(targets).reshape(B, S).to(preds.device, torch.int32).contiguous()
    sums = F.metric_sums(preds, y, "classification", mask=_mask_bs(attention_mask, B, S, preds.device))
    return sums[0] / sums[1]


def compute_mse_metric(targets, preds, attention_mask=None):
    """Regression-era metric (pyc train_step_SMC_T): MSE of the particle-mean prediction.  Computed by smct_metric."""
    preds = preds.to(torch.float32).contiguous()
    B, P, S, Fy = preds.shape
    y = torch.as_tensor(targets).to(preds.device, torch.float32).reshape(B, S, Fy).contiguous()
    sums = F.metric_sums(preds, y, "gaussian", mask=_mask_bs(attention_mask, B, S, preds.device))
    return sums[0] / (sums[1] * Fy)


def EM_update(err, logvar, it, EM_param):
    """train/utils.py:48-51: var <- (1 - it^-a) exp(logvar) + it^-a err ; returns log(var)."""
    g = float(it) ** (-EM_param)
    return math.log((1.0 - g) * math.exp(logvar) + g * err)


def EM(smc_transformer, it, EM_param):
    """train/utils.py:22-46: per batch element j (in order) update the four scalar log-variances from the mean squared
    noises.  One smct_em_update call (two kernels, no host round trip); only scalar ('uni') variances, like the
    reference's EM mode."""
    if not smc_transformer.cell.noise:
        return smc_transformer
    att = smc_transformer.cell.attention_smc
    noises = [smc_transformer.noise_K_resampled.contiguous(), smc_transformer.noise_q.contiguous(),
              smc_transformer.noise_V_resampled.contiguous(), smc_transformer.noise_z.contiguous()]
    logvars = [logvar_vector(getattr(att, "logvar_" + n)) for n in ("k", "q", "v", "z")]   # views of the variables' storage
    F.em_update(noises, logvars, it, EM_param)
    return smc_transformer
