"""Mirror of src/train/train_step_functions.py:31-68 (train_step_SMC_T).

Forward, SMC loss, EM variance update, metric, gradients (smct_backward) and the Adam step all run on the CUDA
path.  With several ranks (torch.distributed initialised) the gradients of the batch shards are all-reduced over
NCCL before the update (one flat buffer), so every rank applies the same step."""
import numpy as np
import torch

from .. import distributed as D
from .. import functional as F
from .utils import EM, compute_categorical_cross_entropy, compute_mse_metric


def compute_gradients(smc_transformer, inputs, targets, inv_n=None, loss_variant="pyc", classic_w=None):
    """tape.gradient(loss, trainable_variables) for the forward pass that was just run on (inputs, targets).
    Returns {parameter name: gradient tensor} keyed like smct_params.

    The forward is replayed with the noise scales it actually used: `_last_forward["logvar_fwd"]` holds clones of the
    four log-variances taken at forward time.  smct_backward regenerates the q and z noises from logvar_q / logvar_z
    (-> the forward's values) and weights the K / V noise terms of the loss with exp(-logvar_k), exp(-logvar_v)
    (-> the CURRENT values, i.e. after an EM update, as compute_SMC_loss sees them — train_step_functions.py:43-54).
    classic_w (B,S): weight of the classic-loss term of every row relative to the uniform mean (masked losses)."""
    m = smc_transformer
    fwd = m._last_forward
    if fwd is None:
        raise RuntimeError("call the model on (inputs, targets) before compute_gradients")
    params = dict(fwd["params"])
    snap = fwd.get("logvar_fwd")
    if snap is not None:
        params["logvar_q"], params["logvar_z"] = snap["logvar_q"], snap["logvar_z"]
    return F.backward(fwd["shape"], params, fwd["x"], fwd["y"], fwd["out"], eps=fwd["eps"], seed=fwd["seed"],
                      inv_n=inv_n, loss_variant=loss_variant, classic_w=classic_w)


def _is_regression(targets):
    t = targets if isinstance(targets, torch.Tensor) else torch.as_tensor(np.asarray(targets))
    return torch.is_floating_point(t)


def train_step_SMC_T(inputs, targets, smc_transformer, optimizer, it, attention_mask=None, EM_param=None):
    """Same signature/return as the reference: (loss, metric-or-None).  optimizer=None: forward + loss only."""
    regression = _is_regression(targets)
    if optimizer is not None and smc_transformer.cell.noise:
        # decide BEFORE the forward: the gradient of the checkout's noise_z = z - z_ term (self_attention_SMC.py:190)
        # involves particles outside the final genealogy and is not served by smct_backward
        mode = smc_transformer.noise_z_mode or ("pyc" if regression else "checkout")
        if mode != "pyc":
            raise NotImplementedError('training needs model.noise_z_mode = "pyc" (noise_z = z - dense(z_), the regression-era '
                                      'form); the checkout\'s noise_z = z - z_ loss term has no gradient on this path')
    (preds, preds_resampl), _, _ = smc_transformer(inputs=inputs, targets=targets, attention_mask=attention_mask)
    if EM_param is not None:
        smc_transformer = EM(smc_transformer, it=it, EM_param=EM_param)
    smc_loss, classic_loss = smc_transformer.compute_SMC_loss(predictions=preds_resampl, inputs=inputs, targets=targets,
                                                              attention_mask=attention_mask)
    loss = smc_loss
    if regression:
        metric = compute_mse_metric(targets, preds, attention_mask=attention_mask)
    else:
        metric = compute_categorical_cross_entropy(targets=targets, preds=preds,
                                                   num_particles=smc_transformer.cell.num_particles,
                                                   attention_mask=attention_mask)
    if optimizer is not None:
        rank, world = D.world()
        B, P, S = preds.shape[0], preds.shape[1], preds.shape[2]
        # regression era (pyc): no log-determinant terms; the nlp checkout: full -log N terms + sparse cross-entropy
        classic_w = None
        mask = smc_transformer._last_mask
        if mask is not None:   # masked mean of compute_classic_loss: mask[b,s] * (global B*S) / (global sum of the mask)
            tot = mask.sum(dtype=torch.float64).reshape(1)
            D.allreduce_sums(tot)
            classic_w = (mask * (float(B) * world * S / float(tot.item()))).contiguous()
        grads = compute_gradients(smc_transformer, inputs, targets, inv_n=1.0 / (float(B) * world * P * S),
                                  loss_variant="pyc" if regression else "checkout", classic_w=classic_w)
        params = smc_transformer._last_forward["params"]
        # optimizer.apply_gradients(zip(gradients, trainable_variables)): only TRAINABLE variables are updated — in EM
        # mode the log-variances are plain tensors outside trainable_variables (self_attention_SMC.py:48-52)
        trainable = {v.value.data_ptr() for v in smc_transformer.trainable_variables}
        att = smc_transformer.cell.attention_smc

        def is_trainable(n):
            if n.startswith("logvar_"):   # "multi" log-variances are (D,D) diagonal matrices: the kernels see a copy of the diagonal
                return bool(getattr(att, n).trainable)
            return params[n].data_ptr() in trainable
        names = sorted(n for n in grads if is_trainable(n))
        if world > 1:
            flat = torch.cat([grads[n].reshape(-1) for n in names])
            D.allreduce_sums(flat)
            o = 0
            for n in names:
                k = grads[n].numel()
                grads[n].copy_(flat[o:o + k].reshape(grads[n].shape))
                o += k
        optimizer.apply_gradients((grads[n], params[n]) for n in names)
        for n in names:   # write the updated diagonal back into the (D,D) variable of noise_dim = "multi"
            if n.startswith("logvar_") and getattr(att, n).value.dim() == 2:
                getattr(att, n).value.diagonal().copy_(params[n])
    if smc_transformer.cell.noise:
        return loss, metric
    return loss, None
