"""Mirror of src/train/train_step_functions.py:31-68 (train_step_SMC_T).

The forward pass, the SMC loss, the EM variance update and the metric run on the CUDA path.  The
gradient step needs the backward pass of the particle filter (reverse-time sweep with scatter-add over
ancestors), which is SURVEY.md §8(f) rank 1 and not built yet: passing an optimizer raises."""
import torch

from .utils import EM, compute_categorical_cross_entropy, compute_mse_metric


def train_step_SMC_T(inputs, targets, smc_transformer, optimizer, it, attention_mask=None, EM_param=None):
    """Same signature/return as the reference: (loss, metric-or-None)."""
    (preds, preds_resampl), _, _ = smc_transformer(inputs=inputs, targets=targets, attention_mask=attention_mask)
    if EM_param is not None:
        smc_transformer = EM(smc_transformer, it=it, EM_param=EM_param)
    smc_loss, classic_loss = smc_transformer.compute_SMC_loss(predictions=preds_resampl, inputs=inputs, targets=targets,
                                                              attention_mask=attention_mask)
    loss = smc_loss
    if torch.is_floating_point(torch.as_tensor(targets)):
        metric = compute_mse_metric(targets, preds)
    else:
        metric = compute_categorical_cross_entropy(targets=targets, preds=preds,
                                                   num_particles=smc_transformer.cell.num_particles,
                                                   attention_mask=attention_mask)
    if optimizer is not None:
        raise NotImplementedError(
            "gradient step: the backward pass of the SMC forward is not built yet (SURVEY.md §8(f) rank 1); "
            "call with optimizer=None for forward + loss (+ EM variance update)")
    if smc_transformer.cell.noise:
        return loss, metric
    return loss, None
