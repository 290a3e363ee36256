"""Mirror of src/train/train_functions.py:146-243 (train_SMC_transformer): the epoch / validation loop around
train_step_SMC_T.  Same signature, same history keys, same CSV file name; every forward / backward / optimiser
step runs on the CUDA path.  `logger` and `ckpt_manager` may be None (logging and tf.train checkpoints are the
reference's control plane, SURVEY.md §2 — out of scope); weights are saved with model.save_weights like the reference."""
import csv
import os
import time

import numpy as np
import torch

from .train_step_functions import train_step_SMC_T
from .utils import compute_categorical_cross_entropy, compute_mse_metric


def _f(x):
    return float(x.item()) if isinstance(x, torch.Tensor) else float(x)


def _info(logger, msg):
    if logger is not None:
        logger.info(msg)


def _unpack(batch):
    """The reference's tf.data pipelines yield (inputs, targets, attention_mask); 2-tuples are accepted as mask=None."""
    if len(batch) == 3:
        return batch
    inp, tar = batch
    return inp, tar, None


def _log_model_variances(smc_transformer, logger):
    """train_functions.py:128-143: log the current noise variances."""
    att = smc_transformer.cell.attention_smc
    from ..models.SMC_Transformer.self_attention_SMC import logvar_vector
    vals = {n: np.exp(logvar_vector(getattr(att, "logvar_" + n)).detach().cpu().numpy()) for n in ("k", "q", "v", "z")}
    _info(logger, "variances k,q,v,z: {}".format({n: (float(v[0]) if v.size == 1 else v.tolist()) for n, v in vals.items()}))


def saving_training_history(keys, values, output_path, csv_fname, logger, start_epoch):
    """utils_train.saving_training_history: one CSV column per history key."""
    if output_path is None:
        return None
    path = os.path.join(output_path, csv_fname)
    rows = max((len(v) for v in values), default=0)
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["epoch"] + list(keys))
        for i in range(rows):
            w.writerow([start_epoch + i + 1] + [(v[i] if i < len(v) else "") for v in values])
    _info(logger, "saving loss and metrics information in {}".format(path))
    return path


def train_SMC_transformer(smc_transformer, optimizer, EPOCHS, train_dataset, val_dataset, output_path, ckpt_manager, logger,
                          start_epoch, num_train, EM_param=None):
    """Returns the losses_history dict (the reference writes it to `smc_t_history_{num_train}.csv`; so does this)."""
    losses_history = {"train_loss": [], "train_mse_metric": [], "train_ppl": [], "val_loss": [], "val_mse_metric": [],
                      "val_ppl": []}
    print("number of trainable variables:", len(smc_transformer.trainable_variables))
    if start_epoch > 0:
        if start_epoch > EPOCHS:
            _info(logger, "adding {} more epochs to existing training".format(EPOCHS))
            start_epoch = 0
        else:
            _info(logger, "starting training after checkpoint restoring from epoch {}".format(start_epoch))
    it = 0
    start = time.time()
    noise = smc_transformer.cell.noise
    for epoch in range(start_epoch, EPOCHS):
        start = time.time()
        _info(logger, "Epoch {}/{}".format(epoch + 1, EPOCHS))
        train_loss = [0., 0.] if noise else [0.]
        val_loss = [0., 0.] if noise else [0.]
        batch = -1
        for batch, item in enumerate(train_dataset):
            inp, tar, attn_mask = _unpack(item)
            it += 1
            train_loss_batch, train_metric_avg_pred = train_step_SMC_T(inputs=inp, targets=tar, smc_transformer=smc_transformer,
                                                                       optimizer=optimizer, it=it, attention_mask=attn_mask,
                                                                       EM_param=EM_param)
            train_loss[0] += _f(train_loss_batch)
            if noise:
                train_loss[1] += _f(train_metric_avg_pred)
        if output_path is not None:
            smc_transformer.save_weights(os.path.join(output_path, "model"))
        batch_val = -1
        for batch_val, item in enumerate(val_dataset):
            inp, tar, attn_mask = _unpack(item)
            (preds_val, preds_val_resampl), _, _ = smc_transformer(inputs=inp, targets=tar, attention_mask=attn_mask)
            val_loss_batch, _ = smc_transformer.compute_SMC_loss(inputs=inp, targets=tar, predictions=preds_val_resampl)   # (the reference passes no mask here, train_functions.py:207)
            val_loss[0] += _f(val_loss_batch)
            if noise:
                if torch.is_floating_point(torch.as_tensor(np.asarray(tar)) if not isinstance(tar, torch.Tensor) else tar):
                    val_metric = compute_mse_metric(tar, preds_val, attention_mask=attn_mask)   # regression era (pyc): MSE of the particle mean
                else:
                    val_metric = compute_categorical_cross_entropy(targets=tar, preds=preds_val,
                                                                   num_particles=smc_transformer.cell.num_particles,
                                                                   attention_mask=attn_mask)
                val_loss[1] += _f(val_metric)
        train_loss = [v / max(batch + 1, 1) for v in train_loss]
        val_loss = [v / max(batch_val + 1, 1) for v in val_loss]
        _info(logger, "train loss: {} - val loss: {}".format(train_loss[0], val_loss[0]))
        losses_history["train_loss"].append(train_loss[0])
        losses_history["val_loss"].append(val_loss[0])
        if noise:
            _info(logger, "train mse metric from avg particule: {} - val mse metric from avg particule: {}".format(
                train_loss[1], val_loss[1]))
            losses_history["train_mse_metric"].append(train_loss[1])
            losses_history["val_mse_metric"].append(val_loss[1])
            losses_history["train_ppl"] = list(np.exp(losses_history["train_mse_metric"]))
            losses_history["val_ppl"] = list(np.exp(losses_history["val_mse_metric"]))
            _log_model_variances(smc_transformer, logger)
        else:
            losses_history["train_ppl"] = list(np.exp(losses_history["train_loss"]))
            losses_history["val_ppl"] = list(np.exp(losses_history["val_loss"]))
        if ckpt_manager is not None:
            ckpt_manager.save()
        _info(logger, "Time taken for 1 epoch: {} secs".format(time.time() - start))
    saving_training_history(keys=list(losses_history.keys()), values=list(losses_history.values()), output_path=output_path,
                            csv_fname="smc_t_history_{}.csv".format(num_train), logger=logger, start_epoch=start_epoch)
    _info(logger, "total training time for {} epochs:{}".format(EPOCHS, time.time() - start))
    return losses_history
