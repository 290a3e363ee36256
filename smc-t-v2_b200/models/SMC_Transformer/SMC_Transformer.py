"""Mirror of src/models/SMC_Transformer/SMC_Transformer.py: the SMC_Transformer model call surface.

Same constructor, `model(inputs, targets, attention_mask=None)` return structure, `compute_SMC_loss`
and attributes (`cell`, `internal_noises`, `noise_K_resampled`, ...).  One `__call__` is one
smct_forward (C-ABI): all S particle-filter timesteps run on the GPU; torch tensors are containers.

Two input conventions are served (SURVEY.md §0 finding 1):
  * regression (north star / regression-era pyc): float inputs (B,1,S,F_x) and targets (B,1,S,F_y);
    Dense 'projection_layer_ts' encoding, Gaussian observation weights, MSE-type classic loss;
  * classification (the nlp checkout): integer tokens (B,1,S,1); Embedding encoding, softmax weights
    passed as logits, sparse cross-entropy classic loss.
The task is inferred from the input dtype unless `task=` is given.
"""
import collections
import math

import numpy as np
import torch

from ... import functional as F
from ..classic_layers import Dense, Embedding, Variable, _device
from .SMC_TransformerCell import SMC_Transf_Cell
from .self_attention_SMC import next_seed

NestedInput = collections.namedtuple("NestedInput", ["x", "y"])
NestedState = collections.namedtuple("NestedState", ["K", "V", "R"])


def _as_tensor(a):
    if isinstance(a, torch.Tensor):
        return a
    return torch.as_tensor(np.asarray(a))


class SMC_Transformer:
    def __init__(self, d_model, output_size, seq_len, full_model, dff, num_layers=1, num_heads=1,
                 maximum_position_encoding=50, rate=0., attn_window=None, init_weights=1, task=None, seed=7):
        if num_layers != 1:
            raise NotImplementedError("only the 1-layer SMC_Transformer is in scope (num_layers=%r needs the Baselines "
                                      "Decoder / GPT2Decoder, SURVEY.md §2 #13)" % (num_layers,))
        self.decoder = None
        self.init_variables = None
        self.cell = SMC_Transf_Cell(d_model=d_model, output_size=output_size, seq_len=seq_len, full_model=full_model,
                                    dff=dff, attn_window=attn_window, num_heads=num_heads, rate=rate)
        self.embedding = Embedding(input_dim=output_size, output_dim=d_model, name="embedding")  # classification
        self.projection_layer_ts = Dense(d_model, name="projection_layer_ts")                    # regression (pyc)
        self.output_size = output_size
        self.seq_len = seq_len
        self.full_model = full_model
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.d_model = d_model
        self.dff = dff
        self.task = task
        self.training = False
        # noise_z recorded for the loss: None -> "pyc" (regression era, z - dense(z_)) for float inputs,
        # "checkout" (z - z_, self_attention_SMC.py:190) for integer tokens.  Gradients need "pyc".
        self.noise_z_mode = None
        self._last_forward = None
        self._last_mask = None           # (B,S) attention mask of the last call (weights the classic loss / its gradient)
        self.algo = "auto"
        self.seed = None                 # int -> reproducible draws; None -> fresh Philox seed per call
        self._rng = np.random.default_rng(seed)
        self.cell.build(self._rng)
        self.internal_noises = None
        self.noise_K_resampled = self.noise_V_resampled = self.noise_q = self.noise_z = None
        self.loss_sums = None
        self.last_indices = None
        self.get_layers()

    # ---- bookkeeping ------------------------------------------------------------------------
    def get_layers(self):
        layers = self.cell.attention_smc.layers + self.cell.layers
        layers.append(self.cell.output_layer)
        self.layers_ = layers

    def _task(self, inputs):
        if self.task is not None:
            return self.task
        return "regression" if torch.is_floating_point(inputs) else "classification"

    def _encoder_params(self, task, F_x):
        if task == "classification":
            if self.embedding.embeddings is None:
                self.embedding.build(self._rng)
            return self.embedding.embeddings.value, torch.zeros(self.d_model, device=_device())
        if not self.projection_layer_ts.built:
            self.projection_layer_ts.build(F_x, self._rng)
        return self.projection_layer_ts.kernel.value, self.projection_layer_ts.bias.value

    @property
    def trainable_variables(self):
        out = []
        if self.embedding.embeddings is not None:
            out += self.embedding.trainable_variables
        if self.projection_layer_ts.built:
            out += self.projection_layer_ts.trainable_variables
        return out + self.cell.trainable_variables

    def save_weights(self, path):
        np.savez(path if path.endswith(".npz") else path + ".npz", **{v.name: v.numpy() for v in self.trainable_variables})

    def load_weights(self, path):
        data = np.load(path if path.endswith(".npz") else path + ".npz")
        for v in self.trainable_variables:
            if v.name in data.files:
                v.assign(data[v.name])

    def _params(self, task, F_x):
        p = self.cell.param_dict()
        p["W_in"], p["b_in"] = self._encoder_params(task, F_x)
        return p

    def _shape(self, B, S, task, F_x, F_y, b_global0=0):
        c = self.cell
        return F.make_shape(B, c.num_particles if c.noise else max(1, c.num_particles), S, self.d_model, H=self.num_heads,
                            dff=self.dff, F_x=F_x, F_y=F_y, full_model=self.full_model,
                            attn_window=c.attention_smc.attn_window, len_resampling=c.len_resampling,
                            weights="classification" if task == "classification" else "gaussian",
                            noise_z_mode=self.noise_z_mode or ("checkout" if task == "classification" else "pyc"),
                            input_mode="embedding" if task == "classification" else "linear",
                            noise=c.noise, logvar_dim=c.attention_smc.logvar_dim(), algo=self.algo,
                            sigma_obs=c.Sigma_obs, b_global0=b_global0)

    def _prep(self, inputs, targets):
        dev = _device()
        inputs, targets = _as_tensor(inputs), _as_tensor(targets)
        if inputs.dim() != 4 or targets.dim() != 4:
            raise ValueError("inputs/targets must be 4-D (B,1,S,F) like the reference's tf.data tuples")
        if inputs.shape[1] != 1 or targets.shape[1] != 1:
            raise NotImplementedError("per-particle inputs (P>1 on axis 1) are not supported by the fused forward; "
                                      "the reference tiles (B,1,S,F) inputs itself (SMC_Transformer.py:168-170)")
        task = self._task(inputs)
        B, _, S, _ = inputs.shape
        if task == "classification":
            x = inputs.reshape(B, S).to(dev, torch.int32).contiguous()
            y = targets.reshape(B, S).to(dev, torch.int32).contiguous()
            F_x, F_y = self.output_size, self.output_size
        else:
            F_x, F_y = inputs.shape[-1], targets.shape[-1]
            x = inputs.reshape(B, S, F_x).to(dev, torch.float32).contiguous()
            y = targets.reshape(B, S, F_y).to(dev, torch.float32).contiguous()
            if F_y != self.output_size:
                raise ValueError("targets have %d features but output_size=%d" % (F_y, self.output_size))
        return task, x, y, B, S, F_x, F_y

    # ---- forward ----------------------------------------------------------------------------
    def get_encoded_input(self, inputs, attention_mask=None):
        """SMC_Transformer.py:166-181 (decoder is None): tile to P particles, embed / project, * sqrt(d_model).
        Returns (B,P,S,D).  (The fused forward never materialises this tensor; provided for API parity.)"""
        inputs = _as_tensor(inputs)
        task = self._task(inputs)
        B, _, S, _ = inputs.shape
        if task == "classification":
            x, F_x = inputs.reshape(B, S).to(_device(), torch.int32).contiguous(), self.output_size
        else:
            F_x = inputs.shape[-1]
            x = inputs.reshape(B, S, F_x).to(_device(), torch.float32).contiguous()
        shape = self._shape(B, S, task, F_x, self.output_size)
        X = F.encode(shape, self._params(task, F_x), x)
        return X.unsqueeze(1).expand(B, shape.P, S, self.d_model)

    def __call__(self, inputs, targets, attention_mask=None, **kw):
        return self.call(inputs, targets, attention_mask=attention_mask, **kw)

    def call(self, inputs, targets, attention_mask=None, eps=None, u=None, i_forced=None, b_global0=0):
        """SMC_Transformer.call — SMC_Transformer.py:183-248.
        inputs (B,1,S,F_x) float | (B,1,S,1) int ; targets (B,1,S,F_y) | (B,1,S,1) int.
        Returns ((pred, pred_resampl), (K, V, R_resampl), last_filtering_weights) with the reference's shapes
        (B,P,S,F_y) / (B,P,S,D) / (B,P,S).  Keyword-only extras inject the random draws (tests)."""
        # attention_mask (B,1,S,1): with num_layers == 1 the reference's forward does not look at it (only the multi-layer
        # decoder does, SMC_Transformer.py:172-181); it weights the classic loss and the metric (compute_classic_loss)
        task, x, y, B, S, F_x, F_y = self._prep(inputs, targets)
        self._last_mask = self._mask_bs(attention_mask, B, S)
        if S != self.seq_len:
            # the reference sizes its state with self.seq_len (SMC_Transformer.py:197); inference mutates it (run_SMC_T.py:191)
            raise ValueError("sequence length %d != model.seq_len %d" % (S, self.seq_len))
        shape = self._shape(B, S, task, F_x, F_y, b_global0)
        dev = _device()
        cvt = lambda a, dt: None if a is None else _as_tensor(a).to(dev, dt).contiguous()
        want = ("i_out", "noise_q", "noise_z", "noise_K", "noise_V", "loss_sums") if self.cell.noise else ("noise_K", "noise_V")
        params = self._params(task, F_x)
        seed = self.seed if self.seed is not None else next_seed()
        eps_t = cvt(eps, torch.float32)
        out = F.forward(shape, params, x, y, eps=eps_t, u=cvt(u, torch.float64), i_forced=cvt(i_forced, torch.int32),
                        seed=seed, want=want)
        # the log-variances the forward USED (an EM update overwrites the variables in place before the backward runs)
        logvar_fwd = {n: params[n].clone() for n in ("logvar_k", "logvar_q", "logvar_v", "logvar_z")}
        self._last_forward = dict(shape=shape, params=params, x=x, y=y, out=out, eps=eps_t, seed=seed, logvar_fwd=logvar_fwd)
        self.cell.dec_timestep = 0   # SMC_Transformer.py:214-215
        self.cell.cell_count = 0
        self.noise_K_resampled = out["noise_K"]
        self.noise_V_resampled = out["noise_V"]
        if self.cell.noise:
            self.noise_q, self.noise_z = out["noise_q"], out["noise_z"]
            self.internal_noises = [self.noise_K_resampled, self.noise_q, self.noise_V_resampled, self.noise_z]
            self.loss_sums = out["loss_sums"]
            self.last_indices = out["i_out"]
        self._last_task = task
        self._last_targets = y
        return (out["pred"], out["pred_resampl"]), (out["K"], out["V"], out["R"]), out["w"]

    # ---- loss -------------------------------------------------------------------------------
    def compute_log_gaussian_density(self, logvar, noise):
        """SMC_Transformer.py:91-99: -log N(noise; 0, diag(exp(logvar))) summed over D, returned as the SUM over
        all (b,p,s) rows (fp64 scalar tensor) plus the row count, so callers can form the mean."""
        from .self_attention_SMC import logvar_vector
        lv = logvar_vector(logvar)
        D = noise.shape[-1]
        rows = noise.numel() // D
        quad = F.sumsq_scaled(noise.contiguous(), lv)
        logdet = float(lv.sum().item()) * (1.0 if lv.numel() == D else D)
        return 0.5 * quad + rows * (0.5 * D * math.log(2 * math.pi) + 0.5 * logdet), rows

    @staticmethod
    def _mask_bs(attention_mask, B, S):
        """(B,1,S,1) attention mask -> (B,S) fp32 device tensor (None stays None)."""
        if attention_mask is None:
            return None
        m = _as_tensor(attention_mask).to(_device(), torch.float32).reshape(B, S).contiguous()
        return m

    def compute_classic_loss(self, targets, predictions, attention_mask=None):
        """SMC_Transformer.py:148-164 (sparse CE, classification) or the regression-era
        mean(0.5*||y-pred||^2/Sigma_obs) (pyc).  With an attention_mask the mean is the masked mean of :158-163:
        sum(loss * mask) / sum(mask), the mask tiled over the particles."""
        task = "regression" if torch.is_floating_point(_as_tensor(targets)) else "classification"
        pred = predictions.to(_device(), torch.float32).contiguous()
        B, P, S, Fy = pred.shape
        t = _as_tensor(targets)
        mask = self._mask_bs(attention_mask, B, S)
        denom = float(B * P * S) if mask is None else float(P) * mask.sum(dtype=torch.float64)
        if task == "classification":
            y = t.reshape(B, S).to(_device(), torch.int32).contiguous()
            return F.classic_loss_sum(pred, y, "classification", mask=mask) / denom
        y = t.reshape(B, S, Fy).to(_device(), torch.float32).contiguous()
        return F.classic_loss_sum(pred, y, "gaussian", mask=mask) * (0.5 / self.cell.Sigma_obs) / denom

    def compute_SMC_loss(self, inputs, targets, predictions, attention_mask=None, variant=None):
        """compute_SMC_loss — SMC_Transformer.py:123-146.  Returns (total_loss, classic_loss) as fp64 scalar tensors.
        variant "checkout" (default for classification): full -log N terms; "pyc" (default for regression):
        0.5*||n||^2/sigma^2 without the log-determinant terms (regression-era code)."""
        assert self.cell.noise == self.cell.attention_smc.noise
        classic = self.compute_classic_loss(targets, predictions, attention_mask)
        if not self.cell.noise:
            return classic, classic
        task = "regression" if torch.is_floating_point(_as_tensor(targets)) else "classification"
        variant = variant or ("pyc" if task == "regression" else "checkout")
        att = self.cell.attention_smc
        logvars = [att.logvar_k, att.logvar_q, att.logvar_v, att.logvar_z]
        total = 0.0
        rows = None
        for noise, lv in zip(self.internal_noises, logvars):
            if variant == "checkout":
                part, rows = self.compute_log_gaussian_density(lv, noise)
            else:
                from .self_attention_SMC import logvar_vector
                part = 0.5 * F.sumsq_scaled(noise.contiguous(), logvar_vector(lv))
                rows = noise.numel() // noise.shape[-1]
            total = total + part
        smc_loss = total / float(rows)
        return (smc_loss + classic).reshape(()), classic.reshape(())
