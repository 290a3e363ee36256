"""Mirror of src/models/SMC_Transformer/self_attention_SMC.py: the stochastic self-attention layer.
Same constructor, attributes and call signature; the computation is smct_attention_step (C-ABI)."""
import math

import numpy as np
import torch

from ... import functional as F
from ..classic_layers import Dense, Variable, _device

_seed_counter = [0x5eed]


def next_seed():
    """Fresh Philox seed per stochastic call (the reference draws from TF's global generator)."""
    _seed_counter[0] += 1
    return _seed_counter[0]


def set_global_seed(seed):
    _seed_counter[0] = int(seed)


def logvar_vector(logvar):
    """Scalar log-variance -> (1,) ; (D,D) diagonal matrix (noise_dim='multi', self_attention_SMC.py:41-44,59-62)
    -> its diagonal (D,) — what add_noise uses through tf.linalg.diag_part (:104)."""
    v = logvar.value if isinstance(logvar, Variable) else torch.as_tensor(np.asarray(logvar, np.float32), device=_device())
    if v.dim() == 2:
        if not isinstance(logvar, Variable):
            return torch.diagonal(v).contiguous()
        # a persistent (D,) buffer per variable (stable data pointer: the optimiser keys its moments by it), refreshed from
        # the matrix on every call; the training step writes the updated diagonal back into the matrix
        d = getattr(logvar, "_diag", None)
        if d is None or d.numel() != v.shape[0]:
            d = torch.diagonal(v).contiguous().clone()
            logvar._diag = d
        else:
            d.copy_(torch.diagonal(v))
        return d
    return v.reshape(-1).contiguous()


class Self_Attention_SMC:
    def __init__(self, d_model, num_heads=1, attn_window=None, init_variables=None):
        if init_variables is not None:
            raise NotImplementedError("GPT-2 weight injection (init_variables) is out of scope (SURVEY.md §2 #5)")
        self.d_model = d_model
        self.attn_window = attn_window
        self.wq = Dense(d_model, name="dense_projection_q")
        self.wk = Dense(d_model, name="dense_projection_k")
        self.wv = Dense(d_model, name="dense_projection_v")
        self.dense = Dense(d_model, name="dense_projection_z")
        self.noise_network = None
        self.layers = [self.wq, self.wk, self.wv, self.dense]
        self.noise = False
        self.num_heads = num_heads
        assert d_model % self.num_heads == 0
        self.depth = d_model // self.num_heads
        self.logvar_k = self.logvar_q = self.logvar_v = self.logvar_z = None
        self.noise_k = self.noise_q = self.noise_v = self.noise_z = None
        self.seed = None  # set to an int for reproducible draws

    def build(self, rng):
        for l in self.layers:
            if not l.built:
                l.build(self.d_model, rng)

    def diagonal_variance_matrix(self, sigmas):
        return np.diag(np.log(np.asarray(sigmas, np.float32)))

    def add_SMC_parameters(self, dict_sigmas, EM=False):
        """self_attention_SMC.py:46-67.  `sigmas` are VARIANCES: logvar = log(sigma)."""
        if isinstance(dict_sigmas["k"], str):
            raise NotImplementedError("noise_net variance is out of scope (SURVEY.md §8 a5)")
        for n in ("k", "q", "v", "z"):
            s = dict_sigmas[n]
            if isinstance(s, (list, tuple, np.ndarray)) and np.ndim(s) == 1:
                val = self.diagonal_variance_matrix(s)
            else:
                val = np.asarray(math.log(float(s)), np.float32)
            setattr(self, "logvar_" + n, Variable(val, "logvar_" + n, trainable=not EM))
        self.noise = True

    @property
    def trainable_variables(self):
        out = []
        for l in self.layers:
            out += l.trainable_variables
        if self.noise:
            out += [v for v in (self.logvar_k, self.logvar_q, self.logvar_v, self.logvar_z) if v is not None and v.trainable]
        return out

    def param_dict(self):
        p = dict(Wq=self.wq.kernel.value, bq=self.wq.bias.value, Wk=self.wk.kernel.value, bk=self.wk.bias.value,
                 Wv=self.wv.kernel.value, bv=self.wv.bias.value, Wz=self.dense.kernel.value, bz=self.dense.bias.value)
        if self.noise:
            for n in ("k", "q", "v", "z"):
                p["logvar_" + n] = logvar_vector(getattr(self, "logvar_" + n))
        else:
            z = torch.zeros(1, device=_device())
            for n in ("k", "q", "v", "z"):
                p["logvar_" + n] = z
        return p

    def logvar_dim(self):
        return int(logvar_vector(self.logvar_k).numel()) if self.noise else 1

    def __call__(self, inputs, timestep, K, V, eps4=None):
        return self.call(inputs, timestep, K, V, eps4=eps4)

    def call(self, inputs, timestep, K, V, eps4=None):
        """self_attention_SMC.py:140-193.  inputs (B,P,1,D); K,V (B,P,S,D).
        Returns ((z, K, V), attention_weights) with z (B,P,1,D), attention_weights (B,P,H,1,S).
        `eps4` (4,B,P,D) optionally injects the N(0,1) draws (order k,q,v,z)."""
        inputs = torch.as_tensor(inputs, dtype=torch.float32, device=_device()) if not isinstance(inputs, torch.Tensor) \
            else inputs.to(_device(), torch.float32)
        assert inputs.dim() == 4, "inputs must be (B,P,1,D)"  # :148
        B, P, _, D = inputs.shape
        if not self.wq.built:
            self.build(np.random.default_rng(next_seed()))
        K = K.to(_device(), torch.float32).contiguous().clone()
        V = V.to(_device(), torch.float32).contiguous().clone()
        S = K.shape[2]
        p = self.param_dict()
        dummy = torch.ones(D, device=_device())
        p.update(ln1_g=dummy, ln1_b=dummy, Wo=dummy)
        shape = F.make_shape(B, P, S, D, H=self.num_heads, dff=0, F_x=1, F_y=1, full_model=False,
                             attn_window=self.attn_window, noise=self.noise, logvar_dim=self.logvar_dim())
        if eps4 is not None:
            eps4 = torch.as_tensor(eps4, dtype=torch.float32, device=_device()).contiguous()
        out = F.attention_step(shape, p, int(timestep), inputs.reshape(B, P, D).contiguous(), K, V, eps4=eps4,
                               seed=self.seed if self.seed is not None else next_seed())
        if self.noise:
            self.noise_k = out["noise_k"].unsqueeze(2)
            self.noise_q = out["noise_q"].unsqueeze(2)
            self.noise_v = out["noise_v"].unsqueeze(2)
            self.noise_z = out["noise_z"].unsqueeze(2)
        return (out["z"].unsqueeze(2), K, V), out["attn"].unsqueeze(3)
