"""Mirror of src/models/SMC_Transformer/SMC_TransformerCell.py: one particle-filter timestep.
Same constructor, attributes and `call(inputs, states)` contract; computed by smct_cell_step (C-ABI)."""
import collections

import numpy as np
import torch

from ... import functional as F
from ..classic_layers import Dense, LayerNormalization, _device, point_wise_feed_forward_network
from .self_attention_SMC import Self_Attention_SMC, next_seed
from .transformer_utils import resample  # noqa: F401  (same import as the reference module)

NestedInput = collections.namedtuple("NestedInput", ["x", "y"])
NestedState = collections.namedtuple("NestedState", ["K", "V", "R"])


class SMC_Transf_Cell:
    def __init__(self, d_model, output_size, seq_len, full_model, dff, num_heads=1, attn_window=None,
                 init_variables=None, rate=0., **kwargs):
        if init_variables is not None:
            raise NotImplementedError("GPT-2 weight injection (init_variables) is out of scope (SURVEY.md §2 #5)")
        self.dec_timestep = 0   # SMC_TransformerCell.py:24-25
        self.cell_count = 0
        self.attention_smc = Self_Attention_SMC(d_model=d_model, num_heads=num_heads, attn_window=attn_window)
        self.d_model = d_model
        self.output_size = output_size
        self.seq_len = seq_len
        self.full_model = full_model
        self.dff = dff
        self.init_variables = None
        self.layernorm1 = LayerNormalization(epsilon=1e-6, name="layer_norm1")
        self.layernorm2 = LayerNormalization(epsilon=1e-6, name="layer_norm2")
        self.ffn = point_wise_feed_forward_network(d_model, dff)
        self.output_layer = Dense(output_size, use_bias=False, name="output_layer")
        self.layers = [self.layernorm1, self.layernorm2, self.ffn]
        self.rate = rate        # dropout is never active in the reference (cell.training stays False, :58)
        self.num_particles = 1
        self.noise = False
        self.len_resampling = None
        self.training = False
        self.Sigma_obs = 0.5    # observation variance of the Gaussian weight (regression era; pyc)
        self.task = None        # "regression" | "classification"; inferred from the target dtype if None
        self.seed = None
        self.state_size = NestedState(K=(self.num_particles, seq_len, d_model), V=(self.num_particles, seq_len, d_model),
                                      R=(self.num_particles, seq_len, d_model))

    # ---- parameters -------------------------------------------------------------------------
    def build(self, rng):
        self.attention_smc.build(rng)
        if self.layernorm1.gamma is None:
            self.layernorm1.build(self.d_model)
            self.layernorm2.build(self.d_model)
            self.ffn.build(self.d_model, rng)
            self.output_layer.build(self.d_model, rng)

    @property
    def built(self):
        return self.layernorm1.gamma is not None and self.attention_smc.wq.built

    def add_SMC_parameters(self, dict_sigmas, num_particles, EM=False, sigma_obs=None):
        """SMC_TransformerCell.py:72-76 (+ the regression-era `sigma_obs` argument recovered from the pyc)."""
        self.noise = True
        self.attention_smc.add_SMC_parameters(dict_sigmas=dict_sigmas, EM=EM)
        self.num_particles = num_particles
        if sigma_obs is not None:
            self.Sigma_obs = float(sigma_obs)

    def add_stop_resampling(self, len_resampling):
        assert self.noise  # :125-127
        self.len_resampling = len_resampling

    @property
    def trainable_variables(self):
        out = self.attention_smc.trainable_variables
        for l in self.layers:
            out += l.trainable_variables
        return out + self.output_layer.trainable_variables

    def param_dict(self):
        p = self.attention_smc.param_dict()
        p.update(ln1_g=self.layernorm1.gamma.value, ln1_b=self.layernorm1.beta.value,
                 ln2_g=self.layernorm2.gamma.value, ln2_b=self.layernorm2.beta.value,
                 W1=self.ffn.dense1.kernel.value, b1=self.ffn.dense1.bias.value,
                 W2=self.ffn.dense2.kernel.value, b2=self.ffn.dense2.bias.value,
                 Wo=self.output_layer.kernel.value)
        return p

    # ---- weights ----------------------------------------------------------------------------
    def compute_w_classification(self, predictions, y):
        """SMC_TransformerCell.py:78-84: w = softmax(predictions)[y], shape (B,P) (unnormalised over particles).
        predictions (B,P,1,V); y (B,P,1,1) integer (identical across particles)."""
        pred = predictions.to(_device(), torch.float32).reshape(predictions.shape[0], predictions.shape[1], -1).contiguous()
        yb = y.reshape(y.shape[0], -1)[:, 0].to(_device(), torch.int32).contiguous()
        return F.logweights(pred, yb, "classification")

    def compute_w_regression(self, predictions, y):
        """Gaussian observation weight (call site SMC_TransformerCell.py:206; north-star form):
        log-weights -0.5||y-pred||^2/Sigma_obs, normalised over particles; returns w = softmax_p(logw) (B,P)."""
        pred = predictions.to(_device(), torch.float32).reshape(predictions.shape[0], predictions.shape[1], -1).contiguous()
        yb = y.to(_device(), torch.float32).reshape(y.shape[0], y.shape[1], -1)[:, 0].contiguous()
        logw = F.logweights(pred, yb, "gaussian", self.Sigma_obs)
        zero_u = torch.zeros(logw.shape, dtype=torch.float64, device=logw.device)
        _, w, _ = F.resample_indices(logw, zero_u, normalise=True)
        return w

    # ---- one timestep -----------------------------------------------------------------------
    def __call__(self, inputs, states, **kw):
        return self.call(inputs, states, **kw)

    def call(self, inputs, states, eps4=None, u=None, i_forced=None):
        """SMC_TransformerCell.py:129-184.  inputs = NestedInput(x (B,P,D), y (B,P,F_y)); states = (K,V,R) (B,P,S,D).
        Returns (output, NestedState) with output = [r (B,P,1,D), attn (B,P,H,1,S), [noise_q, noise_z], w (B,P)]
        (or [r, attn] when noise is off).  Extra keyword arguments inject the random draws."""
        x, y = inputs
        K, V, R = states
        dev = _device()
        x = x.to(dev, torch.float32).contiguous()
        B, P, D = x.shape
        S = K.shape[2]
        if not self.built:
            self.build(np.random.default_rng(next_seed()))
        task = self.task or ("classification" if not torch.is_floating_point(y) else "regression")
        if task == "classification":
            yb = y.reshape(B, P, -1)[:, 0, 0].to(dev, torch.int32).contiguous()
            Fy, weights = self.output_size, "classification"
        else:
            yb = y.to(dev, torch.float32).reshape(B, P, -1)[:, 0].contiguous()
            Fy, weights = yb.shape[-1], "gaussian"
        shape = F.make_shape(B, P, S, D, H=self.attention_smc.num_heads, dff=self.dff, F_x=1, F_y=Fy,
                             full_model=self.full_model, attn_window=self.attention_smc.attn_window,
                             len_resampling=self.len_resampling, weights=weights, noise=self.noise,
                             logvar_dim=self.attention_smc.logvar_dim(), sigma_obs=self.Sigma_obs)
        K, V, R = (a.to(dev, torch.float32).contiguous().clone() for a in (K, V, R))
        cvt = lambda a, dt: None if a is None else torch.as_tensor(a, device=dev).to(dt).contiguous()
        out = F.cell_step(shape, self.param_dict(), self.dec_timestep, x, yb, K, V, R, eps4=cvt(eps4, torch.float32),
                          u=cvt(u, torch.float64), i_forced=cvt(i_forced, torch.int32),
                          seed=self.seed if self.seed is not None else next_seed())
        r = out["r"].unsqueeze(2)
        attn = out["attn"].unsqueeze(3)
        if self.noise:
            self.attention_smc.noise_q = out["noise_q"].unsqueeze(2)
            self.attention_smc.noise_z = out["noise_z"].unsqueeze(2)
            self.last_indices = out["i_t"]
            output = [r, attn, [self.attention_smc.noise_q, self.attention_smc.noise_z], out["w"]]
        else:
            output = [r, attn]
        self.cell_count += 1            # :180-182 — K.rnn calls the step once extra at t=0
        if self.cell_count > 1:
            self.dec_timestep += 1
        return output, NestedState(K=K, V=V, R=R)
