"""A small TensorFlow-API shim over torch-CPU, just large enough to EXECUTE the reference's own,
unmodified Python sources of the hot path (src/models/SMC_Transformer/{SMC_Transformer,
SMC_TransformerCell,self_attention_SMC,transformer_utils}.py, src/models/classic_layers.py).

*** TEST INFRASTRUCTURE — NOT PRODUCT CODE. ***

Why: TensorFlow / tensorflow-probability are not installed (no network), so the reference cannot run
as-is.  With this shim its control flow, layer wiring and quirks (LN1 applied twice, the K.rnn dummy
first call and the cell_count/dec_timestep compensation, noise_z = z - z_, un-normalised X in
noise_K_resampled, weights passed as logits, ...) are executed verbatim from /root/reference; only the
*semantics of the TF ops themselves* are restated here (on torch CPU fp32), which is the part that
stays "parity unpinned at the TensorFlow boundary".

Random draws are injected: `install(...).rng.normal_fn / uniform_fn` are called in call order.
Usage: see oracle/gen_golden_from_reference.py.
"""
import collections
import math
import sys
import types

import numpy as np
import torch

T = torch.Tensor


def _t(x, dtype=None):
    if isinstance(x, T):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, Variable):
        return x.value if dtype is None else x.value.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


TRACK_GRADIENTS = False   # set by the gradient-fixture generator: trainable variables become autograd leaves


class Variable:
    def __init__(self, initial_value=None, name=None, shape=None, trainable=True, dtype=None):
        self.value = _t(initial_value).clone().to(torch.float32) if not isinstance(initial_value, T) else initial_value.clone()
        self.name = name
        self.trainable = trainable
        self._trainable = trainable
        if TRACK_GRADIENTS and trainable and self.value.is_floating_point():
            self.value = self.value.detach().requires_grad_(True)

    def numpy(self):
        return self.value.detach().numpy()

    @property
    def shape(self):
        return self.value.shape

    def assign(self, v):
        self.value = _t(v).detach().clone()
        if TRACK_GRADIENTS and self.trainable and self.value.is_floating_point():
            self.value.requires_grad_(True)

    # arithmetic used on logvar variables
    def __mul__(self, o): return self.value * _t(o)
    __rmul__ = __mul__
    def __add__(self, o): return self.value + _t(o)
    __radd__ = __add__
    def __sub__(self, o): return self.value - _t(o)
    def __rsub__(self, o): return _t(o) - self.value
    def __neg__(self): return -self.value

    @classmethod
    def __torch_function__(cls, func, types_, args=(), kwargs=None):
        kwargs = kwargs or {}
        args = tuple(a.value if isinstance(a, Variable) else a for a in args)
        return func(*args, **kwargs)


class RNG:
    """Injected randomness, consumed in call order (the reference draws from TF's global generator)."""

    def __init__(self):
        self.normal_fn = None      # (shape) -> torch tensor
        self.uniform_fn = None     # (B, num_samples) -> torch float64 tensor in [0,1)
        self.np_rng = np.random.default_rng(0)


def _build_tf(rng: RNG):
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float64, tf.int32, tf.int64 = torch.float32, torch.float64, torch.int32, torch.int64
    tf.newaxis = None
    tf.Variable = Variable
    tf.Tensor = T

    class TensorShape(tuple):
        def __new__(cls, dims):
            return super().__new__(cls, tuple(dims))
    tf.TensorShape = TensorShape

    def _shape_arg(shape):
        return tuple(int(s) for s in (shape if isinstance(shape, (tuple, list, torch.Size)) else [shape]))

    tf.shape = lambda x: _t(x).shape
    tf.constant = lambda v, dtype=None, shape=None: (_t(v, dtype).reshape(_shape_arg(shape)) if shape is not None else _t(v, dtype))
    tf.zeros = lambda shape, dtype=torch.float32: torch.zeros(_shape_arg(shape), dtype=dtype)
    tf.ones = lambda shape, dtype=torch.float32: torch.ones(_shape_arg(shape), dtype=dtype)
    tf.cast = lambda x, dtype: _t(x).to(dtype) if not isinstance(x, (int, float)) else torch.tensor(x, dtype=dtype)
    tf.concat = lambda values, axis: torch.cat([_t(v) for v in values], dim=axis)
    tf.stack = lambda values, axis=0: torch.stack([_t(v) for v in values], dim=axis)
    tf.transpose = lambda x, perm=None: _t(x).permute(*perm) if perm is not None else _t(x).T
    tf.tile = lambda x, multiples: _t(x).repeat(*[int(m) for m in multiples])
    tf.stop_gradient = lambda x: _t(x).detach()

    class GradientTape:
        """tf.GradientTape over torch autograd: the shim's ops ARE torch ops, so the tape is the autograd graph of the
        forward that ran inside the `with` block (trainable variables are leaves when TRACK_GRADIENTS is set)."""

        def __init__(self, persistent=False, **kw):
            self.persistent = persistent

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def gradient(self, target, sources):
            vals = [s.value if isinstance(s, Variable) else s for s in sources]
            return list(torch.autograd.grad(_t(target), vals, allow_unused=True, retain_graph=True))
    tf.GradientTape = GradientTape
    tf.exp = lambda x: torch.exp(_t(x))
    tf.square = lambda x: torch.square(_t(x))
    tf.scalar_mul = lambda s, x: s * _t(x)
    tf.einsum = lambda eq, *ops: torch.einsum(eq, *[_t(o) for o in ops])
    tf.tensordot = lambda a, b, axes: torch.tensordot(_t(a), _t(b), dims=([axes[0] % _t(a).dim()], [axes[1]]))

    def reshape(x, shape):
        return _t(x).reshape(tuple(int(s) for s in shape))
    tf.reshape = reshape

    def squeeze(x, axis=None):
        x = _t(x)
        if axis is None:
            return x.squeeze()
        axes = [axis] if isinstance(axis, int) else list(axis)
        for a in sorted([a % x.dim() for a in axes], reverse=True):
            assert x.shape[a] == 1, "tf.squeeze on a non-unit axis"
            x = x.squeeze(a)
        return x
    tf.squeeze = squeeze
    tf.expand_dims = lambda x, axis: _t(x).unsqueeze(axis)

    def split(value, num_or_size_splits, axis=0):
        v = _t(value)
        if isinstance(num_or_size_splits, int):
            return list(torch.chunk(v, num_or_size_splits, dim=axis))
        return list(torch.split(v, list(num_or_size_splits), dim=axis))
    tf.split = split

    def matmul(a, b, transpose_a=False, transpose_b=False):
        a, b = _t(a), _t(b)
        if transpose_a:
            a = a.transpose(-1, -2)
        if transpose_b:
            b = b.transpose(-1, -2)
        return torch.matmul(a, b)
    tf.matmul = matmul

    def reduce_sum(x, axis=None, keepdims=False):
        x = _t(x)
        return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)
    tf.reduce_sum = reduce_sum

    def reduce_mean(x, axis=None, keepdims=False):
        x = _t(x)
        return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)
    tf.reduce_mean = reduce_mean

    def gather(params, indices, axis=0, batch_dims=0):
        """tf.gather for the two call sites of the hot path: (axis=1, batch_dims=1) and (axis=-1, batch_dims=2)."""
        params, indices = _t(params), _t(indices).long()
        axis = axis % params.dim()
        if batch_dims == 1 and axis == 1:       # transformer_utils.py:47
            idx = indices.reshape(indices.shape + (1,) * (params.dim() - 2)).expand(indices.shape + params.shape[2:])
            return torch.gather(params, 1, idx)
        if batch_dims == 2 and axis == params.dim() - 1:   # SMC_TransformerCell.py:81
            return torch.gather(params, -1, indices.unsqueeze(-1)).squeeze(-1)
        raise NotImplementedError("tf.gather(axis=%d, batch_dims=%d)" % (axis, batch_dims))
    tf.gather = gather

    # ---- tf.math / tf.nn / tf.linalg / tf.nest / tf.random ---------------------------------
    tf.math = types.ModuleType("tensorflow.math")
    tf.math.log = lambda x: torch.log(_t(x, torch.float32) if not isinstance(x, T) else x)
    tf.math.exp = tf.exp
    tf.math.sqrt = lambda x: torch.sqrt(_t(x))
    tf.math.multiply = lambda a, b: _t(a) * _t(b)
    tf.math.is_nan = lambda x: torch.isnan(_t(x))
    tf.math.equal = lambda a, b: _t(a) == _t(b)
    tf.math.logical_not = lambda x: ~_t(x)
    tf.math.softmax = lambda x, axis=-1: torch.softmax(_t(x), dim=axis)
    tf.math.argmax = lambda x, axis=-1: torch.argmax(_t(x), dim=axis)
    tf.nn = types.ModuleType("tensorflow.nn")
    tf.nn.softmax = lambda x, axis=-1: torch.softmax(_t(x), dim=axis)
    tf.nn.relu = lambda x: torch.relu(_t(x))
    tf.nn.gelu = lambda x: torch.nn.functional.gelu(_t(x))
    tf.linalg = types.ModuleType("tensorflow.linalg")
    tf.linalg.diag = lambda x: torch.diag(_t(x))
    tf.linalg.diag_part = lambda x: torch.diagonal(_t(x))
    tf.linalg.band_part = lambda x, lo, hi: torch.tril(_t(x)) if (lo == -1 and hi == 0) else (_ for _ in ()).throw(NotImplementedError())
    tf.nest = types.ModuleType("tensorflow.nest")

    def flatten(struct):
        if isinstance(struct, (T, Variable)) or not isinstance(struct, (list, tuple)):
            return [struct]
        out = []
        for s in struct:
            out += flatten(s)
        return out
    tf.nest.flatten = flatten
    tf.random = types.ModuleType("tensorflow.random")

    def normal(shape, dtype=torch.float32, **kw):
        shp = _shape_arg(shape)
        if rng.normal_fn is not None:
            return rng.normal_fn(shp).to(dtype)
        return torch.from_numpy(rng.np_rng.standard_normal(shp).astype(np.float32)).to(dtype)
    tf.random.normal = normal
    tf.random.uniform = lambda shape, dtype=torch.float32, **kw: torch.from_numpy(rng.np_rng.random(_shape_arg(shape)).astype(np.float32)).to(dtype)

    def categorical(logits, num_samples, **kw):
        """TensorFlow CPU kernel (multinomial_op.cc): max over finite logits, fp64 running sum of
        exp(double(logit) - max), target = U * total, upper_bound.  U is injected."""
        lg = _t(logits).to(torch.float32)
        B, C = lg.shape
        u = rng.uniform_fn(B, num_samples) if rng.uniform_fn is not None else torch.from_numpy(rng.np_rng.random((B, num_samples)))
        l64 = lg.double()
        mx = torch.where(torch.isfinite(l64), l64, torch.full_like(l64, -math.inf)).max(dim=1, keepdim=True).values
        e = torch.where(torch.isfinite(l64), torch.exp(l64 - mx), torch.zeros_like(l64))
        cdf = torch.cumsum(e, dim=1)
        idx = torch.searchsorted(cdf, u.double() * cdf[:, -1:], right=True)
        return idx.clamp_(max=C - 1)
    tf.random.categorical = categorical

    # ---- tf.keras ----------------------------------------------------------------------------
    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    backend = types.ModuleType("tensorflow.keras.backend")
    losses = types.ModuleType("tensorflow.keras.losses")
    utils_m = types.ModuleType("tensorflow.keras.utils")
    utils_m.plot_model = lambda *a, **k: None

    class Layer:
        def __init__(self, name=None, **kwargs):
            self.name = name
            self._built = False
            self._weights = []

        def add_weight(self, name=None, shape=None, initializer="glorot_uniform", trainable=True):
            shape = tuple(int(s) for s in shape)
            if initializer == "zeros":
                v = np.zeros(shape, np.float32)
            elif initializer == "ones":
                v = np.ones(shape, np.float32)
            else:
                lim = math.sqrt(6.0 / (shape[0] + shape[-1]))
                v = rng.np_rng.uniform(-lim, lim, size=shape).astype(np.float32)
            var = Variable(torch.from_numpy(v), name=name, trainable=trainable)
            self._weights.append(var)
            return var

        @property
        def weights(self):
            return self._weights

        def set_weights(self, ws):
            for var, w in zip(self._weights, ws):
                var.assign(torch.from_numpy(np.asarray(w, np.float32)))

        def build(self, input_shape):
            pass

        def __call__(self, *args, **kwargs):
            if not self._built and hasattr(self, "build") and args and isinstance(args[0], T):
                self.build(args[0].shape)
                self._built = True
            return self.call(*args, **kwargs)
    layers.Layer = Layer

    class Dense(Layer):
        def __init__(self, units, activation=None, use_bias=True, name=None):
            super().__init__(name=name)
            self.units, self.activation, self.use_bias = units, activation, use_bias
            self.kernel = self.bias = None

        def build(self, input_shape):
            self.kernel = self.add_weight("kernel", (input_shape[-1], self.units))
            if self.use_bias:
                self.bias = self.add_weight("bias", (self.units,), initializer="zeros")

        def call(self, x):
            y = torch.matmul(_t(x).to(torch.float32), self.kernel.value)
            if self.use_bias:
                y = y + self.bias.value
            if self.activation == "relu":
                y = torch.relu(y)
            elif self.activation is not None:
                raise NotImplementedError(self.activation)
            return y
    layers.Dense = Dense

    class LayerNormalization(Layer):
        """Keras non-fused path (epsilon < 1.001e-5): moments + batch_normalization."""

        def __init__(self, epsilon=1e-3, name=None):
            super().__init__(name=name)
            self.epsilon = epsilon

        def build(self, input_shape):
            self.gamma = self.add_weight("gamma", (input_shape[-1],), initializer="ones")
            self.beta = self.add_weight("beta", (input_shape[-1],), initializer="zeros")

        def call(self, x):
            x = _t(x)
            mean = x.mean(-1, keepdim=True)
            var = (x - mean).square().mean(-1, keepdim=True)
            inv = torch.rsqrt(var + self.epsilon) * _t(self.gamma)
            return x * inv + (_t(self.beta) - mean * inv)
    layers.LayerNormalization = LayerNormalization

    class Embedding(Layer):
        def __init__(self, input_dim, output_dim, name=None):
            super().__init__(name=name)
            self.input_dim, self.output_dim = input_dim, output_dim
            v = rng.np_rng.uniform(-0.05, 0.05, size=(input_dim, output_dim)).astype(np.float32)
            self.embeddings = Variable(torch.from_numpy(v), name="embeddings")
            self._weights.append(self.embeddings)
            self._built = True

        def call(self, x):
            return self.embeddings.value[_t(x).long()]
    layers.Embedding = Embedding

    class Dropout(Layer):
        def __init__(self, rate=0., **kw):
            super().__init__()
            self.rate = rate

        def call(self, x, training=False):
            assert not training or self.rate == 0., "dropout is never active on the reference's hot path"
            return x
    layers.Dropout = Dropout

    class Sequential(Layer):
        def __init__(self, layers_list=None, name=None):
            super().__init__(name=name)
            self.layers = list(layers_list or [])
            self._built = True

        def call(self, x):
            for l in self.layers:
                x = l(x)
            return x

        @property
        def weights(self):
            out = []
            for l in self.layers:
                out += l.weights
            return out
    keras.Sequential = Sequential

    class Model(Layer):
        def __init__(self, *a, **k):
            super().__init__()
            self._built = True
    keras.Model = Model

    def rnn(step_function, inputs, initial_states, **kw):
        """tf.keras.backend.rnn, eager non-unrolled path (TF 2.6 backend.py): inputs (B,S,...) nested;
        the step function is first called ONCE on the t=0 slice to infer the output structure (result
        discarded), then looped over t = 0..S-1.  Returns (last_output, outputs stacked on axis 1, states)."""
        flat = tf.nest.flatten(inputs)
        S = flat[0].shape[1]

        def slice_t(t):
            sl = [f[:, t] for f in flat]
            return type(inputs)(*sl) if hasattr(inputs, "_fields") else sl
        step_function(slice_t(0), tuple(initial_states))   # dummy call (output discarded)
        states = tuple(initial_states)
        outs = []
        for t in range(S):
            out, states = step_function(slice_t(t), tuple(states))
            outs.append(out)

        def stack(items):
            if isinstance(items[0], (list, tuple)):
                return [stack([it[i] for it in items]) for i in range(len(items[0]))]
            return torch.stack(items, dim=1)
        return outs[-1], stack(outs), list(states)
    backend.rnn = rnn

    class SparseCategoricalCrossentropy:
        def __init__(self, from_logits=False, reduction="none"):
            self.from_logits = from_logits

        def __call__(self, y_true, y_pred):
            y_pred = _t(y_pred)
            yt = _t(y_true).long().reshape(y_pred.shape[:-1])
            logp = torch.log_softmax(y_pred, dim=-1) if self.from_logits else torch.log(y_pred.clamp_min(1e-7))
            return -torch.gather(logp, -1, yt.unsqueeze(-1)).squeeze(-1)
    losses.SparseCategoricalCrossentropy = SparseCategoricalCrossentropy
    losses.MSE = lambda a, b: (_t(a) - _t(b)).square().mean(-1)

    keras.layers, keras.backend, keras.losses, keras.utils = layers, backend, losses, utils_m
    tf.keras = keras
    tf.losses = types.ModuleType("tensorflow.losses")

    # ---- tensorflow_probability (only MultivariateNormalDiag.log_prob, SMC_Transformer.py:97-98) ----
    tfp = types.ModuleType("tensorflow_probability")
    tfp.distributions = types.ModuleType("tensorflow_probability.distributions")

    class MultivariateNormalDiag:
        def __init__(self, scale_diag=None, loc=None):
            self.scale = _t(scale_diag)

        def log_prob(self, x):
            x = _t(x)
            D = x.shape[-1]
            z = x / self.scale
            return -0.5 * (z * z).sum(-1) - torch.log(self.scale).sum() * (1.0 if self.scale.numel() == D else D) \
                - 0.5 * D * math.log(2 * math.pi)
    tfp.distributions.MultivariateNormalDiag = MultivariateNormalDiag
    mods = {"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
            "tensorflow.keras.backend": backend, "tensorflow.keras.losses": losses, "tensorflow.keras.utils": utils_m,
            "tensorflow.math": tf.math, "tensorflow.nn": tf.nn, "tensorflow.linalg": tf.linalg,
            "tensorflow.random": tf.random, "tensorflow_probability": tfp}
    return tf, mods


Installed = collections.namedtuple("Installed", ["tf", "rng", "modules"])


def install(reference_root="/root/reference"):
    """Put the shim (and stubs for the reference modules that load GPT-2 / nltk at import time) into
    sys.modules and make `src.*` importable from the read-only reference checkout."""
    rng = RNG()
    tf, mods = _build_tf(rng)
    for k, v in mods.items():
        sys.modules[k] = v
    # src/eval/language_metrics.py loads cache/gpt2 at import (language_metrics.py:6-10): stub it.
    lm = types.ModuleType("src.eval.language_metrics")
    lm.gpt2_perplexity_batch = lm.gpt2_perplexity_batch_2 = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    # multi-layer decoder / GPT-2 decoder: never instantiated with num_layers == 1
    twe = types.ModuleType("src.models.Baselines.Transformer_without_enc")
    twe.Decoder = type("Decoder", (), {})
    g2 = types.ModuleType("src.models.Baselines.GPT2Decoder")
    g2.GPT2Decoder = type("GPT2Decoder", (), {})
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import importlib
    for name in [m for m in sys.modules if m == "src" or m.startswith("src.")]:
        del sys.modules[name]
    importlib.import_module("src")
    importlib.import_module("src.eval")
    sys.modules["src.eval.language_metrics"] = lm
    importlib.import_module("src.models")
    importlib.import_module("src.models.Baselines")
    sys.modules["src.models.Baselines.Transformer_without_enc"] = twe
    sys.modules["src.models.Baselines.GPT2Decoder"] = g2
    return Installed(tf, rng, mods)
