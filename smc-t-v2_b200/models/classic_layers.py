"""Parameter holders mirroring the Keras layers the hot path uses (src/models/classic_layers.py:19-23,
tf.keras.layers.Dense / LayerNormalization / Embedding).  They own torch CUDA tensors (containers
only); all arithmetic happens in libsmct_b200.so."""
import math

import numpy as np
import torch


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("smct_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Variable:
    """A named trainable tensor (tf.Variable stand-in): .value is a contiguous fp32 CUDA tensor."""

    def __init__(self, value, name, trainable=True):
        self.value = torch.as_tensor(np.asarray(value, np.float32), device=_device()).contiguous()
        self.name = name
        self.trainable = trainable

    def numpy(self):
        return self.value.detach().cpu().numpy()

    def assign(self, v):
        self.value.copy_(torch.as_tensor(np.asarray(v, np.float32)).reshape(self.value.shape))

    @property
    def shape(self):
        return tuple(self.value.shape)


class Dense:
    """tf.keras.layers.Dense(units, use_bias): kernel (in, units) Glorot-uniform, bias zeros; built lazily."""

    def __init__(self, units, use_bias=True, name="dense", activation=None):
        self.units, self.use_bias, self.name, self.activation = units, use_bias, name, activation
        self.kernel = None
        self.bias = None

    @property
    def built(self):
        return self.kernel is not None

    def build(self, in_dim, rng):
        lim = math.sqrt(6.0 / (in_dim + self.units))
        self.kernel = Variable(rng.uniform(-lim, lim, size=(in_dim, self.units)), self.name + "/kernel:0")
        self.bias = Variable(np.zeros(self.units), self.name + "/bias:0") if self.use_bias else None

    @property
    def trainable_variables(self):
        return [v for v in (self.kernel, self.bias) if v is not None]


class LayerNormalization:
    """tf.keras.layers.LayerNormalization(epsilon): gamma ones, beta zeros."""

    def __init__(self, epsilon=1e-6, name="layer_norm"):
        self.epsilon, self.name = epsilon, name
        self.gamma = None
        self.beta = None

    def build(self, dim):
        self.gamma = Variable(np.ones(dim), self.name + "/gamma:0")
        self.beta = Variable(np.zeros(dim), self.name + "/beta:0")

    @property
    def trainable_variables(self):
        return [v for v in (self.gamma, self.beta) if v is not None]


class Embedding:
    """tf.keras.layers.Embedding(input_dim, output_dim): table U(-0.05, 0.05)."""

    def __init__(self, input_dim, output_dim, name="embedding"):
        self.input_dim, self.output_dim, self.name = input_dim, output_dim, name
        self.embeddings = None

    def build(self, rng):
        self.embeddings = Variable(rng.uniform(-0.05, 0.05, size=(self.input_dim, self.output_dim)), self.name + "/embeddings:0")

    @property
    def trainable_variables(self):
        return [self.embeddings] if self.embeddings is not None else []


class FFN:
    """point_wise_feed_forward_network(d_model, dff) (classic_layers.py:19-23): Dense(dff, relu) -> Dense(d_model)."""

    def __init__(self, d_model, dff):
        self.dense1 = Dense(dff, name="FFN1_after_mha_dff", activation="relu")
        self.dense2 = Dense(d_model, name="FFN2_after_mha_dmodel")
        self.layers = [self.dense1, self.dense2]

    def build(self, d_model, rng):
        self.dense1.build(d_model, rng)
        self.dense2.build(self.dense1.units, rng)

    @property
    def trainable_variables(self):
        return self.dense1.trainable_variables + self.dense2.trainable_variables


def point_wise_feed_forward_network(d_model, dff):
    return FFN(d_model, dff)
