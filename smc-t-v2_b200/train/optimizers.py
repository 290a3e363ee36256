"""Mirror of the optimizer wiring of src/algos/run_SMC_T.py:18-22: Adam(CustomSchedule(d_model), beta_1=0.9,
beta_2=0.98, epsilon=1e-9) and CustomSchedule (src/utils/utils_train.py:110-123).  The update itself is the
smct_adam_step kernel (C-ABI); this class only keeps the moment buffers and the step counter."""
import math

import torch

from .. import functional as F


class CustomSchedule:
    """lr(step) = d_model^-0.5 * min(step^-0.5, step * warmup^-1.5)  (utils_train.py:110-123)."""

    def __init__(self, d_model, warmup_steps=4000):
        self.d_model = float(d_model)
        self.warmup_steps = warmup_steps

    def __call__(self, step):
        step = float(step)
        arg1 = step ** -0.5 if step > 0 else float("inf")   # tf.math.rsqrt(0) = inf; min() picks arg2 = 0: lr(0) = 0
        return self.d_model ** -0.5 * min(arg1, step * self.warmup_steps ** -1.5)


class Adam:
    """tf.keras.optimizers.Adam semantics: m, v moments; lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.98, epsilon=1e-9):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self._slots = {}

    def _lr(self):
        lr = self.learning_rate
        return lr(self.iterations) if callable(lr) else lr

    def apply_gradients(self, grads_and_vars):
        """grads_and_vars: iterable of (grad tensor, parameter tensor) — parameters are updated in place."""
        # tf.keras: the schedule is evaluated at the PRE-increment `iterations`; iterations + 1 only enters the bias
        # correction (optimizer_v2 _prepare_local / adam.py) — so the first update of CustomSchedule has lr = 0
        lr = self._lr()
        self.iterations += 1
        t = self.iterations
        lr_t = lr * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
        for g, p in grads_and_vars:
            if g is None:
                continue
            key = p.data_ptr()
            if key not in self._slots:
                self._slots[key] = (torch.zeros_like(p), torch.zeros_like(p))
            m, v = self._slots[key]
            F.adam_step(p, g.reshape(p.shape).contiguous(), m, v, lr_t, self.beta_1, self.beta_2, self.epsilon)
