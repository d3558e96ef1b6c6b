"""Shared host-side plumbing of the MF-family models: initialisers (util/tool.py:79-97 of the
reference), TensorFlow-1.12 optimizer bookkeeping (util/learner.py:2-15) and device buffers.
No arithmetic of the hot path happens here: tables, gradients and optimizer slots are torch
tensors used purely as device memory for the sm_100a kernels."""
import numpy as np
import torch

# TF-1.12 constructor defaults behind learner.optimizer (learner.py:4-14)
OPT_HYPER = {
    "gd": lambda lr, mom: [lr],
    "adam": lambda lr, mom: [lr, 0.9, 0.999, 1e-8],
    "adagrad": lambda lr, mom: [lr],
    "rmsprop": lambda lr, mom: [lr, 0.9, 0.0, 1e-10],
    "momentum": lambda lr, mom: [lr, mom],
}
# slot initial values: adam m, v = 0; adagrad accumulator = 1e-8 (learner.py:5-6);
# rmsprop rms = 1, momentum = 0; momentum accumulator = 0
OPT_SLOTS = {"gd": (None, None), "adam": (0.0, 0.0), "adagrad": (1e-8, None),
             "rmsprop": (1.0, 0.0), "momentum": (0.0, None)}


def get_initializer(init_method, stddev, generator):
    """tool.get_initializer (util/tool.py:79-97).  TF's Philox init streams are not
    reproducible without TF, so the draws come from a seeded torch generator instead
    (main.py:12 seeds tf with 2017; we seed the generator with 2017)."""
    def normal(shape, std):
        return torch.randn(shape, generator=generator) * std

    def tnormal(shape, std):  # truncated_normal: redraw beyond 2 sigma
        x = torch.randn(shape, generator=generator)
        bad = x.abs() > 2
        while bad.any():
            x[bad] = torch.randn(int(bad.sum()), generator=generator)
            bad = x.abs() > 2
        return x * std

    def uniform(shape, lim):
        return (torch.rand(shape, generator=generator) * 2 - 1) * lim

    def fans(shape):
        return (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])

    table = {
        "tnormal": lambda s: tnormal(s, stddev),
        "uniform": lambda s: uniform(s, stddev),
        "normal": lambda s: normal(s, stddev),
        "xavier_normal": lambda s: normal(s, (2.0 / sum(fans(s))) ** 0.5),
        "xavier_uniform": lambda s: uniform(s, (6.0 / sum(fans(s))) ** 0.5),
        "he_normal": lambda s: tnormal(s, (2.0 / fans(s)[0]) ** 0.5 / 0.87962566103423978),
        "he_uniform": lambda s: uniform(s, (6.0 / fans(s)[0]) ** 0.5),
    }
    fn = table.get(init_method, table["tnormal"])           # tool.py:96-97 default
    return lambda shape: fn(tuple(shape)).to(torch.float32)


class OptimizerState(object):
    """learner.optimizer(...) bookkeeping: which update rule, its hyper-parameters and, for
    Adam, the fp32 beta-power variables TF multiplies by beta after every step."""

    def __init__(self, learner, learning_rate, momentum=0.9):
        self.kind = learner.lower()
        if self.kind not in OPT_HYPER:
            raise ValueError("please select a suitable optimizer")        # learner.py:14-15
        self.lr = float(learning_rate)
        self.hyper = OPT_HYPER[self.kind](self.lr, momentum)
        self._p1 = np.float32(0.9)
        self._p2 = np.float32(0.999)
        self.stamp = 1
        self._pows_dev = None

    def lr_t(self, steps):
        """fp32 lr*sqrt(1-b2^t)/(1-b1^t) for the next `steps` steps (adam.py::_prepare/_finish)."""
        out = np.full(max(steps, 1), self.lr, dtype=np.float32)
        if self.kind == "adam":
            one, lr = np.float32(1.0), np.float32(self.lr)
            for s in range(steps):
                out[s] = lr * np.sqrt(one - self._p2) / (one - self._p1)
                self._p1 = np.float32(self._p1 * np.float32(0.9))
                self._p2 = np.float32(self._p2 * np.float32(0.999))
        return out

    def device_pows(self):
        """Device copy of TF's beta1_power / beta2_power variables (fp32 [2]); the persistent epoch
        kernels read and advance it themselves, `lr_t(steps)` advances the host mirror."""
        if self.kind != "adam":
            return None
        if self._pows_dev is None:
            self._pows_dev = torch.tensor([float(self._p1), float(self._p2)], dtype=torch.float32, device="cuda")
        return self._pows_dev

    def take_stamps(self, steps):
        first = self.stamp
        self.stamp += steps
        return first

    def slots_like(self, var):
        i0, i1 = OPT_SLOTS[self.kind]
        mk = lambda v: None if v is None else torch.full_like(var, v)
        return mk(i0), mk(i1)
