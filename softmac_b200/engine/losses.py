"""Loss seeds for the adjoint (SURVEY.md 8f row 2 -- next to the hot path, not on it).

The Taichi losses of the reference write ``x.grad[f]`` directly (softmac/engine/losses/loss_grip.py:117-140); here a
loss computes its value and hands the seed to the simulator through ``add_x_grad``.  ``ChamferLoss`` follows
``GripLoss`` with weight (1, 0, 0): sum over current particles of the squared distance to the closest target plus
the same with the roles swapped (loss_grip.py:45-68), nearest neighbours held fixed in the gradient.
"""
import numpy as np


class PointwiseLoss:
    """0.5 * weight * |x[f] - target|^2 summed over particles (used by the coupling tests)."""

    def __init__(self, simulator, target, weight=1.0):
        self.sim, self.target, self.weight = simulator, np.asarray(target, dtype=np.float64), weight

    def initialize(self):
        pass

    def reset(self):
        pass

    def compute_loss(self, f):
        d = self.sim.get_x(f) - self.target
        self.sim.add_x_grad(f, self.weight * d)
        return {"loss": 0.5 * self.weight * float((d * d).sum())}


class ChamferLoss:
    def __init__(self, simulator, target, weight=1.0, chunk=2048):
        self.sim, self.target, self.weight, self.chunk = simulator, np.asarray(target, dtype=np.float64), weight, chunk

    def initialize(self):
        pass

    def reset(self):
        pass

    def _nearest(self, a, b):
        idx = np.empty(len(a), dtype=np.int64)
        for i in range(0, len(a), self.chunk):
            d = ((a[i:i + self.chunk, None, :] - b[None, :, :]) ** 2).sum(-1)
            idx[i:i + self.chunk] = d.argmin(1)
        return idx

    def compute_loss(self, f):
        x, t = self.sim.get_x(f), self.target
        i_cur, i_tar = self._nearest(x, t), self._nearest(t, x)
        d1, d2 = x - t[i_cur], x[i_tar] - t
        loss = float((d1 * d1).sum() + (d2 * d2).sum())
        g = 2 * d1
        np.add.at(g, i_tar, 2 * d2)
        self.sim.add_x_grad(f, self.weight * g)
        return {"loss": self.weight * loss, "chamfer_loss": self.weight * loss}


class DeviceChamferLoss:
    """ChamferLoss evaluated on the GPU (smx_chamfer_loss): no device->host copy of x, no O(N^2) numpy; the seed goes straight
    into the simulator's loss-seed buffer of frame f.  With a batched handle the loss is summed over the rollouts."""

    def __init__(self, simulator, target, weight=1.0):
        import ctypes as C
        from .._capi import lib, check, as_d, d_ptr
        self.sim, self.weight = simulator, float(weight)
        t = as_d(np.asarray(target, dtype=np.float64)).reshape(-1, 3)
        check(lib().smx_set_chamfer_target(simulator._h, d_ptr(t), len(t)))
        self._C, self._lib, self._check = C, lib, check

    def initialize(self):
        pass

    def reset(self):
        pass

    def compute_loss(self, f):
        out = self._C.c_double()
        self._check(self._lib().smx_chamfer_loss(self.sim._h, int(f), self.weight, self._C.byref(out)))
        return {"loss": out.value, "chamfer_loss": out.value}
