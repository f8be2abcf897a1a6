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


# ------------------------------------------------------------------------------------------------------------------
# Mirrors of the reference's loss classes (softmac/engine/losses/loss_{grip,pour,door,transport}.py): same names,
# constructor (cfg, mpm_sim), cfg.weight / cfg.target_path, initialize(), compute_loss(f) -> dict with the reference's
# keys, clear() / reset().  The Taichi versions accumulate into scalar fields under ti.ad.Tape and let autodiff write
# x.grad[f] and position/rotation/v/w.grad[f]; here compute_loss(f) evaluates the value AND hands the seeds to the
# simulator (add_x_grad / Primitive.add_all_states_grad).  The O(N^2) Chamfer part runs on the GPU; the rigid-body
# terms are a handful of scalars on the 13-vector of primitives[0] and stay on the host.
# ------------------------------------------------------------------------------------------------------------------
def _cfg_get(cfg, key, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class _RigidTerms:
    """pose / velocity terms shared by the reference losses; s13 = [x(3) q(4, w first) v(3) w(3)] of primitives[0] at frame f."""

    @staticmethod
    def velocity(s13, w_ang):                       # loss_grip.py:85-88 (w_ang = 0.1), loss_door.py:43-44 (w_ang = 0)
        v, w = s13[7:10], s13[10:13]
        g = np.zeros(13)
        g[7:10] = 2 * v
        g[10:13] = 2 * w_ang * w
        return float(v @ v + w_ang * (w @ w)), g


class _ChamferPoseVelLoss:
    """GripLoss / PourLoss: weight = (chamfer, pose, velocity)."""
    rotation_terms = True

    def __init__(self, cfg, mpm_sim):
        self.cfg, self.sim = cfg, mpm_sim
        self.dim, self.n_particles, self.dtype = mpm_sim.dim, mpm_sim.n_particles, mpm_sim.dtype
        self.rigid_control = mpm_sim.primitives[0] if len(mpm_sim.primitives) else None
        self.weight = (1.0, 0.0, 0.0)
        self.target = None
        self._chamfer = None
        self.loss = 0.0

    def load_target_position(self, path):
        self.set_target(np.load(path))

    def set_target(self, pos):
        self.target = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
        self._chamfer = DeviceChamferLoss(self.sim, self.target, weight=1.0)

    def initialize(self):
        w = _cfg_get(self.cfg, "weight", (1.0, 0.0, 0.0))
        self.weight = tuple(float(x) for x in w)
        path = _cfg_get(self.cfg, "target_path")
        if path is not None and self.target is None:
            self.load_target_position(path)

    def pose(self, s13):
        """10 (y - 0.4)^2 [+ min(0, |q_w| - 0.5)^2 + max(0, |q_w| - 0.9)^2 in GripLoss] (loss_grip.py:76-82, loss_pour.py:76-82)"""
        g = np.zeros(13)
        val = 10.0 * (s13[1] - 0.4) ** 2
        g[1] = 20.0 * (s13[1] - 0.4)
        if self.rotation_terms:
            a = abs(s13[3])
            lo, hi = min(0.0, a - 0.5), max(0.0, a - 0.9)
            val += lo * lo + hi * hi
            g[3] = (2 * lo + 2 * hi) * np.sign(s13[3])
        return float(val), g

    def compute_loss(self, f):
        cw, pw, vw = self.weight
        out = {"loss": 0.0, "chamfer_loss": 0.0, "pose_loss": 0.0, "vel_loss": 0.0}
        if cw > 0:
            self._chamfer.weight = cw
            out["chamfer_loss"] = self._chamfer.compute_loss(f)["loss"]
        if (pw > 0 or vw > 0) and self.rigid_control is not None:
            s13 = self.rigid_control.get_all_states(f)
            g = np.zeros(13)
            if pw > 0:
                val, gp = self.pose(s13)
                out["pose_loss"] = pw * val
                g += pw * gp
            if vw > 0:
                val, gv = _RigidTerms.velocity(s13, 0.1)
                out["vel_loss"] = vw * val
                g += vw * gv
            self.rigid_control.add_all_states_grad(f, g)
        # the reference returns the CUMULATIVE field under 'loss' (sum_up_loss_kernel does `self.loss[None] +=`, loss_grip.py:92-95,
        # and _extract_loss returns self.loss[None], :139-146; demo_grip.py:152 assigns, not adds); 'frame_loss' is this frame's share
        out["frame_loss"] = out["chamfer_loss"] + out["pose_loss"] + out["vel_loss"]
        self.loss += out["frame_loss"]
        out["loss"] = self.loss
        return out

    def clear(self):
        self.loss = 0.0

    reset = clear


class GripLoss(_ChamferPoseVelLoss):
    """softmac/engine/losses/loss_grip.py"""
    rotation_terms = True


class PourLoss(_ChamferPoseVelLoss):
    """softmac/engine/losses/loss_pour.py (the two rotation terms are commented out there, :80-81)"""
    rotation_terms = False


class _PoseVelContactLoss:
    """DoorLoss / TransportLoss: weight = (pose, velocity, contact).  The contact term is min_i max(|x_i - p|^2 - 0.01, 0) per
    controller group, squared (loss_door.py:46-56); its gradient goes to the minimising particle and to the primitive position
    (the subgradient Taichi's reverse mode of ti.atomic_min takes is not pinned by the reference [ext])."""
    n_groups = 1

    def __init__(self, cfg, mpm_sim):
        self.cfg, self.sim = cfg, mpm_sim
        self.dim, self.n_particles, self.dtype = mpm_sim.dim, mpm_sim.n_particles, mpm_sim.dtype
        self.n_particles_per_controller = self.n_particles // self.n_groups
        self.rigid = mpm_sim.primitives[0]
        self.device_contact = hasattr(mpm_sim, "_h")        # the CUDA simulator evaluates the contact term on the device
        self.weight = (1.0, 0.0, 0.0)
        self.target = []
        self.loss = 0.0

    def set_target(self, target):
        self.target = target

    def initialize(self):
        self.weight = tuple(float(x) for x in _cfg_get(self.cfg, "weight", (1.0, 0.0, 0.0)))

    def pose(self, s13):
        raise NotImplementedError

    def contact_device(self, f, weight):
        """The contact term on the GPU (smx_contact_distance_loss): no device->host copy of x; the seed goes straight into the
        loss-seed buffer of frame f and into the position adjoint of primitives[0].  Returns the weighted value (summed over
        batched rollouts)."""
        from .._capi import lib, check, d_ptr
        out = np.zeros(max(int(getattr(self.sim, "n_batch", 1)), 1))
        check(lib().smx_contact_distance_loss(self.sim._h, int(f), int(getattr(self.rigid, "_id", 0)), int(self.n_groups), float(weight), d_ptr(out)))
        return float(out.sum())

    def contact(self, x, pos):
        """numpy restatement of the contact term (the checker of contact_device in tests/test_losses.py)"""
        npc = self.n_particles_per_controller
        val, gx, gp = 0.0, np.zeros_like(x), np.zeros(3)
        for k in range(self.n_groups):
            d = x[k * npc:(k + 1) * npc] - pos
            dist = np.maximum((d * d).sum(1) - 0.01, 0.0)
            i = int(dist.argmin())
            m = min(float(dist[i]), 1e6)
            val += m * m
            if 0.0 < m < 1e6:
                gx[k * npc + i] += 2 * m * 2 * d[i]
                gp -= 2 * m * 2 * d[i]
        return val, gx, gp

    def compute_loss(self, f):
        pw, vw, cw = self.weight
        out = {"loss": 0.0, "pose_loss": 0.0, "vel_loss": 0.0, "contact_loss": 0.0}
        s13 = self.rigid.get_all_states(f)
        g = np.zeros(13)
        if pw > 0:
            val, gp = self.pose(s13)
            out["pose_loss"] = pw * val
            g += pw * gp
        if vw > 0:
            val, gv = _RigidTerms.velocity(s13, 0.0)
            out["vel_loss"] = vw * val
            g += vw * gv
        if cw > 0:
            if self.device_contact:
                out["contact_loss"] = self.contact_device(f, cw)
            else:
                val, gx, gpos = self.contact(self.sim.get_x(f), s13[:3])
                out["contact_loss"] = cw * val
                g[:3] += cw * gpos
                if np.any(gx):
                    self.sim.add_x_grad(f, cw * gx)
        if np.any(g):
            self.rigid.add_all_states_grad(f, g)
        out["frame_loss"] = out["pose_loss"] + out["vel_loss"] + out["contact_loss"]
        self.loss += out["frame_loss"]
        out["loss"] = self.loss          # cumulative, as the reference's loss field (loss_door.py:71-75, 110-117)
        return out

    def clear(self):
        self.loss = 0.0

    reset = clear


class DoorLoss(_PoseVelContactLoss):
    """softmac/engine/losses/loss_door.py: pose = (q_w - cos(pi/8))^2"""
    n_groups = 1

    def pose(self, s13):
        g = np.zeros(13)
        d = s13[3] - np.cos(np.pi / 8)
        g[3] = 2 * d
        return float(d * d), g


class TransportLoss(_PoseVelContactLoss):
    """softmac/engine/losses/loss_transport.py: pose = |position - target|^2, two controller groups"""
    n_groups = 2

    def pose(self, s13):
        g = np.zeros(13)
        d = s13[:3] - np.asarray(self.target, dtype=np.float64)
        g[:3] = 2 * d
        return float(d @ d), g
