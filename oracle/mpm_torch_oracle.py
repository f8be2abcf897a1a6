"""Second, independent restatement of the hot path: PyTorch float64, adjoints from torch.autograd.

TEST INFRASTRUCTURE ONLY (tests/ import it; the product never does).  Parity to Taichi itself remains unpinned (Taichi 1.4.1 cannot
be installed in this image, the reference ships no golden vectors): this file exists so that the C oracle (oracle/mpm_oracle.c, forward
AND hand-written adjoint) is checked against a transliteration written separately, statement by statement, from the reference sources,
whose gradients come from a general-purpose autodiff instead of hand derivation.

Every function cites the reference lines it follows (paths relative to /root/reference/softmac/engine):
  compute_F_tmp            mpm_simulator.py:125-128        svd / backward_svd / clamp   :130-157, :184-192
  p2g                      :198-262                        boundary_condition           :268-281
  grid_op                  :283-297                        g2p                          :299-318
  grid_op_mixed1..4        :396-443                        substep order                :320-337
  Primitive.sdf / normal / collider_v / collide / collide_particle / collide_mixed      primitive/primitive_base.py:53-181
  forward_kinematics       primitive/primitive_base.py:280-283
  Mesh._sdf / _normal      primitive/mesh.py:45-108        length, qrot, qmul, w2quat, inv_trans   primitive/primitive_utils.py:4-46

Taichi autodiff conventions reproduced explicitly:
  * cast(int) and comparisons carry no gradient (``.detach()`` on every integer base / mask);
  * ti.max(a, b) passes its gradient to a iff b < a (else to b), ti.min(a, b) to a iff a < b (else to b): ``ti_max`` / ``ti_min``;
  * the SVD adjoint is the reference's own explicit formula (backward_svd with its +-1e-6 clamp), not torch's;
  * ``life`` = 1 / (substeps - f % substeps) is evaluated in float32 (Taichi default_fp);
  * atomic ``+=`` scatters are index_add (gradient passes through unchanged).
Loops over particles are vectorised (one tensor op per reference statement); branches are evaluated with torch.where on
branch-safe operands, so the taken branch is exactly the reference's.
"""
import numpy as np
import torch

DT = torch.float64
INF = 1e10                                      # mesh.py:12


def ti_max(a, b):                               # gradient to a iff b < a, else to b
    a, b = torch.broadcast_tensors(torch.as_tensor(a, dtype=DT), torch.as_tensor(b, dtype=DT))
    return torch.where(b < a, a, b)


def ti_min(a, b):                               # gradient to a iff a < b, else to b
    a, b = torch.broadcast_tensors(torch.as_tensor(a, dtype=DT), torch.as_tensor(b, dtype=DT))
    return torch.where(a < b, a, b)


def dot(a, b):
    return (a * b).sum(-1)


def length(x):                                  # primitive_utils.py:4-5
    return torch.sqrt(dot(x, x) + 1e-8)


def qrot(rot, v):                               # primitive_utils.py:8-13
    qvec = rot[..., 1:4].expand(v.shape)
    uv = torch.linalg.cross(qvec, v)
    uuv = torch.linalg.cross(qvec, uv)
    return v + 2 * (rot[..., 0:1] * uv + uuv)


def qmul(q, r):                                 # primitive_utils.py:20-27 (terms = r (x) q)
    t = r[:, None] * q[None, :]
    w = t[0, 0] - t[1, 1] - t[2, 2] - t[3, 3]
    x = t[0, 1] + t[1, 0] - t[2, 3] + t[3, 2]
    y = t[0, 2] + t[1, 3] + t[2, 0] - t[3, 1]
    z = t[0, 3] - t[1, 2] + t[2, 1] + t[3, 0]
    out = torch.stack([w, x, y, z])
    return out / torch.sqrt(dot(out, out))


def w2quat(axis_angle):                         # primitive_utils.py:29-40 (norm(1e-12) = sqrt(dot + 1e-12))
    w = torch.sqrt(dot(axis_angle, axis_angle) + 1e-12)
    v = (axis_angle / w) * torch.sin(w / 2)
    return torch.cat([torch.cos(w / 2)[None], v])


def inv_trans(pos, position, rotation):         # primitive_utils.py:43-46
    inv = torch.cat([rotation[0:1], -rotation[1:4]])
    inv = inv / torch.sqrt(dot(inv, inv))
    return qrot(inv, pos - position)


class SvdRef(torch.autograd.Function):
    """ti.svd (mpm_simulator.py:133) with the reference's own adjoint (backward_svd, :140-157).  Conventions of the Taichi routine
    [ext]: U, V proper rotations, sigma sorted by decreasing magnitude, the sign carried by the last one."""

    @staticmethod
    def forward(ctx, F):
        U, S, Vh = torch.linalg.svd(F)
        V = Vh.transpose(-1, -2).contiguous()
        du, dv = torch.linalg.det(U), torch.linalg.det(V)
        U = U.clone(); V = V.clone(); S = S.clone()
        fu, fv = du < 0, dv < 0
        U[fu, :, 2] *= -1; S[fu, 2] *= -1
        V[fv, :, 2] *= -1; S[fv, 2] *= -1
        ctx.save_for_backward(U, S, V)
        return U, S, V

    @staticmethod
    def backward(ctx, gu, gs, gv):
        u, s, v = ctx.saved_tensors
        sig = torch.diag_embed(s)
        gsigma = torch.diag_embed(gs)
        vt, ut = v.transpose(-1, -2), u.transpose(-1, -2)
        sigma_term = u @ gsigma @ vt
        s2 = s ** 2
        d = s2[:, None, :] - s2[:, :, None]                         # d[i, j] = s[j] - s[i]
        cl = torch.where(d >= 0, torch.clamp(d, min=1e-6), torch.clamp(d, max=-1e-6))      # clamp (:184-192)
        Fm = 1.0 / cl
        Fm = Fm * (1 - torch.eye(3, dtype=DT))                      # F[i, i] = 0
        u_term = u @ ((Fm * (ut @ gu - gu.transpose(-1, -2) @ u)) @ sig) @ vt
        v_term = u @ (sig @ ((Fm * (vt @ gv - gv.transpose(-1, -2) @ v)) @ vt))
        return u_term + v_term + sigma_term


class TorchPrimitive:
    """Mesh primitive: trilinear SDF / normal tables + the contact functions of primitive_base.py."""

    def __init__(self, sdf, normal, lower, upper, dx, friction, softness, enabled=True):
        self.sdf_table = torch.as_tensor(np.asarray(sdf), dtype=DT)
        self.normal_table = torch.as_tensor(np.asarray(normal), dtype=DT)
        self.lower, self.upper = torch.as_tensor(np.asarray(lower), dtype=DT), torch.as_tensor(np.asarray(upper), dtype=DT)
        self.inv_dx = 1.0 / float(dx)
        self.friction, self.softness, self.enabled = float(friction), float(softness), bool(enabled)
        self.res = torch.tensor(self.sdf_table.shape)

    # -- mesh.py:45-108 --------------------------------------------------------------------------------------------
    def _lookup(self, gp):
        in_box = ((gp >= self.lower) & (gp < self.upper)).all(-1)
        pos = (gp - self.lower) * self.inv_dx
        base = pos.detach().to(torch.int64)                         # ti.cast(pos, ti.i32): truncation, no gradient
        base = torch.minimum(torch.clamp(base, min=0), self.res - 2)        # only matters outside the box (masked below)
        fx = pos - base
        w = [1.0 - fx, fx]
        return in_box, base, w

    def _sdf(self, gp):
        in_box, base, w = self._lookup(gp)
        out = torch.zeros(gp.shape[0], dtype=DT)
        for i in (0, 1):
            for j in (0, 1):
                for k in (0, 1):
                    weight = w[i][:, 0] * w[j][:, 1] * w[k][:, 2]
                    out = out + weight * self.sdf_table[base[:, 0] + i, base[:, 1] + j, base[:, 2] + k]
        return torch.where(in_box, out, torch.full_like(out, INF))

    def _normal(self, gp):
        in_box, base, w = self._lookup(gp)
        nrm = torch.zeros(gp.shape[0], 3, dtype=DT)
        for i in (0, 1):
            for j in (0, 1):
                for k in (0, 1):
                    weight = w[i][:, 0] * w[j][:, 1] * w[k][:, 2]
                    nrm = nrm + weight[:, None] * self.normal_table[base[:, 0] + i, base[:, 1] + j, base[:, 2] + k]
        safe = torch.where(in_box[:, None], nrm, torch.tensor([0.0, 1.0, 0.0], dtype=DT).expand_as(nrm))
        nrm = safe / torch.sqrt(dot(safe, safe))[:, None]          # .normalized()
        return torch.where(in_box[:, None], nrm, torch.tensor([0.0, 1.0, 0.0], dtype=DT).expand_as(nrm))

    # -- primitive_base.py:53-70 -----------------------------------------------------------------------------------
    def sdf(self, s13, pos):
        return self._sdf(inv_trans(pos, s13[0:3], s13[3:7]))

    def normal(self, s13, pos):
        return qrot(s13[3:7], self._normal(inv_trans(pos, s13[0:3], s13[3:7])))

    def collider_v(self, s13, r):
        quat = s13[3:7] / torch.sqrt(dot(s13[3:7], s13[3:7]))
        inv = torch.cat([quat[0:1], -quat[1:4]])
        r_local = qrot(inv, r)
        local = s13[7:10] + torch.linalg.cross(s13[10:13].expand(r_local.shape), r_local)
        return qrot(quat, local)

    def _friction_projection(self, v_t, normal_component):        # shared by collide (:86-89) and collide_mixed (:153-156)
        nrm = length(v_t)
        fric = v_t / nrm[:, None] * ti_max(0.0, nrm + normal_component * self.friction)[:, None]
        flag = ((normal_component < 0) & (torch.sqrt(dot(v_t, v_t)) > 1e-30)).detach().to(DT)[:, None]
        return fric * flag + v_t * (1 - flag)

    # -- primitive_base.py:72-103: grid contact -----------------------------------------------------------------------
    def collide(self, s13, grid_pos, v_out, dt, grid_m):
        dist = self.sdf(s13, grid_pos)
        influence = ti_min(torch.exp(-torch.clamp(dist, max=1.0) * self.softness), 1.0)    # clamp only guards exp() of the 1e10 sentinel
        act = (((self.softness > 0) & (influence > 0.1)) | (dist <= 0)).detach()
        v_in = v_out
        D = self.normal(s13, grid_pos)
        r = grid_pos - s13[0:3]
        cv = self.collider_v(s13, r)
        input_v = v_out - cv
        nc = dot(input_v, D)
        v_t = input_v - ti_min(nc, 0.0)[:, None] * D
        v_t = self._friction_projection(v_t, nc)
        new = cv + input_v * (1 - influence)[:, None] + v_t * influence[:, None]
        v_new = torch.where(act[:, None], new, v_out)
        b_f = grid_m[:, None] * (v_in - v_new) * (1.0 / dt)
        b_t = torch.linalg.cross(r, b_f)
        m = act.to(DT)[:, None]
        return v_new, torch.cat([(b_f * m).sum(0), (b_t * m).sum(0)])

    # -- primitive_base.py:105-137: particle (penalty) contact --------------------------------------------------------
    def collide_particle(self, s13, p_pos, p_v, dt):
        dist = self.sdf(s13, p_pos)
        c = dist - 5e-3
        act = (c < 0.0).detach()
        D = self.normal(s13, p_pos)
        r = p_pos - s13[0:3]
        cv = self.collider_v(s13, r)
        input_v = p_v - cv
        nc = dot(input_v, D)
        v_t = input_v - nc[:, None] * D
        f1 = -D * (torch.where(act, c, torch.zeros_like(c)) * 50.0)[:, None]
        v_t_norm = torch.sqrt(dot(v_t, v_t) + 1e-8)
        f2 = -v_t / v_t_norm[:, None] * (torch.abs(nc) * self.friction)[:, None]
        m = act.to(DT)[:, None]
        p_f = (f1 + f2) * m
        b_f = -(f1 + f2) * m
        b_t = torch.linalg.cross(r, b_f)
        return p_f * dt, torch.cat([b_f.sum(0), b_t.sum(0)])

    # -- primitive_base.py:139-181: forecast-based contact ------------------------------------------------------------
    def collide_mixed(self, s13, p_pos, p_v, p_mass, dt, life):
        dist = self.sdf(s13, p_pos)
        act = (dist <= 5e-3).detach()
        p_v_in = p_v
        D = self.normal(s13, p_pos)
        r = p_pos - s13[0:3]
        cv = self.collider_v(s13, r)
        input_v = p_v - cv
        nc = dot(input_v, D)
        neg = (nc < 0).detach()
        v_t = input_v - nc[:, None] * D
        v_t = self._friction_projection(v_t, nc)
        pv1 = cv + v_t
        influence = ti_min(torch.exp(-torch.clamp(dist, max=1.0) * self.softness), 1.0)
        pv2 = cv + input_v * (1 - influence)[:, None] + v_t * influence[:, None]
        pv = torch.where((neg & (dist > 0).detach())[:, None], pv2, pv1)
        pv = torch.where(neg[:, None], pv, p_v)
        x_new = pv * dt + p_pos
        s2 = self.sdf(s13, x_new)
        pen = (s2 < 0).detach()
        n = self.normal(s13, x_new)
        pv = torch.where(pen[:, None], pv - (torch.where(pen, s2, torch.zeros_like(s2)) / dt)[:, None] * n * life, pv)
        out = torch.where(act[:, None], pv, p_v)
        b_f = p_mass * (p_v_in - out) * (1.0 / dt)
        b_t = torch.linalg.cross(r, b_f)
        m = act.to(DT)[:, None]
        return out, torch.cat([(b_f * m).sum(0), (b_t * m).sum(0)])


class TorchOracle:
    def __init__(self, n_grid=64, dt=2e-4, E=3e3, nu=0.2, gravity=(0., -9.8, 0.), ground_friction=20., material_model=0, ptype=0,
                 collision_type=2, substeps=5, n_control=0, rigid_velocity_control=False):
        self.n_grid, self.dt = int(n_grid), float(dt)
        self.dx, self.inv_dx = 1.0 / n_grid, float(n_grid)                      # mpm_simulator.py:32
        self.p_vol = (self.dx * 0.5) ** 2; self.p_mass = self.p_vol             # :34-35
        mu, lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))         # :41
        if ptype == 1:
            mu, lam = 0.3 * mu, 0.3 * lam                                       # :42-43
        elif ptype == 2:
            mu = 0.0                                                            # :44-45
        self.mu, self.lam = mu, lam
        self.gravity = torch.tensor(gravity, dtype=DT)
        self.sticky = ground_friction >= 10.0
        self.material_model, self.ptype, self.collision_type = material_model, ptype, collision_type
        self.substeps, self.n_control, self.vctrl = int(substeps), int(n_control), bool(rigid_velocity_control)
        self.prims = []
        self.plasticity, self.yield_stress = 0, 0.0     # 1: von Mises return mapping of the soft_cloth variant (set_plasticity)

    def set_plasticity(self, mode, yield_stress):
        self.plasticity, self.yield_stress = int(mode), float(yield_stress)

    def _von_mises(self, F_tmp, U, sig, V):
        """compute_von_mises, soft_cloth/engine/mpm_simulator.py:172-189 (3-D): log-strain return mapping; F stays F_tmp unless the
        particle yields.  norm(x) = sqrt(x.x + 1e-8) (:201-202)."""
        sig = ti_max(sig, 0.05)
        epsilon = torch.log(sig)
        epsilon_hat = epsilon - (epsilon.sum(-1, keepdim=True) / 3)
        epsilon_hat_norm = torch.sqrt(dot(epsilon_hat, epsilon_hat) + 1e-8)
        delta_gamma = epsilon_hat_norm - self.yield_stress / (2 * self.mu)
        yields = delta_gamma > 0
        eps_new = epsilon - (delta_gamma / epsilon_hat_norm)[:, None] * epsilon_hat
        F_y = U @ torch.diag_embed(torch.exp(eps_new)) @ V.transpose(-1, -2)
        return torch.where(yields[:, None, None], F_y, F_tmp)

    def add_primitive(self, *a, **k):
        self.prims.append(TorchPrimitive(*a, **k))

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _stencil(self, x):
        base = (x * self.inv_dx - 0.5).detach().to(torch.int64)                 # .cast(int): truncation, no gradient
        fx = x * self.inv_dx - base
        w = [0.5 * (1.5 - fx) ** 2, 0.75 - (fx - 1.0) ** 2, 0.5 * (fx - 0.5) ** 2]
        return base, fx, w

    def _lin(self, idx):
        n = self.n_grid
        return (idx[:, 0] * n + idx[:, 1]) * n + idx[:, 2]

    def _boundary(self, I, v):                                                  # mpm_simulator.py:268-281
        n, bound = self.n_grid, 3
        cols = [v[:, 0], v[:, 1], v[:, 2]]
        zero = torch.zeros_like(cols[0])
        for d in range(3):
            cols[d] = torch.where((I[:, d] < bound) & (cols[d] < 0).detach(), zero, cols[d])
            cols[d] = torch.where((I[:, d] > n - bound) & (cols[d] > 0).detach(), zero, cols[d])
            if d == 1 and self.sticky:
                st = I[:, 1] < bound
                cols = [torch.where(st, zero, c) for c in cols]
        return torch.stack(cols, -1)

    def _grid_velocity(self, grid_v_in, grid_m, s13s, f_unused, contact_grid):
        """grid_op (:283-297) / grid_op_mixed1 (:396-404) on the nodes with mass > 1e-10; returns the dense velocity field + wrenches."""
        n = self.n_grid
        nodes = torch.nonzero(grid_m.detach() > 1e-10)[:, 0]
        m = grid_m[nodes]
        v_out = (1.0 / m)[:, None] * grid_v_in[nodes] + self.dt * self.gravity
        I = torch.stack([nodes // (n * n), (nodes // n) % n, nodes % n], -1)
        wrench = [torch.zeros(6, dtype=DT) for _ in self.prims]
        if contact_grid:
            for i, p in enumerate(self.prims):
                if p.enabled:
                    v_out, w6 = p.collide(s13s[i], I.to(DT) * self.dx, v_out, self.dt, m)
                    wrench[i] = wrench[i] + w6
        v_out = self._boundary(I, v_out)
        return torch.zeros(n ** 3, 3, dtype=DT).index_add(0, nodes, v_out), wrench

    # -- one substep (mpm_simulator.py:320-337) -------------------------------------------------------------------
    def substep(self, f, state, s13s=(), action=None, control_idx=None):
        """state (n, 24) = [x v F C] (get_state layout, :481-489); s13s: per primitive [pos(3) quat(4) v(3) w(3)] of frame f.
        Returns (state of frame f+1, wrench contributions [6] per primitive, primitive states of frame f+1 under velocity control)."""
        dt, n = self.dt, self.n_grid
        N = state.shape[0]
        x, v = state[:, 0:3], state[:, 3:6]
        F, C = state[:, 6:15].reshape(N, 3, 3), state[:, 15:24].reshape(N, 3, 3)
        eye = torch.eye(3, dtype=DT)
        wrench = [torch.zeros(6, dtype=DT) for _ in self.prims]
        # compute_F_tmp (:125-128), svd (:130-133)
        F_tmp = (eye + dt * C) @ F
        if self.material_model == 0:
            U, sig, V = SvdRef.apply(F_tmp)
        # p2g (:198-262)
        collision_impulse = torch.zeros(N, 3, dtype=DT)
        if self.collision_type == 1:
            for i, p in enumerate(self.prims):
                if p.enabled:
                    imp, w6 = p.collide_particle(s13s[i], x, v, dt)
                    collision_impulse = collision_impulse + imp
                    wrench[i] = wrench[i] + w6
        control_impulse = torch.zeros(N, 3, dtype=DT)
        if self.n_control > 0:
            ci = torch.as_tensor(control_idx, dtype=torch.int64)
            on = ci >= 0
            control_impulse = torch.where(on[:, None], 6e-4 * action[torch.clamp(ci, min=0)] * dt, control_impulse)
        base, fx, w = self._stencil(x)
        new_F = F_tmp
        J = torch.linalg.det(F_tmp)
        if self.material_model == 0:
            if self.ptype == 0 and self.plasticity == 1:
                new_F = self._von_mises(F_tmp, U, sig, V)
            elif self.ptype == 0:
                sig_new = ti_min(ti_max(sig, 1 - 2e-3), 1 + 3e-3)
                new_F = U @ torch.diag_embed(sig_new) @ V.transpose(-1, -2)
            elif self.ptype == 2:
                new_F = eye * torch.pow(J, 1.0 / 3)[:, None, None]
            r = U @ V.transpose(-1, -2)
            stress = 2 * self.mu * (new_F - r) @ new_F.transpose(-1, -2) + eye * (self.lam * J * (J - 1))[:, None, None]
        else:
            if self.ptype == 2:
                sq = torch.sqrt(J)
                new_F = torch.diag_embed(torch.stack([sq, sq, torch.ones_like(sq)], -1))
            stress = self.mu * (new_F @ new_F.transpose(-1, -2)) + eye * (self.lam * torch.log(J) - self.mu)[:, None, None]
        stress = (-dt * self.p_vol * 4 * self.inv_dx * self.inv_dx) * stress
        affine = stress + self.p_mass * C
        grid_v_in = torch.zeros(n ** 3, 3, dtype=DT)
        grid_m = torch.zeros(n ** 3, dtype=DT)
        offs = [(a, b, c) for a in range(3) for b in range(3) for c in range(3)]
        for (a, b, c) in offs:
            off = torch.tensor([a, b, c], dtype=DT)
            dpos = (off - fx) * self.dx
            weight = w[a][:, 0] * w[b][:, 1] * w[c][:, 2]
            idx = self._lin(base + torch.tensor([a, b, c]))
            val = weight[:, None] * (self.p_mass * v + (affine @ dpos[:, :, None])[:, :, 0] + collision_impulse + control_impulse)
            grid_v_in = grid_v_in.index_add(0, idx, val)
            grid_m = grid_m.index_add(0, idx, weight * self.p_mass)
        # forward_kinematics (:329-331; primitive_base.py:280-283)
        next_s13s = None
        if self.vctrl:
            next_s13s = []
            for s in s13s:
                pos = s[0:3] + s[7:10] * dt
                rot = qmul(w2quat(s[10:13] * dt), s[3:7])
                next_s13s.append(torch.cat([pos, rot]))
        # grid update
        if self.collision_type == 2:
            grid_v_mixed, _ = self._grid_velocity(grid_v_in, grid_m, s13s, f, False)             # mixed1
            grid_v_out = grid_v_mixed                                                             # grid_v_out (cleared) += grid_v_mixed
            v_tmp = torch.zeros(N, 3, dtype=DT)                                                   # mixed2
            for (a, b, c) in offs:
                weight = w[a][:, 0] * w[b][:, 1] * w[c][:, 2]
                v_tmp = v_tmp + weight[:, None] * grid_v_mixed[self._lin(base + torch.tensor([a, b, c]))]
            life = float(np.float32(1.0) / np.float32(self.substeps - f % self.substeps))       # mixed3 (f32 in Taichi)
            v_tgt = v_tmp
            for i, p in enumerate(self.prims):
                if p.enabled:
                    v_tgt, w6 = p.collide_mixed(s13s[i], x, v_tgt, self.p_mass, dt, life)
                    wrench[i] = wrench[i] + w6
            for (a, b, c) in offs:                                                                # mixed4
                weight = w[a][:, 0] * w[b][:, 1] * w[c][:, 2]
                idx = self._lin(base + torch.tensor([a, b, c]))
                on = (grid_m.detach()[idx] > 1e-10).to(DT)
                grid_v_out = grid_v_out.index_add(0, idx, -2.0 * (weight * on)[:, None] * (v_tmp - v_tgt))
        else:
            grid_v_out, wg = self._grid_velocity(grid_v_in, grid_m, s13s, f, self.collision_type == 0)
            wrench = [a + b for a, b in zip(wrench, wg)]
        # g2p (:299-318)
        new_v = torch.zeros(N, 3, dtype=DT)
        new_C = torch.zeros(N, 3, 3, dtype=DT)
        for (a, b, c) in offs:
            dpos = torch.tensor([a, b, c], dtype=DT) - fx
            g_v = grid_v_out[self._lin(base + torch.tensor([a, b, c]))]
            weight = w[a][:, 0] * w[b][:, 1] * w[c][:, 2]
            new_v = new_v + weight[:, None] * g_v
            new_C = new_C + 4 * self.inv_dx * weight[:, None, None] * (g_v[:, :, None] * dpos[:, None, :])
        new_x = x + dt * new_v
        out = torch.cat([new_x, new_v, new_F.reshape(N, 9), new_C.reshape(N, 9)], -1)
        return out, wrench, next_s13s

    # -- forward + vector-Jacobian product of one substep ---------------------------------------------------------------
    def substep_with_vjp(self, f, state, s13s, cot_state, ext_f_grads, action=None, control_idx=None, cot_next_pose=None):
        """Returns (frame f+1, wrenches, next poses) and, for the objective cot_state . frame[f+1] + sum_i ext_f_grad_i . wrench_i
        (+ cot_next_pose_i . pose_i[f+1] under velocity control), the gradients w.r.t. frame f, each primitive's 13-state and the action."""
        st = torch.as_tensor(np.asarray(state), dtype=DT).clone().requires_grad_(True)
        ps = [torch.as_tensor(np.asarray(s), dtype=DT).clone().requires_grad_(True) for s in s13s]
        act = None if action is None else torch.as_tensor(np.asarray(action), dtype=DT).clone().requires_grad_(True)
        out, wrench, nxt = self.substep(f, st, ps, act, control_idx)
        L = (out * torch.as_tensor(np.asarray(cot_state), dtype=DT)).sum()
        for i, g in enumerate(ext_f_grads):
            L = L + (wrench[i] * torch.as_tensor(np.asarray(g), dtype=DT)).sum()
        if nxt is not None and cot_next_pose is not None:
            for i, g in enumerate(cot_next_pose):
                L = L + (nxt[i] * torch.as_tensor(np.asarray(g), dtype=DT)[:7]).sum()
        inputs = [st] + ps + ([act] if act is not None else [])
        grads = torch.autograd.grad(L, inputs, allow_unused=True)
        z = lambda g, ref: np.zeros(tuple(ref.shape)) if g is None else g.detach().numpy()
        g_state = z(grads[0], st)
        g_prims = [z(grads[1 + i], ps[i]) for i in range(len(ps))]
        g_act = z(grads[-1], act) if act is not None else None
        return (out.detach().numpy(), [w.detach().numpy() for w in wrench], None if nxt is None else [p.detach().numpy() for p in nxt],
                g_state, g_prims, g_act)
