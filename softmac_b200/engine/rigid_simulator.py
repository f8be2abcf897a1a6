"""Stand-in for the Jade bridge ``softmac/engine/rigid_simulator.py:RigidSimulator``.

Jade (nimblephysics) is an external CPU simulator that is not installable here and is OUT OF SCOPE (SURVEY.md 2.1
row 4); what IS on the hot path is its interface to the primitives: once per env step it reads the averaged contact
wrench ``primitive.ext_f / substeps`` (rigid_simulator.py:92-93), advances the bodies, writes pose + twist of the next
``substeps`` frames with ``set_all_states`` (:200-201) and, in the backward pass, pulls ``get_all_states_grad`` (:207-208)
and pushes ``ext_f_grad`` (:166-168).  This class reproduces exactly that call pattern around a tiny articulated-body
integrator (fixed / prismatic / revolute / free joints, semi-implicit Euler) whose Jacobians are taken by central differences in
f64 (state dimension <= 12 per body; closed form for fixed / prismatic / revolute joints, ``_pose_jac``), so the coupling loop and
its gradient chain can be run and tested end to end.  The revolute joint covers the door scene (config/demo_door_config.py:31-56,
assets/door/door.urdf: one hinge about y through the link origin).
"""
import numpy as np


def _quat_mul(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])


def _exp2quat(e):
    th = np.linalg.norm(e)
    if th < 1e-12:
        return np.array([1.0, 0.5 * e[0], 0.5 * e[1], 0.5 * e[2]])
    return np.concatenate([[np.cos(th / 2)], np.sin(th / 2) * e / th])


def _quat_rot(q, v):
    qv = q[1:]
    uv = np.cross(qv, v)
    return v + 2 * (q[0] * uv + np.cross(qv, uv))


class Body:
    def __init__(self, joint="fixed", origin=(0, 0, 0), quat=(1, 0, 0, 0), axis=(1, 0, 0), mass=1.0, inertia=1.0, gravity=True):
        assert joint in ("fixed", "prismatic", "revolute", "free")
        self.joint, self.origin, self.quat0 = joint, np.asarray(origin, float), np.asarray(quat, float)
        self.axis = np.asarray(axis, float) / np.linalg.norm(axis)
        self.mass, self.inertia, self.gravity = float(mass), float(inertia), gravity
        self.ndof = {"fixed": 0, "prismatic": 1, "revolute": 1, "free": 6}[joint]


def bodies_from_urdf(urdf_path):
    """One Body per link with a collision mesh, in document order (the order ``Primitives`` creates the primitives in,
    primitives.py:25-41): joint type (fixed / prismatic / floating -> free), axis, origin accumulated along the parent
    chain (translations only; the reference assets use rpy = 0), mass of the child link."""
    import xml.etree.ElementTree as ET
    root = ET.parse(urdf_path).getroot()
    joints = {j.find("child").attrib["link"]: j for j in root.findall("joint")}

    def origin_of(link):
        o = np.zeros(3)
        while link in joints:
            j = joints[link]
            org = j.find("origin")
            if org is not None:
                o += np.array([float(v) for v in org.attrib.get("xyz", "0 0 0").split()])
            link = j.find("parent").attrib["link"]
        return o

    out = []
    for link in root.findall("link"):
        if link.find("collision/geometry/mesh") is None:
            continue
        name = link.attrib["name"]
        j = joints.get(name)
        jt = j.attrib.get("type", "fixed") if j is not None else "floating"
        joint = {"fixed": "fixed", "prismatic": "prismatic", "floating": "free", "revolute": "revolute", "continuous": "revolute"}.get(jt)
        if joint is None:
            raise NotImplementedError(f"stand-in rigid simulator: joint type {jt!r} of link {name!r}")
        axis = (1, 0, 0)
        if j is not None and j.find("axis") is not None:
            axis = tuple(float(v) for v in j.find("axis").attrib["xyz"].split())
        mass = link.find("inertial/mass")
        spec = dict(joint=joint, axis=axis, origin=tuple(origin_of(name)), mass=float(mass.attrib["value"]) if mass is not None else 1.0,
                    gravity=False)
        it = link.find("inertial/inertia")
        if joint == "revolute" and it is not None:      # moment of inertia about the hinge axis (the link frame sits on the hinge)
            g = lambda k: float(it.attrib.get(k, 0.0))
            I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
            a = np.asarray(axis, float) / np.linalg.norm(axis)
            spec["inertia"] = float(a @ I @ a)
        out.append(spec)
    return out


class RigidSimulator:
    def __init__(self, cfg, primitives, substeps=20, env_dt=2e-3, bodies=None, fp32_bridge=True):
        self.cfg, self.primitives = cfg, primitives
        # the Jade bridge truncates wrench and body state to float32 (torch.FloatTensor, rigid_simulator.py:92,185)
        self.fp32_bridge = fp32_bridge
        self.n_primitive = len(primitives)
        self.substeps, self.dt = substeps, env_dt
        self.gravity = np.asarray(getattr(cfg, "gravity", (0., 0., 0.)), float)
        specs = bodies if bodies is not None else getattr(cfg, "bodies", [])
        self.bodies = [b if isinstance(b, Body) else Body(**b) for b in specs]
        assert len(self.bodies) == self.n_primitive, "one body per primitive (rigid_simulator.py:42)"
        self.offsets = np.cumsum([0] + [b.ndof for b in self.bodies])
        self.state_dim_half = int(self.offsets[-1])
        self.state_dim = 2 * self.state_dim_half
        self.action_dim = self.state_dim_half
        init = np.asarray(getattr(cfg, "init_state", ()), float)
        self.init_state = init if init.size == self.state_dim else np.zeros(self.state_dim)
        self.ext_grad_scale = 1.0
        self.transform_action = False
        self.obs_ext_f = np.zeros(6 * self.n_primitive)
        self.state_grad = np.zeros(self.state_dim)
        self.states = []
        self._clear_tape()

    def _clear_tape(self):
        self.jacob_ds_df, self.jacob_ds_ds, self.jacob_ds_da, self.jacob_external = [], [], [], []

    def initialize(self):
        pass

    # -- dynamics ------------------------------------------------------------------------------------------
    def _pose(self, state, i):
        """[x(3) q(4) v(3) w(3)] of body i: position / quaternion in the world, body-frame twist (rigid_simulator.py:176-186)."""
        b, o, h = self.bodies[i], self.offsets[i], self.state_dim_half
        q, qd = state[o:o + b.ndof], state[h + o:h + o + b.ndof]
        if b.joint == "fixed":
            return np.concatenate([b.origin, b.quat0, np.zeros(6)])
        if b.joint == "prismatic":
            ax_w = _quat_rot(b.quat0, b.axis)
            return np.concatenate([b.origin + ax_w * q[0], b.quat0, b.axis * qd[0], np.zeros(3)])
        if b.joint == "revolute":       # hinge through the link origin: R = R0 Rot(axis, theta); body-frame twist (0, axis * omega)
            qt = np.concatenate([[np.cos(0.5 * q[0])], np.sin(0.5 * q[0]) * b.axis])
            return np.concatenate([b.origin, _quat_mul(b.quat0, qt), np.zeros(3), b.axis * qd[0]])
        quat = _quat_mul(_exp2quat(q[:3]), b.quat0)
        inv = np.array([quat[0], -quat[1], -quat[2], -quat[3]])
        return np.concatenate([b.origin + q[3:], quat, _quat_rot(inv, qd[3:]), _quat_rot(inv, qd[:3])])

    def _advance(self, state, action, wrenches):
        h, dt = self.state_dim_half, self.dt
        new = state.copy()
        for i, b in enumerate(self.bodies):
            o = self.offsets[i]
            f, tq = wrenches[6 * i:6 * i + 3], wrenches[6 * i + 3:6 * i + 6]
            a = action[o:o + b.ndof]
            if b.joint == "prismatic":
                ax_w = _quat_rot(b.quat0, b.axis)
                g = self.gravity if b.gravity else 0.0
                qd = state[h + o] + dt * (a[0] + ax_w @ (f + b.mass * g)) / b.mass
                new[h + o] = qd
                new[o] = state[o] + dt * qd
            elif b.joint == "revolute":
                # generalized force = hinge axis (world) . torque about the link origin; a force through the hinge does no work, and
                # gravity has no moment about the vertical hinge of the door (a tilted hinge would make _advance non-affine: not modelled)
                ax_w = _quat_rot(b.quat0, b.axis)
                qd = state[h + o] + dt * (a[0] + ax_w @ tq) / b.inertia
                new[h + o] = qd
                new[o] = state[o] + dt * qd
            elif b.joint == "free":
                g = self.gravity if b.gravity else 0.0
                w = state[h + o:h + o + 3] + dt * (a[:3] + tq) / b.inertia
                v = state[h + o + 3:h + o + 6] + dt * ((a[3:] + f) / b.mass + g)
                new[h + o:h + o + 3], new[h + o + 3:h + o + 6] = w, v
                new[o:o + 3] = state[o:o + 3] + dt * w          # small-rotation update of the exponential coordinates
                new[o + 3:o + 6] = state[o + 3:o + 6] + dt * v
        return new

    def _pose_jac(self, state, i):
        """d pose_i / d state, (13, state_dim): closed form for fixed / prismatic / revolute joints, central differences for the
        free joint (exponential coordinates)."""
        b, o, h = self.bodies[i], self.offsets[i], self.state_dim_half
        if b.joint == "free":
            return self._jac(lambda x: self._pose(x, i), state)
        J = np.zeros((13, self.state_dim))
        if b.joint == "prismatic":
            J[0:3, o] = _quat_rot(b.quat0, b.axis)
            J[7:10, h + o] = b.axis
        elif b.joint == "revolute":
            th = state[o]
            dqt = np.concatenate([[-0.5 * np.sin(0.5 * th)], 0.5 * np.cos(0.5 * th) * b.axis])
            J[3:7, o] = _quat_mul(b.quat0, dqt)
            J[10:13, h + o] = b.axis
        return J

    @staticmethod
    def _jac(fn, x, eps=1e-6):
        x = np.asarray(x, float)
        y0 = fn(x)
        J = np.zeros((y0.size, x.size))
        for k in range(x.size):
            d = np.zeros_like(x); d[k] = eps
            J[:, k] = (fn(x + d) - fn(x - d)) / (2 * eps)
        return J

    # -- reference interface -----------------------------------------------------------------------------------
    def reset(self):
        self.states = [self.init_state.copy()]
        self._clear_tape()
        self.state_grad = np.zeros(self.state_dim)
        self.set_ext_state(-1)

    def step(self, s, action=None):
        if self.n_primitive == 0:
            return
        wr = np.zeros(6 * self.n_primitive)
        for i in range(self.n_primitive):
            ext_f = np.asarray(self.primitives[i].ext_f.to_numpy(), dtype=np.float64)
            if self.fp32_bridge:
                ext_f = ext_f.astype(np.float32).astype(np.float64)
            ext_f = ext_f / self.substeps
            self.obs_ext_f[6 * i:6 * i + 6] = ext_f
            if (np.abs(ext_f) > 1e-10).any() and self.primitives[i].enable_external_force:
                wr[6 * i:6 * i + 6] = ext_f
            self.primitives[i].clear_ext_f()
        a = np.zeros(self.action_dim) if action is None else np.asarray(action, dtype=np.float64).reshape(-1)
        st = self.states[-1]
        self.states.append(self._advance(st, a, wr))
        self.jacob_ds_ds.append(self._jac(lambda x: self._advance(x, a, wr), st))
        self.jacob_ds_da.append(self._jac(lambda x: self._advance(st, x, wr), a) if a.size else np.zeros((self.state_dim, 0)))
        Jf = self._jac(lambda x: self._advance(st, a, x), wr)
        mask = np.repeat([(np.abs(wr[6 * i:6 * i + 6]) > 0).any() for i in range(self.n_primitive)], 6)
        self.jacob_ds_df.append([Jf[:, 6 * i:6 * i + 6] * mask[6 * i] for i in range(self.n_primitive)])
        self.set_ext_state(s)

    def step_grad(self, s, action=None):
        if self.n_primitive == 0:
            return None, []
        self.state_grad = self.state_grad + self.get_ext_state_grad(s + 1) * self.ext_grad_scale
        action_grad = self.state_grad @ self.jacob_ds_da[s]
        state_grad_t = self.state_grad @ self.jacob_ds_ds[s]
        ext_f_grad_list = []
        for i in range(self.n_primitive):
            ext_f_grad_list.append(self.state_grad @ self.jacob_ds_df[s][i] / self.substeps)
            self.primitives[i].clear_ext_f()
        self.state_grad = state_grad_t
        return action_grad, ext_f_grad_list

    def set_ext_state(self, s):
        st = self.states[-1]
        self.jacob_external.append([self._pose_jac(st, i) for i in range(self.n_primitive)])
        for i in range(self.n_primitive):
            pose = self._pose(st, i)
            if self.fp32_bridge:
                pose = pose.astype(np.float32).astype(np.float64)
            for j in range((s + 1) * self.substeps, (s + 2) * self.substeps):
                self.primitives[i].set_all_states(j, pose)

    def get_ext_state_grad(self, s):
        g = np.zeros(self.state_dim)
        for i in range(self.n_primitive):
            tmp = np.zeros(13)
            for j in range(s * self.substeps, (s + 1) * self.substeps):
                tmp += self.primitives[i].get_all_states_grad(j)
            g += tmp @ self.jacob_external[s][i]
        return g
