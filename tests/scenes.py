"""Seeded synthetic scenes shared by the oracle tests (CPU) and the CUDA parity tests (GPU)."""
import numpy as np


def sphere_table(radius=0.08, dx=0.01, margin=0.05):
    """Analytic SDF / normal table of a sphere centred at the primitive origin, in the on-disk layout of
    softmac/engine/primitive/mesh.py:235-241 (sdf[res], normal[res,3], position=(lower, upper), dx)."""
    half = radius + margin
    res = int(np.ceil(2 * half / dx)) + 1
    lower = -np.ones(3) * (res - 1) * dx / 2
    upper = lower + (res - 1) * dx
    ax = lower[0] + np.arange(res) * dx
    P = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1)
    r = np.linalg.norm(P, axis=-1)
    sdf = r - radius
    normal = P / np.maximum(r, 1e-9)[..., None]
    normal[r < 1e-9] = (0, 1, 0)
    return dict(sdf=sdf, normal=normal, lower=lower, upper=upper, dx=dx)


def box_table(half=(0.10, 0.04, 0.10), dx=0.01, margin=0.04):
    """Exact box SDF with nearest-face normals (piecewise constant, like the shipped tables)."""
    half = np.asarray(half, float)
    res = (np.ceil(2 * (half + margin) / dx)).astype(int) + 1
    lower = -(res - 1) * dx / 2
    upper = lower + (res - 1) * dx
    axs = [lower[i] + np.arange(res[i]) * dx for i in range(3)]
    P = np.stack(np.meshgrid(*axs, indexing="ij"), -1)
    q = np.abs(P) - half
    outside = np.linalg.norm(np.maximum(q, 0), axis=-1)
    inside = np.minimum(q.max(-1), 0)
    sdf = outside + inside
    k = q.argmax(-1)
    normal = np.zeros_like(P)
    idx = np.indices(k.shape)
    normal[idx[0], idx[1], idx[2], k] = np.sign(P[idx[0], idx[1], idx[2], k]) + (P[idx[0], idx[1], idx[2], k] == 0)
    return dict(sdf=sdf, normal=normal, lower=lower, upper=upper, dx=dx)


def random_quat(rng, spread=0.3):
    q = np.array([1.0, 0, 0, 0]) + spread * rng.normal(size=4)
    return q / np.linalg.norm(q)


def blob_state(n, rng, center=(0.5, 0.3, 0.5), width=0.12, vel=0.5, Fdev=0.02, Cdev=2.0, fp32=True):
    """(n,24) state [x v F C]: a random cloud with non-trivial v, F, C."""
    x = (rng.random((n, 3)) * 2 - 1) * 0.5 * width + np.asarray(center)
    v = vel * rng.normal(size=(n, 3))
    F = np.eye(3)[None] + Fdev * rng.normal(size=(n, 3, 3))
    Cm = Cdev * rng.normal(size=(n, 3, 3))
    st = np.hstack([x, v, F.reshape(n, 9), Cm.reshape(n, 9)])
    if fp32:
        st = st.astype(np.float32).astype(np.float64)
    return st


def cube_state(n, seed=0, init_pos=(0.5, 0.30, 0.5), width=0.390625):
    """The reference generator Shapes.add_box (softmac/engine/shapes/shape_maker.py:51-60) with
    np.random.seed(0): x uniform in the box, v = 0, F = I, C = 0."""
    state = np.random.get_state()
    np.random.seed(seed)
    p = (np.random.random((n, 3)) * 2 - 1) * (0.5 * np.array([width] * 3)) + np.array(init_pos)
    np.random.set_state(state)
    st = np.zeros((n, 24))
    st[:, :3] = p
    st[:, 6] = st[:, 10] = st[:, 14] = 1.0
    return st.astype(np.float32).astype(np.float64)


def contact_rollout_state(n, rng, sphere_center, radius=0.08, gap=5e-4, width=0.10, speed=1.0):
    """Particles in a box resting just outside a sphere (none starts inside it), moving towards it: a scene that
    stays physical over a multi-substep rollout (particles that start deep inside a primitive are ejected at
    sdf/dt ~ 100 m/s by the forecast contact model and leave the domain within a few substeps)."""
    c = np.asarray(sphere_center, float)
    box_c = c + np.array([0.0, radius + 0.5 * width - 0.02, 0.0])
    x = np.zeros((0, 3))
    while len(x) < n:
        cand = (rng.random((2 * n, 3)) * 2 - 1) * 0.5 * width + box_c
        cand = cand[np.linalg.norm(cand - c, axis=1) > radius + gap]
        x = np.vstack([x, cand])
    x = x[:n]
    v = np.tile([0.0, -speed, 0.0], (n, 1)) + 0.05 * rng.normal(size=(n, 3))
    F = np.eye(3)[None] + 0.003 * rng.normal(size=(n, 3, 3))
    Cm = 0.5 * rng.normal(size=(n, 3, 3))
    st = np.hstack([x, v, F.reshape(n, 9), Cm.reshape(n, 9)])
    return st.astype(np.float32).astype(np.float64)


def write_demo_assets(dirpath):
    """Materialises the demo_grip / demo_pour rigid assets (gripper = palm + two fingers, glass, bowl) as OBJ + URDF files under
    `dirpath` from the arrays in tests/golden/demo_meshes.npz (collision meshes and per-body joint type / origin / axis / mass, read
    off the reference's assets by tests/golden/make_fixtures.py --meshes), for the URDF-driven builders (``Primitives(cfgs)``,
    ``bodies_from_urdf``).  Returns {"gripper": urdf path, "glass": ..., "bowl": ...}."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demo_meshes.npz"))
    out = {}
    for name in ("gripper", "glass", "bowl"):
        d = os.path.join(dirpath, name)
        os.makedirs(d, exist_ok=True)
        k = 0
        while f"{name}_mesh{k}_V" in z.files:
            V, F = z[f"{name}_mesh{k}_V"], z[f"{name}_mesh{k}_F"]
            with open(os.path.join(d, f"mesh{k}.obj"), "w") as fh:
                fh.write("".join("v %.17g %.17g %.17g\n" % tuple(v) for v in V))
                fh.write("".join("f %d %d %d\n" % tuple(f + 1) for f in F))
            k += 1
        links = z[f"{name}_links"]
        xml = ['<?xml version="1.0" ?>', f'<robot name="{name}">', '  <link name="world"/>']
        for i, row in enumerate(links):
            jt, mesh, parent = int(row[0]), int(row[1]), int(row[2])
            org, ax, mass, rgba = row[3:6], row[6:9], row[9], row[10:14]
            xml += [f'  <joint name="joint{i}" type="{("fixed", "prismatic", "floating")[jt]}">',
                    f'    <parent link="{"world" if parent == 0 else "link0"}"/> <child link="link{i}"/>',
                    '    <origin xyz="%.17g %.17g %.17g" rpy="0 0 0"/> <axis xyz="%.17g %.17g %.17g"/>' % (*org, *ax), '  </joint>',
                    f'  <link name="link{i}">', '    <inertial> <mass value="%.17g"/> </inertial>' % mass,
                    f'    <visual> <geometry> <mesh filename="mesh{mesh}.obj"/> </geometry> <material name="m{i}"> '
                    '<color rgba="%.6g %.6g %.6g %.6g"/> </material> </visual>' % tuple(rgba),
                    f'    <collision> <geometry> <mesh filename="mesh{mesh}.obj"/> </geometry> </collision>', '  </link>']
        xml.append('</robot>')
        out[name] = os.path.join(d, f"{name}.urdf")
        with open(out[name], "w") as fh:
            fh.write("\n".join(xml) + "\n")
    return out
