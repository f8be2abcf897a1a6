"""Mirrors of the reference's loss classes (softmac/engine/losses/loss_{grip,pour,door,transport}.py) in
softmac_b200/engine/losses.py: values against a plain numpy restatement of the Taichi kernels, seeds against finite differences."""
import numpy as np
import pytest

from harness import rel_l2


def fd(fun, x, eps=1e-6):
    g = np.zeros_like(x)
    for i in range(x.size):
        a, b = x.copy(), x.copy()
        a.flat[i] += eps; b.flat[i] -= eps
        g.flat[i] = (fun(a) - fun(b)) / (2 * eps)
    return g


def test_rigid_terms_match_the_reference_formulas_and_their_derivatives():
    from softmac_b200.engine import losses as L
    rng = np.random.default_rng(5)
    for cls in (L.GripLoss, L.PourLoss, L.DoorLoss, L.TransportLoss):
        obj = cls.__new__(cls)
        obj.target = [0.3, 0.2, 0.6]
        for _ in range(4):
            s = rng.normal(size=13)
            s[3] = rng.choice([-1, 1]) * rng.uniform(0.3, 1.0)        # exercises both |q_w| windows of GripLoss
            val, g = obj.pose(s)
            assert np.allclose(g, fd(lambda z: obj.pose(z)[0], s), atol=1e-6)
            if cls is L.GripLoss:       # loss_grip.py:76-82
                a = abs(s[3])
                assert np.isclose(val, 10 * (s[1] - 0.4) ** 2 + min(0, a - 0.5) ** 2 + max(0, a - 0.9) ** 2)
            if cls is L.PourLoss:       # loss_pour.py:76-82 (rotation terms commented out)
                assert np.isclose(val, 10 * (s[1] - 0.4) ** 2)
            if cls is L.DoorLoss:       # loss_door.py:33-34
                assert np.isclose(val, (s[3] - np.cos(np.pi / 8)) ** 2)
            if cls is L.TransportLoss:  # loss_transport.py:38-41
                assert np.isclose(val, ((s[:3] - np.array(obj.target)) ** 2).sum())
    s = rng.normal(size=13)
    for w_ang in (0.1, 0.0):            # loss_grip.py:85-88, loss_door.py:43-44
        val, g = L._RigidTerms.velocity(s, w_ang)
        assert np.isclose(val, s[7:10] @ s[7:10] + w_ang * (s[10:13] @ s[10:13]))
        assert np.allclose(g, fd(lambda z: L._RigidTerms.velocity(z, w_ang)[0], s), atol=1e-6)


@pytest.mark.parametrize("cls_name,groups", [("DoorLoss", 1), ("TransportLoss", 2)])
def test_contact_distance_term(cls_name, groups):
    from softmac_b200.engine import losses as L
    cls = getattr(L, cls_name)
    rng = np.random.default_rng(6)
    obj = cls.__new__(cls)
    n = 40
    obj.n_particles_per_controller = n // groups
    x = rng.uniform(0.2, 0.8, size=(n, 3))
    pos = np.array([1.2, 0.5, 0.4])                                # away from every particle: the max(.., 0) is inactive
    val, gx, gp = obj.contact(x, pos)
    npc = n // groups
    ref = sum(min(np.maximum(((x[k * npc:(k + 1) * npc] - pos) ** 2).sum(1) - 0.01, 0)) ** 2 for k in range(groups))   # loss_door.py:46-56
    assert np.isclose(val, ref) and val > 0
    assert np.allclose(gx, fd(lambda z: obj.contact(z, pos)[0], x), atol=1e-6)
    assert np.allclose(gp, fd(lambda z: obj.contact(x, z)[0], pos), atol=1e-6)
    assert np.count_nonzero(np.abs(gx).sum(1)) == groups        # only the closest particle of every group


@pytest.mark.gpu
def test_grip_loss_on_the_simulator_matches_the_numpy_restatement():
    """GripLoss.compute_loss(f): Chamfer on the GPU + pose / velocity terms of primitives[0]; the seeds land in x.grad[f] and in the
    primitive's state adjoint exactly as the Taichi tape would leave them (loss_grip.py:45-140)."""
    import scenes
    from harness import Pair
    from softmac_b200.engine.losses import GripLoss, ChamferLoss
    rng = np.random.default_rng(7)
    n = 1500
    pair = Pair(n, tables=[scenes.sphere_table()], prim_params=[(0.5, 666.)], max_steps=6)
    s13 = np.concatenate([[0.5, 0.47, 0.5], [0.95, 0.1, 0.2, 0.1], [0.1, -0.2, 0.05], [0.3, 0.0, -0.4]])
    pair.prims[0].set_all_states(0, s13, f_end=6)
    pair.gpu.reset(scenes.blob_state(n, rng))
    pair.gpu.substep(0)
    target = pair.gpu.get_x(1) * np.array([0.9, 1.05, 1.0]) + rng.normal(size=(n, 3)) * 1e-3
    loss = GripLoss(dict(weight=(2.0, 0.5, 0.25)), pair.gpu)
    loss.initialize()
    loss.set_target(target)
    pair.gpu.clear_all_gradients()
    info = loss.compute_loss(1)
    # numpy restatement of the Chamfer part on the host
    x = pair.gpu.get_x(1)
    d2 = ((x[:, None, :] - target[None, :, :]) ** 2).sum(-1)
    i_cur, i_tar = d2.argmin(1), d2.argmin(0)
    ch = ((x - target[i_cur]) ** 2).sum() + ((x[i_tar] - target) ** 2).sum()
    assert abs(info["chamfer_loss"] / (2.0 * ch) - 1) < 1e-5
    gx = 2 * (x - target[i_cur])
    np.add.at(gx, i_tar, 2 * (x[i_tar] - target))
    seed = pair.gpu.get_state_grad(1)
    assert rel_l2(seed[:, :3], 2.0 * gx) <= 1e-5 and np.abs(seed[:, 3:]).max() == 0
    st = pair.prims[0].get_all_states(1)
    a = abs(st[3])
    pose = 10 * (st[1] - 0.4) ** 2 + min(0, a - 0.5) ** 2 + max(0, a - 0.9) ** 2
    vel = st[7:10] @ st[7:10] + 0.1 * (st[10:13] @ st[10:13])
    assert np.isclose(info["pose_loss"], 0.5 * pose, rtol=1e-6) and np.isclose(info["vel_loss"], 0.25 * vel, rtol=1e-6)
    assert np.isclose(info["loss"], info["chamfer_loss"] + info["pose_loss"] + info["vel_loss"])
    g13 = pair.prims[0].get_all_states_grad(1)
    exp = np.zeros(13)
    exp[1] = 0.5 * 20 * (st[1] - 0.4)
    exp[3] = 0.5 * (2 * min(0, a - 0.5) + 2 * max(0, a - 0.9)) * np.sign(st[3])
    exp[7:10] = 0.25 * 2 * st[7:10]; exp[10:13] = 0.25 * 0.2 * st[10:13]
    assert np.allclose(g13, exp, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("cls_name", ["DoorLoss", "TransportLoss"])
def test_contact_distance_term_on_the_device_matches_the_numpy_restatement(cls_name):
    """DoorLoss / TransportLoss with weight (pose, velocity, contact): the contact term runs on the GPU (smx_contact_distance_loss);
    value, particle seed and primitive-position adjoint against the numpy restatement (loss_door.py:46-56, loss_transport.py:54-70),
    after a re-sort (storage order != particle id) and with the primitive far enough that the max(.., 0) is inactive."""
    import scenes
    from harness import Pair
    from softmac_b200.engine import losses as L
    rng = np.random.default_rng(17)
    n = 1501                                            # odd: the last particle belongs to no group of TransportLoss
    pair = Pair(n, tables=[scenes.sphere_table()], prim_params=[(0.5, 666.)], max_steps=6, sort_every=1)
    s13 = np.concatenate([[0.5, 0.75, 0.45], [0.95, 0.1, 0.2, 0.1], [0.1, -0.2, 0.05], [0.3, 0.0, -0.4]])
    pair.prims[0].set_all_states(0, s13, f_end=6)
    pair.gpu.reset(scenes.blob_state(n, rng))
    pair.gpu.substep(0); pair.gpu.substep(1)
    f = 2
    cls = getattr(L, cls_name)
    loss = cls(dict(weight=(0.5, 0.25, 3.0)), pair.gpu)
    loss.initialize()
    if cls_name == "TransportLoss":
        loss.set_target([0.4, 0.5, 0.6])
    assert loss.device_contact
    pair.gpu.clear_all_gradients()
    info = loss.compute_loss(f)
    x, st = pair.gpu.get_x(f), pair.prims[0].get_all_states(f)
    val, gx, gp = loss.contact(x, st[:3])               # numpy restatement
    assert val > 0 and np.isclose(info["contact_loss"], 3.0 * val, rtol=1e-12)
    seed = pair.gpu.get_state_grad(f)
    assert rel_l2(seed[:, :3], 3.0 * gx) <= 1e-6 and np.abs(seed[:, 3:]).max() == 0
    assert np.count_nonzero(np.abs(seed[:, :3]).sum(1)) == loss.n_groups
    g13 = pair.prims[0].get_all_states_grad(f)
    pv, gpose = loss.pose(st)
    vv, gvel = L._RigidTerms.velocity(st, 0.0)
    exp = 0.5 * gpose + 0.25 * gvel
    exp[:3] += 3.0 * gp
    assert np.allclose(g13, exp, rtol=1e-6, atol=1e-9), (g13, exp)
    assert np.isclose(info["loss"], 0.5 * pv + 0.25 * vv + 3.0 * val)
    # inside the 0.1 ball around the primitive the term and its gradient vanish (max(.., 0) and the 0 < m test)
    near = s13.copy(); near[:3] = x[7]
    pair.prims[0].set_all_states(f, near)
    pair.gpu.clear_all_gradients()
    loss.clear()
    loss.weight = (0.0, 0.0, 3.0)
    info = loss.compute_loss(f)
    val2 = loss.contact(x, pair.prims[0].get_all_states(f)[:3])[0]
    assert np.isclose(info["contact_loss"], 3.0 * val2, rtol=1e-12, atol=1e-300)
    if cls_name == "DoorLoss":
        assert info["contact_loss"] == 0.0 and np.abs(pair.gpu.get_state_grad(f)).max() == 0
