"""SDF table builder against GOLDEN VECTORS OF THE REFERENCE ITSELF: the two cached tables the
reference ships (gripper palm, door), produced by its own Mesh.trimesh2sdf (mesh.py:178-241, trimesh 3.21.5).
CPU: the numpy restatement (oracle/sdf_builder.py).  GPU: smx_build_sdf_table.
Distances must agree everywhere (up to documented trimesh artefacts on the door); the nearest-face normal must agree wherever the nearest face is unique, and be one
of the equidistant faces' normals elsewhere (trimesh breaks those ties by rounding noise)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_sdf_reference_tables.npz"))


def check_against_reference(name, built):
    from oracle import sdf_builder as sb
    V, Fc = G[name + "_V"], G[name + "_F"]
    assert np.array_equal(np.asarray(built["res"]), G[name + "_res"])
    assert np.allclose(built["position"][0], G[name + "_lower"], atol=1e-15) and np.allclose(built["position"][1], G[name + "_upper"], atol=1e-15)
    assert abs(float(np.ravel(built["dx"])[0]) - float(G[name + "_dx"])) < 1e-15
    ref_sdf = G[name + "_sdf"].astype(np.float64)                                         # golden is stored in fp32
    # |distance|: exact up to trimesh's own closest-point noise (16 door samples are off by up to 1.6e-6 in the reference table)
    assert np.abs(np.abs(built["sdf"]) - np.abs(ref_sdf)).max() < 5e-6
    assert np.median(np.abs(np.abs(built["sdf"]) - np.abs(ref_sdf))) < 2e-9
    # sign: identical on the palm; on the door the reference's ray-parity containment test misfires on 209 of 97,524 samples
    # next to the handle (they are interior: generalised winding number 1), the only place where the tables differ in sign
    flips = (np.sign(built["sdf"]) != np.sign(ref_sdf)) & (np.abs(ref_sdf) > 1e-7)        # samples on the surface have no sign
    assert flips.mean() <= (0.0 if name == "palm" else 0.003), flips.sum()
    assert np.all(built["sdf"][flips] < 0)
    ref_n = G[name + "_normal"].astype(np.float64).reshape(-1, 3)
    got_n = (built["normal"] * (1 + 1e-8)).reshape(-1, 3)
    assert np.allclose(np.linalg.norm(got_n, axis=1), 1.0, atol=1e-9)
    # candidate (equidistant) faces of every sample
    res, lower, upper, dx = sb.grid_spec(V)
    ax = [lower[d] + np.arange(res[d]) * dx for d in range(3)]
    P = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    A, B, C = V[Fc[:, 0]], V[Fc[:, 1]], V[Fc[:, 2]]
    fn = np.cross(B - A, C - A); fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    bad_unique = bad_tie = n_tie = 0
    for i in range(0, len(P), 8192):
        p = P[i:i + 8192]
        d2 = ((sb.closest_point_on_triangles(p, A[None], B[None], C[None]) - p[:, None, :]) ** 2).sum(-1)
        cand = d2 <= d2.min(1, keepdims=True) * (1 + 1e-9) + 1e-30
        ok_got = (cand & (np.abs(fn[None] - got_n[i:i + 8192, None, :]).max(-1) < 1e-6)).any(1)
        distinct = np.array([len({tuple(np.round(fn[j], 6)) for j in np.nonzero(c)[0]}) for c in cand])
        tie = distinct > 1
        n_tie += tie.sum()
        bad_tie += (~ok_got).sum()                                                      # must always be a candidate's normal
        bad_unique += (np.abs(got_n[i:i + 8192] - ref_n[i:i + 8192]).max(-1) > 1e-6)[~tie].sum()
    assert bad_tie == 0 and bad_unique == 0, (bad_tie, bad_unique)
    assert n_tie > 0        # the fixture does contain edge / corner samples


@pytest.mark.parametrize("name", ["palm", "door"])
def test_numpy_restatement_matches_reference_tables(name):
    from oracle import sdf_builder as sb
    check_against_reference(name, sb.build_sdf(G[name + "_V"], G[name + "_F"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["palm", "door"])
def test_cuda_builder_matches_reference_tables(name):
    from softmac_b200.engine.primitive.sdf_builder import build_sdf
    check_against_reference(name, build_sdf(G[name + "_V"], G[name + "_F"]))


def icosphere(r=0.05):
    t = (1 + 5 ** 0.5) / 2
    V = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], float)
    V = V / np.linalg.norm(V, axis=1, keepdims=True) * r
    Fc = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2],
                   [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int32)
    return V, Fc


@pytest.mark.gpu
def test_cuda_builder_matches_numpy_on_non_box_mesh():
    from oracle import sdf_builder as sb
    from softmac_b200.engine.primitive.sdf_builder import build_sdf
    V, Fc = icosphere()
    a, b = build_sdf(V, Fc), sb.build_sdf(V, Fc)
    assert np.abs(a["sdf"] - b["sdf"]).max() < 1e-12
    assert (a["sdf"] < 0).sum() > 0 and (np.abs(a["normal"] - b["normal"]).max(-1) > 1e-9).mean() < 0.02      # ties only
