"""numpy restatement of the SDF-table construction of ``softmac/engine/primitive/mesh.py:167-241`` (Mesh.task +
trimesh2sdf).  TEST INFRASTRUCTURE (oracle for softmac_b200's CUDA table builder).

Third-party arithmetic: ``trimesh.proximity.ProximityQuery.signed_distance / on_surface`` (trimesh==3.21.5,
requirements.txt:8) -- exact point-to-triangle distance, sign from containment, nearest triangle id.  Restated here
with the Ericson closest-point-on-triangle routine and the generalized winding number for the sign.  Pinned against the
two cached tables the reference ships (assets/gripper/6895..., assets/door/e7ab...): see tests/test_sdf_builder.py.
Semantics: sdf = signed distance at lower + (i,j,k)*dx, negative inside; normal = unit normal of
the nearest triangle divided by (1 + 1e-8).
"""
import numpy as np


def grid_spec(vertices):
    """mesh.py:167-176 + 190-192, 232-233: (res, lower, upper, dx) with lower/upper at the first/last sample."""
    V = np.asarray(vertices, dtype=np.float64)
    b0, b1 = V.min(0), V.max(0)
    length = np.max(b1 - b0)
    dx = min(0.01, length / 80)
    margin = max(dx * 3, 0.01)
    center = (b0 + b1) / 2
    res = np.ceil((b1 - b0 + margin * 2) / dx).astype(int)
    lower = center - res * dx / 2.0
    lower = lower + dx / 2.0
    upper = lower + (res - 1) * dx
    return res, lower, upper, dx


def closest_point_on_triangles(P, A, B, C):
    """Closest points of every P[i] to every triangle (A[j], B[j], C[j]) -> (n, m, 3).  Ericson, Real-Time Collision
    Detection 5.1.5, vectorised."""
    P = P[:, None, :]
    ab, ac, ap = B - A, C - A, P - A
    d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
    bp = P - B
    d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
    cp = P - C
    d5, d6 = (ab * cp).sum(-1), (ac * cp).sum(-1)
    vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
    out = np.empty(np.broadcast_shapes(P.shape, A.shape))
    done = np.zeros(out.shape[:-1], dtype=bool)

    def put(mask, val):
        m = mask & ~done
        out[m] = np.broadcast_to(val, out.shape)[m]
        done[m] = True

    with np.errstate(divide="ignore", invalid="ignore"):
        put((d1 <= 0) & (d2 <= 0), A)
        put((d3 >= 0) & (d4 <= d3), B)
        put((vc <= 0) & (d1 >= 0) & (d3 <= 0), A + (d1 / (d1 - d3))[..., None] * ab)
        put((d6 >= 0) & (d5 <= d6), C)
        put((vb <= 0) & (d2 >= 0) & (d6 <= 0), A + (d2 / (d2 - d6))[..., None] * ac)
        put((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), B + ((d4 - d3) / ((d4 - d3) + (d5 - d6)))[..., None] * (C - B))
        den = 1.0 / (va + vb + vc)
        put(np.ones_like(done), A + (vb * den)[..., None] * ab + (vc * den)[..., None] * ac)
    return out


def winding_number(P, A, B, C):
    """Generalised winding number (Van Oosterom & Strackee solid angles); ~1 inside a closed mesh, ~0 outside."""
    a, b, c = A - P[:, None, :], B - P[:, None, :], C - P[:, None, :]
    la, lb, lc = np.linalg.norm(a, axis=-1), np.linalg.norm(b, axis=-1), np.linalg.norm(c, axis=-1)
    num = (a * np.cross(b, c)).sum(-1)
    den = la * lb * lc + (a * b).sum(-1) * lc + (b * c).sum(-1) * la + (c * a).sum(-1) * lb
    return (2 * np.arctan2(num, den)).sum(-1) / (4 * np.pi)


def build_sdf(vertices, faces, chunk=4096):
    """-> dict(sdf, normal, position=(lower, upper), dx, res) in the layout of mesh.py:235-241."""
    V, Fc = np.asarray(vertices, dtype=np.float64), np.asarray(faces, dtype=np.int64)
    res, lower, upper, dx = grid_spec(V)
    ax = [lower[d] + np.arange(res[d]) * dx for d in range(3)]
    P = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    A, B, C = V[Fc[:, 0]], V[Fc[:, 1]], V[Fc[:, 2]]
    fn = np.cross(B - A, C - A)
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    sdf, nrm = np.empty(len(P)), np.empty((len(P), 3))
    for i in range(0, len(P), chunk):
        p = P[i:i + chunk]
        cp = closest_point_on_triangles(p, A[None], B[None], C[None])
        d2 = ((cp - p[:, None, :]) ** 2).sum(-1)
        tri = nearest_triangle(d2)
        dist = np.sqrt(d2[np.arange(len(p)), tri])
        inside = winding_number(p, A[None], B[None], C[None]) > 0.5
        sdf[i:i + chunk] = np.where(inside, -dist, dist)
        nrm[i:i + chunk] = fn[tri] / (1.0 + 1e-8)              # mesh.py:215 with |face normal| = 1
    return dict(sdf=sdf.reshape(res), normal=nrm.reshape(tuple(res) + (3,)), position=(lower, upper), dx=np.ones(3) * dx, res=res)


def nearest_triangle(d2, rel_tol=1e-9):
    """Index of the nearest triangle; among (numerically) equidistant ones the LOWEST index."""
    dmin = d2.min(1, keepdims=True)
    cand = d2 <= dmin * (1 + rel_tol) + 1e-30
    return cand.argmax(1)


def load_obj(path):
    """Minimal OBJ reader: 'v' and 'f' records (polygons are fan-triangulated; v/vt/vn indices accepted)."""
    V, Fc = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            V.append([float(t[1]), float(t[2]), float(t[3])])
        elif t[0] == "f":
            idx = [int(s.split("/")[0]) for s in t[1:]]
            idx = [i - 1 if i > 0 else len(V) + i for i in idx]
            for k in range(1, len(idx) - 1):
                Fc.append([idx[0], idx[k], idx[k + 1]])
    return np.array(V, dtype=np.float64), np.array(Fc, dtype=np.int64)
