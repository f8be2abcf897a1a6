"""Mesh -> SDF / normal tables on the GPU: the set-up step of ``softmac/engine/primitive/mesh.py:136-241``
(Mesh.preprocess_sdf / task / trimesh2sdf) without trimesh.  Grid specification is the reference's (mesh.py:167-176,
190-192, 232-233); the distance / nearest-face queries run in ``smx_build_sdf_table`` (softmac_b200/csrc/smx_sdf.cuh).
"""
import hashlib
import os

import numpy as np

from ..._capi import lib, check, as_d, d_ptr, ip


def load_obj(path):
    """'v' and 'f' records of a Wavefront OBJ (polygons fan-triangulated, v/vt/vn index triples accepted)."""
    V, Fc = [], []
    with open(path) as fh:
        for line in fh:
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
    return np.array(V, dtype=np.float64), np.array(Fc, dtype=np.int32)


def grid_spec(vertices):
    """(res, lower, upper, dx): dx = min(0.01, extent/80), margin = max(3 dx, 0.01), samples at cell centres of the padded box."""
    V = np.asarray(vertices, dtype=np.float64)
    b0, b1 = V.min(0), V.max(0)
    dx = min(0.01, float(np.max(b1 - b0)) / 80)
    margin = max(dx * 3, 0.01)
    res = np.ceil((b1 - b0 + margin * 2) / dx).astype(int)
    lower = (b0 + b1) / 2 - res * dx / 2.0 + dx / 2.0
    upper = lower + (res - 1) * dx
    return res, lower, upper, dx


def build_sdf(vertices, faces, device=0):
    """-> dict(sdf, normal, position=(lower, upper), dx, res): the "sdf" entry of the reference's cache pickle (mesh.py:235-241)."""
    V = as_d(vertices).reshape(-1, 3)
    Fc = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
    res, lower, upper, dx = grid_spec(V)
    res32 = np.ascontiguousarray(res, dtype=np.int32)
    sdf = np.zeros(tuple(res))
    nrm = np.zeros(tuple(res) + (3,))
    lo = as_d(lower)
    check(lib().smx_build_sdf_table(d_ptr(V), len(V), Fc.ctypes.data_as(ip), len(Fc), res32.ctypes.data_as(ip), d_ptr(lo), float(dx),
                                    d_ptr(sdf), d_ptr(nrm), int(device)))
    return dict(sdf=sdf, normal=nrm, position=(lower, upper), dx=np.ones(3) * dx, res=res)


def cached_sdf(mesh_path, cache_dir=None, device=0):
    """Build (or load) the tables of an OBJ mesh.  Cache: <cache_dir>/<sha256 of "smx-v1" + vertices + faces>.npz."""
    V, Fc = load_obj(mesh_path)
    h = hashlib.sha256()
    h.update(b"smx-v1"); h.update(V.tobytes()); h.update(Fc.tobytes())
    cache_dir = cache_dir or os.path.join(os.path.expanduser("~"), ".cache", "softmac_b200", "sdf")
    path = os.path.join(cache_dir, h.hexdigest() + ".npz")
    if os.path.exists(path):
        z = np.load(path)
        return dict(sdf=z["sdf"], normal=z["normal"], position=(z["lower"], z["upper"]), dx=z["dx"], res=z["res"]), (V, Fc)
    sdf = build_sdf(V, Fc, device=device)
    try:
        os.makedirs(cache_dir, exist_ok=True)
        np.savez_compressed(path, sdf=sdf["sdf"], normal=sdf["normal"], lower=sdf["position"][0], upper=sdf["position"][1], dx=sdf["dx"], res=sdf["res"])
    except OSError:
        pass
    return sdf, (V, Fc)
