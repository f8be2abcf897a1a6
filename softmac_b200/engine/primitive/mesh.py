"""``Mesh`` primitive: a rigid body described by an SDF table and a normal table.

Mirror of ``softmac/engine/primitive/mesh.py:Mesh`` for the run-time part (table lookup happens on the
device, smx_contact.cuh).  Tables come either from the reference's cached pickle
(mesh.py:148-163: ``{"signature", "sdf": {"sdf","normal","position","dx","res"}, "meshes"}``) or from
arrays passed in directly.  Building tables from a triangle mesh (mesh.py:167-241, trimesh) is a
set-up-time step outside the hot path.
"""
import pickle

import numpy as np

from .primitive_base import Primitive


class Mesh(Primitive):
    def __init__(self, mesh_path=None, color=None, sdf=None, cache_dir=None, device=0, **kwargs):
        super().__init__(**kwargs)
        self.mesh_path, self.color = mesh_path, color
        self.urdf_path = self.cfg.get("urdf_path", "")
        self.mesh_rest = None
        if sdf is not None:
            self.load_sdf(sdf)
        elif mesh_path is not None:
            # Mesh.preprocess_sdf (mesh.py:136-165): tables built on the GPU from the OBJ (or loaded from the cache)
            from .sdf_builder import cached_sdf
            tables, self.mesh_rest = cached_sdf(str(mesh_path), cache_dir=cache_dir, device=device)
            self.load_sdf(tables)

    def load_sdf(self, sdf):
        """sdf: dict with keys sdf, normal, position=(lower, upper), dx (the "sdf" entry of the pickle)."""
        self.sdf_table = np.ascontiguousarray(sdf["sdf"], dtype=np.float64)
        self.normal_table = np.ascontiguousarray(sdf["normal"], dtype=np.float64)
        pos = sdf["position"] if "position" in sdf else (sdf["lower"], sdf["upper"])
        self.sdf_lower = np.asarray(pos[0], dtype=np.float64)
        self.sdf_upper = np.asarray(pos[1], dtype=np.float64)
        dx = sdf["dx"]
        self.sdf_dx = float(dx[0] if np.ndim(dx) else dx)
        self.inv_sdf_dx = 1.0 / self.sdf_dx
        self.sdf_res = list(self.sdf_table.shape)

    @classmethod
    def from_pickle(cls, path, **kwargs):
        with open(path, "rb") as f:
            blob = pickle.load(f)
        m = cls(mesh_path=path, sdf=blob["sdf"], **kwargs)
        m.meshes = blob.get("meshes")
        return m
