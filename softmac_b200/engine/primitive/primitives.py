"""Container of primitives -- mirror of ``softmac/engine/primitive/primitives.py:Primitives``."""
from .primitive_base import Primitive


import os
import xml.etree.ElementTree as ET

import numpy as np


class Primitives:
    def __init__(self, cfgs=(), max_timesteps=2048, rigid_velocity_control=False, primitives=None, cache_dir=None, device=0):
        """cfgs: the reference's cfg.PRIMITIVES entries (friction, urdf_path, enable_external_force): one ``Mesh`` per
        ``<collision><geometry><mesh>`` of each URDF, as in primitives.py:16-41.  `primitives`: ready-made objects instead."""
        self.primitives = list(primitives) if primitives is not None else []
        self.urdfs = list(cfgs)
        self.max_timesteps = max_timesteps
        self.rigid_velocity_control = rigid_velocity_control
        if primitives is None:
            from .mesh import Mesh
            for c in cfgs:
                paths, colors = self.load_info_from_urdf(c["urdf_path"] if isinstance(c, dict) else c.urdf_path)
                for mesh_path, color in zip(paths, colors):
                    self.primitives.append(Mesh(mesh_path, color=color, cfg=c, max_timesteps=max_timesteps,
                                                rigid_velocity_control=rigid_velocity_control, cache_dir=cache_dir, device=device))

    @staticmethod
    def load_info_from_urdf(urdf_path):
        """primitives.py:25-41: collision mesh files and visual colours of a URDF."""
        root = ET.parse(urdf_path).getroot()
        meshes = root.findall(".//collision/geometry/mesh")
        paths = [os.path.join(os.path.dirname(urdf_path), m.attrib.get("filename", "")) for m in meshes]
        colors = [np.array([float(v) for v in c.attrib.get("rgba", "1 1 1 1").split()]) for c in root.findall(".//visual/material/color")]
        colors += [np.ones(4)] * (len(paths) - len(colors))
        return paths, colors

    def append(self, primitive: Primitive):
        self.primitives.append(primitive)

    def set_softness(self, softness=666.):
        for i in self.primitives:
            i.softness[None] = softness

    def __getitem__(self, item):
        if isinstance(item, tuple):
            item = item[0]
        return self.primitives[item]

    def __len__(self):
        return len(self.primitives)

    def __iter__(self):
        return iter(self.primitives)

    def view(self, batch):
        """Per-rollout container for a batched handle: the same primitives addressed to batch `batch`."""
        return _PrimitivesView([p.view(batch) for p in self.primitives])

    def initialize(self):
        self.set_softness(666.)

    def reset(self):
        for i in self.primitives:
            i.reset()


class _PrimitivesView(list):
    def initialize(self):
        pass

    def reset(self):
        for p in self:
            p.reset()
