"""Container of primitives -- mirror of ``softmac/engine/primitive/primitives.py:Primitives``."""
from .primitive_base import Primitive


class Primitives:
    def __init__(self, cfgs=(), max_timesteps=2048, rigid_velocity_control=False, primitives=None):
        self.primitives = list(primitives) if primitives is not None else []
        self.urdfs = list(cfgs)
        self.max_timesteps = max_timesteps
        self.rigid_velocity_control = rigid_velocity_control

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
