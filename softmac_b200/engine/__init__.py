from .mpm_simulator import MPMSimulator  # noqa: F401
from .primitive import Primitive, Primitives, Mesh  # noqa: F401
