from .primitive_base import Primitive  # noqa: F401
from .mesh import Mesh  # noqa: F401
from .primitives import Primitives  # noqa: F401
