"""softmac_b200 -- B200-native (sm_100a) implementation of SoftMAC's differentiable MLS-MPM substep loop.

The package keeps the ``softmac/engine`` simulator API of the reference (``MPMSimulator``, ``Primitive``,
``Primitives``, ``TaichiEnv``) and replaces the Taichi kernels with hand-written CUDA behind the C ABI in
``include/softmac_b200.h`` (``softmac_b200/lib/libsoftmac_b200.so``).  There is no CPU fallback: importing
the engine works anywhere, creating a simulator requires the built library and a CUDA device.
"""
__version__ = "0.1.0"
