"""mvregfus_b200 — B200-native (sm_100a) view-fusion hot path of MVRegFus.

Hand-written CUDA kernels behind a C-ABI (``include/mvregfus_b200.h``), wrapped by drop-in
replacements for the ``mvregfus.multiview`` entry points (``mvregfus_b200.multiview``).
There is no CPU fallback: importing the compute modules without the built extension raises.
"""
from .image_array import ImageArray  # noqa: F401

__version__ = "0.1.0"
