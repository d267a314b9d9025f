"""frave_b200 — B200-native (sm_100a) fractal transform + quantization path of the frave codec.

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/fri_cuda.h), capi
(ctypes binding of that ABI), stages (mirror of the reference's stage interface) and sharding
(frame partitioning across the GPUs of one node; no collective).
"""
from . import capi, sharding, stages  # noqa: F401

__all__ = ["capi", "sharding", "stages"]
