"""Builds frave_b200/libfri_cuda.so in-tree with nvcc for sm_100a (no JIT cache, no torch).

    python -m frave_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libfri_cuda.so")
SOURCES = ["fri_api.cu", "fri_kernels.cu", "fri_predict.cu", "fri_plan.cpp", "fri_order.cpp", "fri_codec.cpp"]
HEADERS = ["fri_geometry.h", "fri_plan.h", "fri_kernels.cuh", "fri_codec.h", os.path.join("..", "..", "include", "fri_cuda.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-ffp-contract=off",  # no FMA contraction in host f32 code (bit parity with Rust)
    "--shared",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfri_cuda cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str | None = None, extra: list[str] | None = None) -> str:
    """Builds the library.  `out`/`extra` build a tuning variant (extra nvcc flags, e.g. -DFRI_...)
    next to the product library; capi loads it when FRI_CUDA_LIB points at it."""
    target = out or LIB_PATH
    if not force and out is None and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, *(extra or [])]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", target, *[os.path.join(_CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return target


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a not in ("--force", "--verbose")]
    out = None
    if args and args[0] == "--variant":  # python -m frave_b200.build --variant <tag> -DFOO=1 ...
        out = os.path.join(_HERE, f"libfri_cuda_{args[1]}.so")
        args = args[2:]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=out, extra=args))
