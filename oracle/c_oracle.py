"""ctypes loader for oracle/libfri_oracle.so (the C restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs — never by frave_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfri_oracle.so")


class Raster(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("channels", C.c_uint32),
        ("sample_bytes", C.c_uint32),
        ("data", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fri_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfri_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        p32 = C.POINTER(C.c_int32)
        pu8 = C.POINTER(C.c_uint8)
        L.fri_oracle_fractal_divide.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(p32), C.POINTER(C.c_size_t)]
        L.fri_oracle_fractal_divide.restype = C.c_int
        L.fri_oracle_from_raster.argtypes = [C.POINTER(Raster), C.c_int, C.POINTER(p32), C.POINTER(p32),
                                             C.POINTER(pu8), C.POINTER(C.c_size_t)]
        L.fri_oracle_from_raster.restype = C.c_int
        L.fri_oracle_extract_tiles.argtypes = [C.POINTER(Raster), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                               C.c_void_p, C.c_int]
        L.fri_oracle_extract_tiles.restype = None
        L.fri_oracle_quantize.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_void_p, C.c_int]
        L.fri_oracle_quantize.restype = None
        L.fri_oracle_extract_values.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32,
                                                C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
        L.fri_oracle_extract_values.restype = None
        L.fri_oracle_encode_tiles.argtypes = [C.POINTER(Raster), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                              C.c_void_p, C.c_int]
        L.fri_oracle_encode_tiles.restype = None
        L.fri_oracle_decode_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32,
                                              C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.fri_oracle_decode_tiles.restype = None
        L.fri_oracle_image_positions.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_void_p]
        L.fri_oracle_image_positions.restype = None
        L.fri_oracle_prev_power_two.argtypes = [C.c_size_t]
        L.fri_oracle_prev_power_two.restype = C.c_size_t
        L.fri_oracle_free.argtypes = [C.c_void_p]
        L.fri_oracle_free.restype = None
        _lib = L
    return _lib


def _raster(img: np.ndarray) -> tuple[Raster, np.ndarray]:
    assert img.ndim == 3 and img.dtype in (np.uint8, np.uint16)
    img = np.ascontiguousarray(img)
    h, w, c = img.shape
    return Raster(w, h, c, img.dtype.itemsize, img.ctypes.data), img


def fractal_divide(width: int, height: int, depth: int = 9) -> np.ndarray:
    """All built tiles (in-bounds + fringe), wavelet_transform.rs:450-484."""
    L = lib()
    ptr = C.POINTER(C.c_int32)()
    n = C.c_size_t()
    if L.fri_oracle_fractal_divide(width, height, depth, C.byref(ptr), C.byref(n)):
        raise MemoryError
    out = np.ctypeslib.as_array(ptr, shape=(n.value, 2)).copy() if n.value else np.zeros((0, 2), np.int32)
    L.fri_oracle_free(ptr)
    return out


def from_raster(img: np.ndarray, depth: int = 9):
    """(centers[n,2], coef[n,C,2^depth] i32, some[n,C,2^depth] bool), tiles sorted by (im, re)."""
    L = lib()
    r, img = _raster(img)
    pc, pk, ps = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    if L.fri_oracle_from_raster(C.byref(r), depth, C.byref(pc), C.byref(pk), C.byref(ps), C.byref(n)):
        raise MemoryError
    c = img.shape[2]
    nn = n.value
    if nn:
        centers = np.ctypeslib.as_array(pc, shape=(nn, 2)).copy()
        coef = np.ctypeslib.as_array(pk, shape=(nn, c, 1 << depth)).copy()
        some = np.ctypeslib.as_array(ps, shape=(nn, c, 1 << depth)).copy().astype(bool)
    else:
        centers = np.zeros((0, 2), np.int32)
        coef = np.zeros((0, c, 1 << depth), np.int32)
        some = np.zeros((0, c, 1 << depth), bool)
    for p in (pc, pk, ps):
        L.fri_oracle_free(p)
    return centers, coef, some


def extract_tiles(img: np.ndarray, centers: np.ndarray, depth: int = 9, nthreads: int = 1, want_some: bool = True):
    """Forward transform for a given tile list (no BFS / retain)."""
    L = lib()
    r, img = _raster(img)
    centers = np.ascontiguousarray(centers, dtype=np.int32)
    n, c = len(centers), img.shape[2]
    coef = np.empty((n, c, 1 << depth), np.int32)
    some = np.empty((n, c, 1 << depth), np.uint8) if want_some else None
    L.fri_oracle_extract_tiles(C.byref(r), depth, centers.ctypes.data, n, coef.ctypes.data,
                               some.ctypes.data if want_some else None, nthreads)
    return coef, (some.astype(bool) if want_some else None)


def quantize(coef: np.ndarray, some: np.ndarray | None, q, depth: int = 9, multiply: bool = False) -> np.ndarray:
    L = lib()
    out = np.ascontiguousarray(coef, dtype=np.int32).copy()
    n, c, _ = out.shape
    qa = np.ascontiguousarray(q, dtype=np.int32)
    assert qa.shape == (32,)
    s8 = np.ascontiguousarray(some, dtype=np.uint8) if some is not None else None
    L.fri_oracle_quantize(out.ctypes.data, s8.ctypes.data if s8 is not None else None, n, c, depth,
                          qa.ctypes.data, int(multiply))
    return out


def extract_values(centers: np.ndarray, coef: np.ndarray, some: np.ndarray | None, h: int, w: int, depth: int = 9,
                   dtype=np.uint8, nthreads: int = 1) -> np.ndarray:
    L = lib()
    centers = np.ascontiguousarray(centers, dtype=np.int32)
    coef = np.ascontiguousarray(coef, dtype=np.int32)
    n, c, _ = coef.shape
    s8 = np.ascontiguousarray(some, dtype=np.uint8) if some is not None else None
    out = np.zeros((h, w, c), dtype=dtype)
    L.fri_oracle_extract_values(centers.ctypes.data, coef.ctypes.data, s8.ctypes.data if s8 is not None else None,
                                n, depth, w, h, c, out.dtype.itemsize, out.ctypes.data, nthreads)
    return out


def encode_tiles(img: np.ndarray, centers: np.ndarray, q, depth: int = 9, nthreads: int = 1,
                 out: np.ndarray | None = None) -> np.ndarray:
    """Timed-baseline driver: forward transform + quantization over a tile list (None -> 0)."""
    L = lib()
    r, img = _raster(img)
    centers = np.ascontiguousarray(centers, dtype=np.int32)
    qa = np.ascontiguousarray(q, dtype=np.int32)
    n, c = len(centers), img.shape[2]
    if out is None:
        out = np.empty((n, c, 1 << depth), np.int32)
    L.fri_oracle_encode_tiles(C.byref(r), depth, centers.ctypes.data, n, qa.ctypes.data, out.ctypes.data, nthreads)
    return out


def decode_tiles(centers: np.ndarray, coef: np.ndarray, some: np.ndarray | None, q, h: int, w: int, depth: int = 9,
                 dtype=np.uint8, nthreads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
    """Timed-baseline driver: dividing dequantization + inverse transform over a tile list."""
    L = lib()
    centers = np.ascontiguousarray(centers, dtype=np.int32)
    coef = np.ascontiguousarray(coef, dtype=np.int32)
    qa = np.ascontiguousarray(q, dtype=np.int32)
    n, c, _ = coef.shape
    s8 = np.ascontiguousarray(some, dtype=np.uint8) if some is not None else None
    if out is None:
        out = np.zeros((h, w, c), dtype=dtype)
    L.fri_oracle_decode_tiles(centers.ctypes.data, coef.ctypes.data, s8.ctypes.data if s8 is not None else None, n,
                              depth, w, h, c, out.dtype.itemsize, qa.ctypes.data, out.ctypes.data, nthreads)
    return out


def fractal_new_cost(centers: np.ndarray, channels: int, depth: int = 9) -> int:
    """Cost model (not a restatement): the containers Fractal::new builds per tile — see fri_oracle.c."""
    L = lib()
    L.fri_oracle_fractal_new_cost.restype = C.c_uint64
    L.fri_oracle_fractal_new_cost.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint32]
    centers = np.ascontiguousarray(centers, dtype=np.int32)
    return int(L.fri_oracle_fractal_new_cost(depth, centers.ctypes.data, len(centers), channels))
