"""ctypes binding of libfri_cuda.so — the same C ABI (include/fri_cuda.h) that the `libfri-cuda`
Rust crate binds.  There is no CPU fallback: if the shared library is missing it is built with
nvcc, and if that is impossible importing callers get a RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

FRI_OK = 0
FRI_E_INVALID = -1
FRI_E_CUDA = -2
FRI_E_NOMEM = -3
FRI_E_UNSUPPORTED = -4
FRI_DEQUANT_DIVIDE = 0
FRI_DEQUANT_MULTIPLY = 1
FRI_BASE_DEPTH = 9

# every symbol include/fri_cuda.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SYMBOLS = [
    ("fri_version", C.c_char_p, []),
    ("fri_last_error", C.c_char_p, []),
    ("fri_device_count", C.c_int, []),
    ("fri_plan_create", C.c_int, [C.POINTER(_P), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    ("fri_plan_destroy", None, [_P]),
    ("fri_plan_num_tiles", C.c_uint32, [_P]),
    ("fri_plan_num_built", C.c_uint32, [_P]),
    ("fri_plan_num_full_tiles", C.c_uint32, [_P]),
    ("fri_plan_coefs_per_frame", C.c_uint64, [_P]),
    ("fri_plan_pixels_covered", C.c_uint64, [_P]),
    ("fri_plan_launch_info", C.c_int, [_P, _P]),
    ("fri_plan_centers", C.c_int, [_P, _P]),
    ("fri_plan_masks", C.c_int, [_P, _P]),
    ("fri_encode_tq_device", C.c_int, [_P, _P, C.c_uint32, _P, _P, _P]),
    ("fri_decode_tq_device", C.c_int, [_P, _P, C.c_uint32, _P, C.c_int, _P, _P]),
    ("fri_encode_tq_device16", C.c_int, [_P, _P, C.c_uint32, _P, _P, _P]),
    ("fri_decode_tq_device16", C.c_int, [_P, _P, C.c_uint32, _P, C.c_int, _P, _P]),
    ("fri_encode_tq", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_decode_tq", C.c_int, [_P, _P, C.c_uint32, _P, C.c_int, _P]),
    ("fri_encode_tq16", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_decode_tq16", C.c_int, [_P, _P, C.c_uint32, _P, C.c_int, _P]),
    ("fri_plan_emission_count", C.c_uint64, [_P]),
    ("fri_plan_emission_order", C.c_int, [_P, _P]),
    ("fri_emit_device", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_encode_tq_emit", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_emit_device16", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_encode_tq_emit16", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_unemit_device", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_unemit_device16", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_decode_tq_emit", C.c_int, [_P, _P, C.c_uint32, _P, C.c_int, _P]),
    ("fri_decode_tq_emit16", C.c_int, [_P, _P, C.c_uint32, _P, C.c_int, _P]),
    ("fri_plan_emission_packed_bytes", C.c_uint64, [_P]),
    ("fri_plan_emission_packed_size", C.c_uint64, [_P, C.c_int]),
    ("fri_emit_device_packed", C.c_int, [_P, _P, C.c_uint32, C.c_int, _P, _P]),
    ("fri_encode_tq_emit_packed", C.c_int, [_P, _P, C.c_uint32, _P, C.c_int, _P]),
    ("fri_unemit_device_packed", C.c_int, [_P, _P, C.c_uint32, C.c_int, _P, _P]),
    ("fri_decode_tq_emit_packed", C.c_int, [_P, _P, C.c_uint32, C.c_int, _P, C.c_int, _P]),
    ("fri_emit_device10", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_encode_tq_emit10", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_unemit_device10", C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    ("fri_decode_tq_emit10", C.c_int, [_P, _P, C.c_uint32, _P, C.c_int, _P]),
    ("fri_predict_device", C.c_int, [_P, _P, C.c_uint32, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("fri_fit_parameters", C.c_int, [_P, _P, _P, _P]),
    ("fri_fit_device", C.c_int, [_P, _P, _P, _P, _P]),
    ("fri_plan_part", C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, _P]),
    ("fri_encode_tq_device_part", C.c_int, [_P, _P, _P, _P, C.c_uint32, C.c_uint32, _P]),
    ("fri_decode_tq_device_part", C.c_int, [_P, _P, _P, C.c_int, _P, C.c_uint32, C.c_uint32, _P]),
    ("fri_plan_groups_in_rows", C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P, _P]),
    ("fri_decode_tq_device_groups", C.c_int, [_P, _P, C.c_uint32, _P, C.c_int, _P, C.c_int32, C.c_uint32, C.c_uint32, _P]),
    ("fri_predict_host", C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("fri_frv_pack", C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    ("fri_frv_unpack", C.c_int, [_P, _P, C.c_size_t, _P]),
    ("fri_frv_info", C.c_int, [_P, C.c_size_t, _P, _P, _P]),
    ("fri_frv_encode", C.c_int, [_P, _P, _P, C.c_int, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    ("fri_frv_decode", C.c_int, [_P, _P, C.c_size_t, _P, C.c_int, _P]),
    ("fri_frv_free", None, [_P]),
    ("fri_host_alloc", C.c_int, [C.POINTER(_P), C.c_size_t]),
    ("fri_host_free", None, [_P]),
    ("fri_plan_last_launches", C.c_uint32, [_P]),
    ("fri_plan_set_bands", C.c_int, [_P, C.c_int]),
    ("fri_plan_set_async", C.c_int, [_P, C.c_int]),
    ("fri_plan_sync", C.c_int, [_P]),
    ("fri_plan_set_independent_calls", C.c_int, [_P, C.c_int]),
    ("fri_quant_divide", C.c_int32, [C.c_int32, C.c_int32]),
    ("fri_quant_divide_small", C.c_int32, [C.c_int32, C.c_int32]),
    ("fri_quant_divide_magic", C.c_int32, [C.c_int32, C.c_int32]),
]
SYMBOL_NAMES = [s[0] for s in _SYMBOLS]

_lib = None


class FriError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libfri_cuda error {code}: {message}")
        self.code = code


def lib_path() -> str:
    """The product library, or a tuning variant when FRI_CUDA_LIB names one (see build.py)."""
    return os.environ.get("FRI_CUDA_LIB") or _build.LIB_PATH


def lib() -> C.CDLL:
    """Loads (building first if necessary) libfri_cuda.so.  Never falls back to a CPU path."""
    global _lib
    if _lib is None:
        if not os.environ.get("FRI_CUDA_LIB") and not os.path.exists(_build.LIB_PATH):
            _build.build()
        L = C.CDLL(lib_path())
        for name, res, args in _SYMBOLS:
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != FRI_OK:
        raise FriError(rc, lib().fri_last_error().decode("utf-8", "replace"))


def version() -> str:
    return lib().fri_version().decode()


def device_count() -> int:
    return int(lib().fri_device_count())


def quant_divide(value: int, q: int) -> int:
    """value / q with the kernels' multiply-high division routine (host evaluation, for tests)."""
    return int(lib().fri_quant_divide(int(value), int(q)))


def frv_info(data: bytes) -> tuple[int, int, int]:
    """(width, height, channels) of a `frif` container."""
    buf = np.frombuffer(data, dtype=np.uint8)
    w, h, c = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    _check(lib().fri_frv_info(buf.ctypes.data, len(data), C.addressof(w), C.addressof(h), C.addressof(c)))
    return int(w.value), int(h.value), int(c.value)


def pack_bits(values: np.ndarray, bits: int = 10) -> np.ndarray:
    """Host restatement of the packed transport for tests: int stream [..., n] -> uint8 [..., 8 * bits * ceil(n / 64)]
    (zig-zag pack_signed of utils.rs:34-40 saturated to `bits` bits, 64 symbols per block, little-endian bit order,
    zero padding)."""
    v = np.asarray(values).astype(np.int64)
    n = v.shape[-1]
    pad = (-n) % 64
    v = np.clip(v, -(1 << (bits - 1)), (1 << (bits - 1)) - 1)
    sym = np.where(v >= 0, 2 * v, -2 * v - 1).astype(np.uint8 if False else np.uint64)
    sym = np.concatenate([sym, np.zeros(v.shape[:-1] + (pad,), np.uint64)], axis=-1)
    # symbol i of a 64-symbol block occupies bits [bits * i, bits * i + bits) of the block's 8 * bits bytes
    b = ((sym[..., None] >> np.arange(bits, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8)  # [..., n_pad, bits], LSB first
    b = b.reshape(v.shape[:-1] + (-1,))
    return np.packbits(b, axis=-1, bitorder="little")


def unpack_bits(packed: np.ndarray, n: int, bits: int = 10) -> np.ndarray:
    """Inverse of pack_bits: uint8 [..., 8 * bits * ceil(n / 64)] -> int32 [..., n] (unpack_signed, utils.rs:42-48)."""
    p = np.asarray(packed, dtype=np.uint8)
    b = np.unpackbits(p, axis=-1, bitorder="little").reshape(p.shape[:-1] + (-1, bits)).astype(np.int64)
    sym = (b << np.arange(bits, dtype=np.int64)).sum(axis=-1)[..., :n]
    return np.where(sym % 2 == 0, sym // 2, -((sym + 1) // 2)).astype(np.int32)


def pack10(values: np.ndarray) -> np.ndarray:
    return pack_bits(values, 10)


def unpack10(packed: np.ndarray, n: int) -> np.ndarray:
    return unpack_bits(packed, n, 10)


def _q_array(q):
    if q is None:
        return None, None
    qa = np.ascontiguousarray(q, dtype=np.int32)
    if qa.shape != (32,):
        raise ValueError("the quantization matrix has 32 entries (quantization.rs:3-5)")
    return qa, qa.ctypes.data


class PinnedBuffer:
    """Page-locked host memory from fri_host_alloc, viewed as a numpy array."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        ptr = _P()
        _check(lib().fri_host_alloc(C.byref(ptr), nbytes))
        self.ptr = ptr.value
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=nbytes // self.dtype.itemsize).reshape(self.shape)

    def free(self) -> None:
        if self.ptr:
            self.array = None
            lib().fri_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Plan:
    """fri_plan: lattice + launch geometry for one (width, height, channels, depth, sample size).

    device >= 0 uploads the tables to that CUDA device; device = -1 makes a host-only plan on
    which only the metadata queries work (compute calls raise FriError(FRI_E_CUDA)).
    """

    def __init__(self, width: int, height: int, channels: int, depth: int = FRI_BASE_DEPTH, sample_bytes: int = 1,
                 device: int = 0):
        self._h = _P()
        _check(lib().fri_plan_create(C.byref(self._h), device, width, height, channels, depth, sample_bytes))
        self.width, self.height, self.channels = int(width), int(height), int(channels)
        self.depth, self.sample_bytes, self.device = int(depth), int(sample_bytes), int(device)
        L = lib()
        self.n_tiles = int(L.fri_plan_num_tiles(self._h))
        self.n_built = int(L.fri_plan_num_built(self._h))
        self.n_full = int(L.fri_plan_num_full_tiles(self._h))
        self.coefs_per_frame = int(L.fri_plan_coefs_per_frame(self._h))
        self.pixels_covered = int(L.fri_plan_pixels_covered(self._h))
        self.pixel_dtype = np.uint8 if sample_bytes == 1 else np.uint16
        self.frame_shape = (self.height, self.width, self.channels)
        self.coef_shape = (self.n_tiles, self.channels, 1 << self.depth)

    def close(self) -> None:
        if self._h:
            lib().fri_plan_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- metadata -------------------------------------------------------------------------
    def centers(self) -> np.ndarray:
        out = np.empty((self.n_tiles, 2), np.int32)
        _check(lib().fri_plan_centers(self._h, out.ctypes.data))
        return out

    def mask_words(self) -> np.ndarray:
        out = np.empty((self.n_tiles, (1 << self.depth) // 32), np.uint32)
        _check(lib().fri_plan_masks(self._h, out.ctypes.data))
        return out

    def masks(self) -> np.ndarray:
        """bool [n_tiles, 2^depth]: coefficient i of tile t is `Some` in the reference."""
        w = self.mask_words()
        bits = np.unpackbits(w.view(np.uint8), axis=1, bitorder="little")
        return bits.astype(bool)

    def launch_info(self) -> dict:
        info = (C.c_int32 * 16)()
        _check(lib().fri_plan_launch_info(self._h, C.addressof(info)))
        keys = ["group_a", "group_b", "region_w", "region_h", "smem_pitch", "smem_bytes", "n_groups", "n_base_tiles",
                "threads", "chunks_per_row", "depth", "sub_bits", "chunks_full", "chunks_owned"]
        return {k: int(info[i]) for i, k in enumerate(keys)}

    def set_bands(self, bands: int) -> None:
        """Bands per frame of the host-buffer entry points (0 = automatic; 1 for concurrent callers)."""
        _check(lib().fri_plan_set_bands(self._h, int(bands)))

    def set_async(self, on: bool) -> None:
        """Host-buffer calls return after enqueueing (pinned buffers only); sync() waits for them."""
        _check(lib().fri_plan_set_async(self._h, 1 if on else 0))

    def sync(self) -> None:
        _check(lib().fri_plan_sync(self._h))

    def set_independent_calls(self, on: bool) -> None:
        """Promise that consecutive *_device calls on one stream touch disjoint buffers (fri_plan_set_independent_calls)."""
        _check(lib().fri_plan_set_independent_calls(self._h, 1 if on else 0))

    @property
    def last_launches(self) -> int:
        return int(lib().fri_plan_last_launches(self._h))

    # ---- host-buffer entry points -----------------------------------------------------------
    def _frames(self, pixels: np.ndarray) -> tuple[np.ndarray, int]:
        px = np.asarray(pixels)
        if px.dtype != self.pixel_dtype:
            raise ValueError(f"pixels must be {np.dtype(self.pixel_dtype).name}")
        if px.shape == self.frame_shape:
            px = px[None]
        if px.ndim != 4 or px.shape[1:] != self.frame_shape:
            raise ValueError(f"pixels must have shape [F]{self.frame_shape}")
        return np.ascontiguousarray(px), px.shape[0]

    def encode(self, pixels: np.ndarray, q=None, out: np.ndarray | None = None, dtype=np.int32) -> np.ndarray:
        """HWC pixels [F, H, W, C] -> quantized coefficients [F, n_tiles, C, 2^depth]: int32, or int16
        through the 16-bit transport (fri_encode_tq16; `out.dtype` decides when `out` is given)."""
        px, n = self._frames(pixels)
        if out is None:
            out = np.empty((n,) + self.coef_shape, dtype)
        assert out.dtype in (np.int32, np.int16) and out.flags.c_contiguous and out.size == n * self.coefs_per_frame
        qa, qp = _q_array(q)
        fn = lib().fri_encode_tq16 if out.dtype == np.int16 else lib().fri_encode_tq
        _check(fn(self._h, px.ctypes.data, n, qp, out.ctypes.data))
        return out

    def decode(self, coefs: np.ndarray, q=None, multiply: bool = False, out: np.ndarray | None = None) -> np.ndarray:
        """Quantized coefficients [F, n_tiles, C, 2^depth] -> HWC pixels [F, H, W, C].  int16 input
        takes the 16-bit transport (fri_decode_tq16), anything else is passed as int32."""
        half = np.asarray(coefs).dtype == np.int16
        cf = np.ascontiguousarray(coefs, dtype=np.int16 if half else np.int32)
        if cf.shape == self.coef_shape:
            cf = cf[None]
        if cf.ndim != 4 or cf.shape[1:] != self.coef_shape:
            raise ValueError(f"coefs must have shape [F]{self.coef_shape}")
        n = cf.shape[0]
        if out is None:
            out = np.empty((n,) + self.frame_shape, self.pixel_dtype)
        assert out.dtype == self.pixel_dtype and out.flags.c_contiguous and out.shape[1:] == self.frame_shape
        qa, qp = _q_array(q)
        mode = FRI_DEQUANT_MULTIPLY if multiply else FRI_DEQUANT_DIVIDE
        fn = lib().fri_decode_tq16 if half else lib().fri_decode_tq
        _check(fn(self._h, cf.ctypes.data, n, qp, mode, out.ctypes.data))
        return out

    # ---- emission order (depth 9) -----------------------------------------------------------------
    def emission_order(self) -> np.ndarray:
        """uint32 [n_tiles * 512]: tile_index * 512 + coefficient_index in the order the reference's
        entropy coder consumes one channel (None slots included).  Raises FriError(FRI_E_UNSUPPORTED)
        where the reference's own scan asserts."""
        out = np.empty(self.n_tiles << self.depth, np.uint32)
        _check(lib().fri_plan_emission_order(self._h, out.ctypes.data))
        return out

    def emission_count(self) -> int:
        n = int(lib().fri_plan_emission_count(self._h))
        if n == 0 and self.n_tiles:
            _check(lib().fri_plan_emission_order(self._h, np.empty(self.n_tiles << self.depth, np.uint32).ctypes.data))
        return n

    def emission_packed_bytes(self) -> int:
        """Bytes of one channel's stream in the 10-bit packed transport (80 bytes per 64 symbols)."""
        self.emission_count()
        return int(lib().fri_plan_emission_packed_bytes(self._h))

    def emission_packed_size(self, bits: int) -> int:
        """Bytes of one channel's stream packed at `bits` (9 or 10) bits per symbol."""
        self.emission_count()
        n = int(lib().fri_plan_emission_packed_size(self._h, int(bits)))
        if n == 0 and self.n_tiles:
            raise FriError(FRI_E_INVALID, lib().fri_last_error().decode("utf-8", "replace"))
        return n

    def encode_emit_packed(self, pixels: np.ndarray, q=None, bits: int = 10, out: np.ndarray | None = None) -> np.ndarray:
        """HWC pixels [F, H, W, C] -> uint8 [F, C, emission_packed_size(bits)] (fri_encode_tq_emit_packed)."""
        px, n = self._frames(pixels)
        nb = self.emission_packed_size(bits)
        if out is None:
            out = np.empty((n, self.channels, nb), np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == (n, self.channels, nb)
        qa, qp = _q_array(q)
        _check(lib().fri_encode_tq_emit_packed(self._h, px.ctypes.data, n, qp, int(bits), out.ctypes.data))
        return out

    def decode_emit_packed(self, packed: np.ndarray, q=None, bits: int = 10, multiply: bool = False,
                           out: np.ndarray | None = None) -> np.ndarray:
        """uint8 [F, C, emission_packed_size(bits)] -> HWC pixels [F, H, W, C] (fri_decode_tq_emit_packed)."""
        pk = np.ascontiguousarray(packed, dtype=np.uint8)
        nb = self.emission_packed_size(bits)
        if pk.shape == (self.channels, nb):
            pk = pk[None]
        if pk.ndim != 3 or pk.shape[1:] != (self.channels, nb):
            raise ValueError(f"packed streams must have shape [F, {self.channels}, {nb}]")
        n = pk.shape[0]
        if out is None:
            out = np.empty((n,) + self.frame_shape, self.pixel_dtype)
        assert out.dtype == self.pixel_dtype and out.flags.c_contiguous and out.shape == (n,) + self.frame_shape
        qa, qp = _q_array(q)
        mode = FRI_DEQUANT_MULTIPLY if multiply else FRI_DEQUANT_DIVIDE
        _check(lib().fri_decode_tq_emit_packed(self._h, pk.ctypes.data, n, int(bits), qp, mode, out.ctypes.data))
        return out

    def emit_device_packed(self, d_coefs: int, n_frames: int, bits: int, d_out: int, stream: int = 0) -> None:
        _check(lib().fri_emit_device_packed(self._h, d_coefs, n_frames, int(bits), d_out, stream))

    def unemit_device_packed(self, d_packed: int, n_frames: int, bits: int, d_coefs: int, stream: int = 0) -> None:
        _check(lib().fri_unemit_device_packed(self._h, d_packed, n_frames, int(bits), d_coefs, stream))

    def encode_emit10(self, pixels: np.ndarray, q=None, out: np.ndarray | None = None) -> np.ndarray:
        """HWC pixels [F, H, W, C] -> uint8 [F, C, emission_packed_bytes()]: emission-ordered streams in the
        10-bit packed transport (fri_encode_tq_emit10); unpack10() gives the coefficients back."""
        px, n = self._frames(pixels)
        nb = self.emission_packed_bytes()
        if out is None:
            out = np.empty((n, self.channels, nb), np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == (n, self.channels, nb)
        qa, qp = _q_array(q)
        _check(lib().fri_encode_tq_emit10(self._h, px.ctypes.data, n, qp, out.ctypes.data))
        return out

    def decode_emit10(self, packed: np.ndarray, q=None, multiply: bool = False, out: np.ndarray | None = None) -> np.ndarray:
        """uint8 [F, C, emission_packed_bytes()] packed streams -> HWC pixels [F, H, W, C] (fri_decode_tq_emit10)."""
        pk = np.ascontiguousarray(packed, dtype=np.uint8)
        nb = self.emission_packed_bytes()
        if pk.shape == (self.channels, nb):
            pk = pk[None]
        if pk.ndim != 3 or pk.shape[1:] != (self.channels, nb):
            raise ValueError(f"packed streams must have shape [F, {self.channels}, {nb}]")
        n = pk.shape[0]
        if out is None:
            out = np.empty((n,) + self.frame_shape, self.pixel_dtype)
        assert out.dtype == self.pixel_dtype and out.flags.c_contiguous and out.shape == (n,) + self.frame_shape
        qa, qp = _q_array(q)
        mode = FRI_DEQUANT_MULTIPLY if multiply else FRI_DEQUANT_DIVIDE
        _check(lib().fri_decode_tq_emit10(self._h, pk.ctypes.data, n, qp, mode, out.ctypes.data))
        return out

    def emit_device10(self, d_coefs: int, n_frames: int, d_out: int, stream: int = 0) -> None:
        _check(lib().fri_emit_device10(self._h, d_coefs, n_frames, d_out, stream))

    def unemit_device10(self, d_packed: int, n_frames: int, d_coefs: int, stream: int = 0) -> None:
        _check(lib().fri_unemit_device10(self._h, d_packed, n_frames, d_coefs, stream))

    def predict_device(self, d_coefs: int, n_frames: int, value_params, width_params, d_bucket: int, d_pred: int, d_sym: int,
                       d_hist: int, d_overflow: int = 0, stream: int = 0) -> None:
        """Prediction + context bucketing of quantized dense blocks (fri_predict_device): value_params / width_params are
        float32 [C, 3, 6]; outputs are device pointers (uint8 / int32 / uint16 [F, C, emission_count()], uint32
        [F, C, 10, 1024], optional uint32 overflow counter)."""
        vp = np.ascontiguousarray(value_params, dtype=np.float32)
        wp = np.ascontiguousarray(width_params, dtype=np.float32)
        if vp.shape != (self.channels, 3, 6) or wp.shape != (self.channels, 3, 6):
            raise ValueError(f"predictor parameters must have shape ({self.channels}, 3, 6)")
        _check(lib().fri_predict_device(self._h, d_coefs, n_frames, vp.ctypes.data, wp.ctypes.data, d_bucket, d_pred, d_sym,
                                        d_hist, d_overflow or None, stream))

    def emit_device(self, d_coefs: int, n_frames: int, d_out: int, stream: int = 0, half: bool = False) -> None:
        """d_out: int32 (or, with half=True, int16) [n_frames, C, emission_count()] on the device."""
        fn = lib().fri_emit_device16 if half else lib().fri_emit_device
        _check(fn(self._h, d_coefs, n_frames, d_out, stream))

    def encode_emit(self, pixels: np.ndarray, q=None, out: np.ndarray | None = None, dtype=np.int32) -> np.ndarray:
        """HWC pixels [F, H, W, C] -> [F, C, emission_count()]: the quantized Some coefficients of every
        channel in emission order, int32 or int16 (`out.dtype` decides when `out` is given)."""
        px, n = self._frames(pixels)
        cnt = self.emission_count()
        if out is None:
            out = np.empty((n, self.channels, cnt), dtype)
        assert out.dtype in (np.int32, np.int16) and out.flags.c_contiguous and out.shape == (n, self.channels, cnt)
        qa, qp = _q_array(q)
        fn = lib().fri_encode_tq_emit16 if out.dtype == np.int16 else lib().fri_encode_tq_emit
        _check(fn(self._h, px.ctypes.data, n, qp, out.ctypes.data))
        return out

    def unemit_device(self, d_streams: int, n_frames: int, d_coefs: int, stream: int = 0, half: bool = False) -> None:
        """Emitted streams int32 (half: int16) [n_frames, C, emission_count()] -> dense coefficient blocks."""
        fn = lib().fri_unemit_device16 if half else lib().fri_unemit_device
        _check(fn(self._h, d_streams, n_frames, d_coefs, stream))

    def decode_emit(self, streams: np.ndarray, q=None, multiply: bool = False, out: np.ndarray | None = None) -> np.ndarray:
        """Emitted streams [F, C, emission_count()] (int32 or int16) -> HWC pixels [F, H, W, C]."""
        st = np.asarray(streams)
        half = st.dtype == np.int16
        st = np.ascontiguousarray(st, dtype=np.int16 if half else np.int32)
        cnt = self.emission_count()
        if st.shape == (self.channels, cnt):
            st = st[None]
        if st.ndim != 3 or st.shape[1:] != (self.channels, cnt):
            raise ValueError(f"streams must have shape [F, {self.channels}, {cnt}]")
        n = st.shape[0]
        if out is None:
            out = np.empty((n,) + self.frame_shape, self.pixel_dtype)
        assert out.dtype == self.pixel_dtype and out.flags.c_contiguous and out.shape == (n,) + self.frame_shape
        qa, qp = _q_array(q)
        mode = FRI_DEQUANT_MULTIPLY if multiply else FRI_DEQUANT_DIVIDE
        fn = lib().fri_decode_tq_emit16 if half else lib().fri_decode_tq_emit
        _check(fn(self._h, st.ctypes.data, n, qp, mode, out.ctypes.data))
        return out

    # ---- host codec behind the transform: parameter fit, context model, rANS, frif container -----------
    def _dense(self, coefs: np.ndarray) -> np.ndarray:
        cf = np.ascontiguousarray(coefs, dtype=np.int32)
        if cf.shape != self.coef_shape:
            raise ValueError(f"coefs must have shape {self.coef_shape} (one frame)")
        return cf

    def fit_parameters(self, coefs: np.ndarray):
        """Quantized dense blocks of one frame -> (value_params, width_params), float32 [C, 3, 6] each."""
        cf = self._dense(coefs)
        vp = np.zeros((self.channels, 3, 6), np.float32)
        wp = np.zeros((self.channels, 3, 6), np.float32)
        _check(lib().fri_fit_parameters(self._h, cf.ctypes.data, vp.ctypes.data, wp.ctypes.data))
        return vp, wp

    def fit_device(self, d_coefs: int, stream: int = 0):
        """The same fit for one device-resident frame (sums on the device, solve on the host): bit-identical
        parameters to fit_parameters on the same coefficients."""
        vp = np.zeros((self.channels, 3, 6), np.float32)
        wp = np.zeros((self.channels, 3, 6), np.float32)
        _check(lib().fri_fit_device(self._h, C.c_void_p(d_coefs), vp.ctypes.data, wp.ctypes.data, C.c_void_p(stream)))
        return vp, wp

    def predict_host(self, coefs: np.ndarray, value_params, width_params):
        """Host form of predict_device for one frame: (bucket u8 [C, n], prediction i32 [C, n], symbol u16 [C, n],
        histograms u32 [C, 10, 1024], overflow count)."""
        cf = self._dense(coefs)
        vp = np.ascontiguousarray(value_params, dtype=np.float32)
        wp = np.ascontiguousarray(width_params, dtype=np.float32)
        n = self.emission_count()
        b = np.zeros((self.channels, n), np.uint8)
        p = np.zeros((self.channels, n), np.int32)
        s = np.zeros((self.channels, n), np.uint16)
        h = np.zeros((self.channels, 10, 1024), np.uint32)
        o = C.c_uint32(0)
        _check(lib().fri_predict_host(self._h, cf.ctypes.data, vp.ctypes.data, wp.ctypes.data, b.ctypes.data, p.ctypes.data,
                                      s.ctypes.data, h.ctypes.data, C.addressof(o)))
        return b, p, s, h, int(o.value)

    @staticmethod
    def _take_bytes(ptr, n) -> bytes:
        try:
            return C.string_at(ptr.value, n.value)
        finally:
            lib().fri_frv_free(ptr)

    def frv_pack(self, value_params, width_params, bucket, sym, hist, colorspace: int = 0) -> bytes:
        """Symbols + buckets + histograms of one frame -> `frif` container bytes (rANS on the host)."""
        vp = np.ascontiguousarray(value_params, dtype=np.float32)
        wp = np.ascontiguousarray(width_params, dtype=np.float32)
        b = np.ascontiguousarray(bucket, dtype=np.uint8)
        s = np.ascontiguousarray(sym, dtype=np.uint16)
        h = np.ascontiguousarray(hist, dtype=np.uint32)
        n = self.emission_count()
        assert b.shape == (self.channels, n) and s.shape == (self.channels, n) and h.shape == (self.channels, 10, 1024)
        out, ln = _P(), C.c_size_t(0)
        _check(lib().fri_frv_pack(self._h, colorspace, vp.ctypes.data, wp.ctypes.data, b.ctypes.data, s.ctypes.data, h.ctypes.data,
                                  C.byref(out), C.byref(ln)))
        return self._take_bytes(out, ln)

    def frv_unpack(self, data: bytes) -> np.ndarray:
        """Container bytes -> quantized dense blocks [n_tiles, C, 512] (serial entropy decode on the host)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        out = np.empty(self.coef_shape, np.int32)
        _check(lib().fri_frv_unpack(self._h, buf.ctypes.data, len(data), out.ctypes.data))
        return out

    def frv_encode(self, pixels: np.ndarray, q=None, colorspace: int = 0) -> bytes:
        """HWC pixels of one frame -> container bytes (FRIEncoder::encode with the transform, quantizer and
        prediction on the device)."""
        px, n = self._frames(pixels)
        assert n == 1
        qa, qp = _q_array(q)
        out, ln = _P(), C.c_size_t(0)
        _check(lib().fri_frv_encode(self._h, px.ctypes.data, qp, colorspace, C.byref(out), C.byref(ln)))
        return self._take_bytes(out, ln)

    def frv_decode(self, data: bytes, q=None, multiply: bool = False) -> np.ndarray:
        """Container bytes -> HWC pixels [H, W, C] (FRIDecoder::decode)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        out = np.empty(self.frame_shape, self.pixel_dtype)
        qa, qp = _q_array(q)
        mode = FRI_DEQUANT_MULTIPLY if multiply else FRI_DEQUANT_DIVIDE
        _check(lib().fri_frv_decode(self._h, buf.ctypes.data, len(data), qp, mode, out.ctypes.data))
        return out

    # ---- one image split over several GPUs by ranges of tile groups -----------
    def part(self, part: int, n_parts: int) -> dict:
        """Groups, tiles (plan order) and pixel rows of part `part` of `n_parts` (fri_plan_part)."""
        v = [C.c_uint32() for _ in range(6)]
        _check(lib().fri_plan_part(self._h, part, n_parts, *[C.byref(x) for x in v]))
        keys = ("group_begin", "group_end", "tile_begin", "tile_end", "row_begin", "row_end")
        return {k: int(x.value) for k, x in zip(keys, v)}

    def encode_device_part(self, d_pixel_rows: int, d_coef_tiles: int, part: int, n_parts: int, q=None, stream: int = 0) -> None:
        """d_pixel_rows: the band of pixel rows [row_begin, row_end) of the part; d_coef_tiles: int32 blocks of its tiles."""
        qa, qp = _q_array(q)
        _check(lib().fri_encode_tq_device_part(self._h, d_pixel_rows, qp, d_coef_tiles, part, n_parts, stream))

    def decode_device_part(self, d_coef_tiles: int, d_pixel_rows: int, part: int, n_parts: int, q=None, multiply: bool = False,
                           stream: int = 0) -> None:
        """Writes only the pixels the part's tiles own into the band (rows [row_begin, row_end))."""
        qa, qp = _q_array(q)
        mode = FRI_DEQUANT_MULTIPLY if multiply else FRI_DEQUANT_DIVIDE
        _check(lib().fri_decode_tq_device_part(self._h, d_coef_tiles, qp, mode, d_pixel_rows, part, n_parts, stream))

    def groups_in_rows(self, group_begin: int, group_end: int, row_begin: int, row_end: int) -> dict:
        """Smallest contiguous sub-range of groups [group_begin, group_end) holding every group that touches pixel rows
        [row_begin, row_end), and the rows that sub-range touches in all (fri_plan_groups_in_rows)."""
        v = [C.c_uint32() for _ in range(4)]
        _check(lib().fri_plan_groups_in_rows(self._h, group_begin, group_end, row_begin, row_end, *[C.byref(x) for x in v]))
        return {k: int(x.value) for k, x in zip(("first", "last", "span_begin", "span_end"), v)}

    def decode_device_groups(self, d_coef_tiles: int, tile_first: int, d_pixel_rows: int, row_first: int, group_begin: int,
                             group_end: int, q=None, multiply: bool = False, stream: int = 0) -> None:
        """Inverse transform of groups [group_begin, group_end): d_coef_tiles points at tile `tile_first`'s block,
        d_pixel_rows at frame row `row_first` (may be negative); only the pixels those groups' tiles own are written."""
        qa, qp = _q_array(q)
        mode = FRI_DEQUANT_MULTIPLY if multiply else FRI_DEQUANT_DIVIDE
        _check(lib().fri_decode_tq_device_groups(self._h, d_coef_tiles, tile_first, qp, mode, d_pixel_rows, row_first, group_begin,
                                                 group_end, stream))

    # ---- device-resident entry points (raw device pointers, e.g. torch.Tensor.data_ptr()) -----
    def encode_device(self, d_pixels: int, n_frames: int, d_coefs: int, q=None, stream: int = 0, half: bool = False) -> None:
        """d_coefs: int32 (or, with half=True, int16) [n_frames, n_tiles, C, 2^depth] on the device."""
        qa, qp = _q_array(q)
        fn = lib().fri_encode_tq_device16 if half else lib().fri_encode_tq_device
        _check(fn(self._h, d_pixels, n_frames, qp, d_coefs, stream))

    def decode_device(self, d_coefs: int, n_frames: int, d_pixels: int, q=None, multiply: bool = False,
                      stream: int = 0, half: bool = False) -> None:
        qa, qp = _q_array(q)
        mode = FRI_DEQUANT_MULTIPLY if multiply else FRI_DEQUANT_DIVIDE
        fn = lib().fri_decode_tq_device16 if half else lib().fri_decode_tq_device
        _check(fn(self._h, d_coefs, n_frames, qp, mode, d_pixels, stream))
