"""Host-side mirror of the reference's stage interface for the transform + quantization path.

The reference (pure Rust, crates/libfri/src) runs four stage functions on this path:

    wavelet_transform::encode(RasterImage, &EncoderOpts) -> Result<WaveletImage, String>   wavelet_transform.rs:708-713
    quantization::encode(WaveletImage)                   -> Result<WaveletImage, String>   quantization.rs:7-25
    quantization::decode(WaveletImage)                   -> Result<WaveletImage, String>   quantization.rs:27-45
    wavelet_transform::decode(WaveletImage)              -> Result<RasterImage, String>    wavelet_transform.rs:715-717

Here the same four names exist with the same argument meaning; the arithmetic always runs in
libfri_cuda (fused: the transform stage also applies the quantization matrix, the quantization
stage then only marks the image).  Errors surface as StageError, the analogue of the
reference's Err(String); there is no CPU fallback.
"""
from __future__ import annotations

import dataclasses
import enum

import numpy as np

from . import capi


class StageError(RuntimeError):
    """Err(String) of a stage function."""


class ColorSpace(enum.Enum):  # images.rs:8-39
    Luma = 0
    YCbCr = 1
    RGB = 2

    @property
    def num_channels(self) -> int:  # images.rs:15-21
        return 1 if self is ColorSpace.Luma else 3


@dataclasses.dataclass
class ImageMetadata:  # images.rs:68-79
    height: int
    width: int
    colorspace: ColorSpace = ColorSpace.RGB
    variant: str = "TameTwindragon"  # encoder.rs:96 — the only variant ever produced


@dataclasses.dataclass
class RasterImage:  # images.rs:82-85 — data is HWC interleaved
    metadata: ImageMetadata
    data: np.ndarray

    @staticmethod
    def from_array(a: np.ndarray, colorspace: ColorSpace | None = None) -> "RasterImage":
        a = np.ascontiguousarray(a)
        if a.ndim == 2:
            a = a[:, :, None]
        cs = colorspace or (ColorSpace.Luma if a.shape[2] == 1 else ColorSpace.RGB)
        if cs.num_channels != a.shape[2]:
            raise StageError("colorspace does not match the channel count")
        return RasterImage(ImageMetadata(a.shape[0], a.shape[1], cs), a)


@dataclasses.dataclass
class EncoderOpts:  # encoder.rs:58-64 (+ the matrix the reference hard-wires, quantization.rs:3-5)
    verbose: bool = False
    emit_coefficients: bool = False
    quantization_matrix: np.ndarray = dataclasses.field(default_factory=lambda: np.ones(32, np.int32))
    depth: int = capi.FRI_BASE_DEPTH
    device: int = 0


@dataclasses.dataclass
class WaveletImage:
    """wavelet_transform.rs:384-389.  fractal_lattice (a HashMap centre -> Fractal in the
    reference) is kept dense: centers[n, 2], coefficients[n, C, 2^depth] and the Some/None
    mask some[n, 2^depth] (channel independent)."""

    metadata: ImageMetadata
    centers: np.ndarray
    coefficients: np.ndarray
    some: np.ndarray
    depth: int
    quantization_matrix: np.ndarray
    sample_bytes: int = 1
    device: int = 0
    quantized: bool = False

    def fractal_lattice(self) -> dict:
        """centre -> per-channel list of Option<i32> (None where the reference holds None)."""
        out = {}
        for i, (re, im) in enumerate(self.centers.tolist()):
            chans = []
            for ch in range(self.coefficients.shape[1]):
                v = self.coefficients[i, ch]
                chans.append([int(x) if s else None for x, s in zip(v, self.some[i])])
            out[(re, im)] = chans
        return out


_plans: dict = {}


def _plan(width, height, channels, depth, sample_bytes, device) -> capi.Plan:
    key = (width, height, channels, depth, sample_bytes, device)
    if key not in _plans:
        try:
            _plans[key] = capi.Plan(width, height, channels, depth, sample_bytes, device)
        except capi.FriError as e:
            raise StageError(str(e)) from e
    return _plans[key]


class wavelet_transform:  # noqa: N801 — module name in the reference
    @staticmethod
    def encode(image: RasterImage, opts: EncoderOpts | None = None) -> WaveletImage:
        opts = opts or EncoderOpts()
        h, w, c = image.data.shape
        sb = image.data.dtype.itemsize
        plan = _plan(w, h, c, opts.depth, sb, opts.device)
        try:
            coefs = plan.encode(image.data, opts.quantization_matrix)[0]
        except capi.FriError as e:
            raise StageError(str(e)) from e
        return WaveletImage(image.metadata, plan.centers(), coefs, plan.masks(), opts.depth,
                            np.asarray(opts.quantization_matrix, np.int32), sb, opts.device, quantized=True)

    @staticmethod
    def decode(image: WaveletImage, multiply: bool = False) -> RasterImage:
        md = image.metadata
        c = image.coefficients.shape[1]
        plan = _plan(md.width, md.height, c, image.depth, image.sample_bytes, image.device)
        if not np.array_equal(plan.centers(), image.centers):
            raise StageError("WaveletImage tile order does not match the plan for this image size")
        try:
            q = image.quantization_matrix if not image.quantized else None
            px = plan.decode(image.coefficients, q, multiply=multiply)[0]
        except capi.FriError as e:
            raise StageError(str(e)) from e
        return RasterImage(md, px)


class quantization:  # noqa: N801
    @staticmethod
    def encode(image: WaveletImage) -> WaveletImage:
        """quantization.rs:7-25.  The division already happened inside the fused transform kernel."""
        if not image.quantized:
            raise StageError("coefficients did not come from wavelet_transform.encode")
        return image

    @staticmethod
    def decode(image: WaveletImage) -> WaveletImage:
        """quantization.rs:27-45: hands the matrix to the fused decode kernel (which divides
        again, as the reference does, unless wavelet_transform.decode(multiply=True))."""
        return dataclasses.replace(image, quantized=False)
