"""Host-side mirror of the reference's stage interface: the transform + quantization path and, behind it,
the prediction / entropy-coding / serialization stages and the two public drivers FRIEncoder / FRIDecoder.

The reference (pure Rust, crates/libfri/src) runs four stage functions on this path:

    wavelet_transform::encode(RasterImage, &EncoderOpts) -> Result<WaveletImage, String>   wavelet_transform.rs:708-713
    quantization::encode(WaveletImage)                   -> Result<WaveletImage, String>   quantization.rs:7-25
    quantization::decode(WaveletImage)                   -> Result<WaveletImage, String>   quantization.rs:27-45
    wavelet_transform::decode(WaveletImage)              -> Result<RasterImage, String>    wavelet_transform.rs:715-717

Here the same four names exist with the same argument meaning; the arithmetic always runs in
libfri_cuda (fused: the transform stage also applies the quantization matrix, the quantization
stage then only marks the image).  Errors surface as StageError, the analogue of the
reference's Err(String); there is no CPU fallback.

The later stages keep the reference's names too (prediction.rs:224, entropy_coding.rs:266 / :354,
serialize.rs:48 / :119, encoder.rs:87, decoder.rs:48).  FRIEncoder runs prediction on the device
(fri_predict_device inside fri_frv_encode) with parameters fitted on the host; the stand-alone prediction stage
below works on host arrays (fri_predict_host, the same arithmetic).  rANS and the `frif` container are this
library's own host code (fri_codec.cpp) — containers written here decode here; byte identity with the
reference's files is unpinned.
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


# ---- the stages behind the quantizer (encoder.rs:34-45, decoder.rs:19-26) -------------------------------------
_CS_CODE = {ColorSpace.Luma: 1, ColorSpace.RGB: 2, ColorSpace.YCbCr: 3}  # images.rs:23-29
_CS_FROM_CODE = {v: k for k, v in _CS_CODE.items()}


@dataclasses.dataclass
class Contexts:
    """What prediction::encode hands to entropy_coding::encode: the fitted predictor parameters and, per emitted
    coefficient, (bucket, symbol) — the reference keeps them in Fractal::parameter_predictors and ten AnsContexts
    per channel (prediction.rs:224-323)."""

    value_params: np.ndarray  # [C, 3, 6] float32
    width_params: np.ndarray  # [C, 3, 6]
    bucket: np.ndarray        # [C, n] uint8, emission order
    symbol: np.ndarray        # [C, n] uint16, pack_signed(value - prediction)
    histograms: np.ndarray    # [C, 10, 1024] uint32


@dataclasses.dataclass
class CompressedImage:  # images.rs:121-124, serialized
    metadata: ImageMetadata
    data: bytes


def _plan_of(image: WaveletImage) -> capi.Plan:
    md = image.metadata
    if image.depth != capi.FRI_BASE_DEPTH or image.sample_bytes != 1:
        raise StageError("the entropy stages exist for depth 9 and 8-bit samples (the reference's reach)")
    return _plan(md.width, md.height, image.coefficients.shape[1], image.depth, 1, image.device)


class prediction:  # noqa: N801
    @staticmethod
    def encode(image: WaveletImage, opts: EncoderOpts | None = None) -> Contexts:
        """prediction.rs:224-323: parameter fit (host least squares), then bucket / prediction / symbol of every
        coefficient and the per-context histograms."""
        plan = _plan_of(image)
        try:
            vp, wp = plan.fit_parameters(image.coefficients)
            b, _p, s, h, over = plan.predict_host(image.coefficients, vp, wp)
        except capi.FriError as e:
            raise StageError(str(e)) from e
        if over:
            raise StageError(f"{over} residual(s) fall outside the 1024-symbol alphabet (the reference panics here)")
        return Contexts(vp, wp, b, s, h)


class entropy_coding:  # noqa: N801
    @staticmethod
    def encode(image: WaveletImage, contexts: Contexts, opts: EncoderOpts | None = None) -> CompressedImage:
        """entropy_coding.rs:266-352 fused with serialize.rs:48-117 (the container is what crosses the boundary)."""
        plan = _plan_of(image)
        try:
            data = plan.frv_pack(contexts.value_params, contexts.width_params, contexts.bucket, contexts.symbol,
                                 contexts.histograms, _CS_CODE[image.metadata.colorspace])
        except capi.FriError as e:
            raise StageError(str(e)) from e
        return CompressedImage(image.metadata, data)

    @staticmethod
    def decode(image: CompressedImage, device: int = 0) -> WaveletImage:
        """entropy_coding.rs:354-449: serial decode with the host predictor."""
        md = image.metadata
        plan = _plan(md.width, md.height, md.colorspace.num_channels, capi.FRI_BASE_DEPTH, 1, device)
        try:
            coefs = plan.frv_unpack(image.data)
        except capi.FriError as e:
            raise StageError(str(e)) from e
        return WaveletImage(md, plan.centers(), coefs, plan.masks(), capi.FRI_BASE_DEPTH, np.ones(32, np.int32), 1, device,
                            quantized=True)


class serialize:  # noqa: N801
    @staticmethod
    def encode(image: CompressedImage) -> bytes:  # serialize.rs:48-117 (done inside entropy_coding.encode)
        return image.data

    @staticmethod
    def decode(data: bytes) -> CompressedImage:  # serialize.rs:119-149: header only; the payload stays packed
        try:
            w, h, c = capi.frv_info(data)
        except capi.FriError as e:
            raise StageError(str(e)) from e
        code = (int.from_bytes(data[12:16], "little") >> 30) & 3
        return CompressedImage(ImageMetadata(h, w, _CS_FROM_CODE[code]), bytes(data))


class FRIEncoder:
    """encoder.rs:83-109: FRIEncoder::new(opts).encode(data, height, width, colorspace) -> bytes."""

    def __init__(self, opts: EncoderOpts | None = None):
        self.opts = opts or EncoderOpts()

    def encode(self, data, height: int, width: int, colorspace: ColorSpace = ColorSpace.RGB) -> bytes:
        px = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else np.asarray(data, np.uint8)
        try:
            px = px.reshape(height, width, colorspace.num_channels)
            plan = _plan(width, height, colorspace.num_channels, capi.FRI_BASE_DEPTH, 1, self.opts.device)
            return plan.frv_encode(px, self.opts.quantization_matrix, _CS_CODE[colorspace])
        except (capi.FriError, ValueError) as e:
            raise StageError("Failed to decode: " + str(e)) from e  # the reference's message, encoder.rs:106


class FRIDecoder:
    """decoder.rs:44-59: FRIDecoder{}.decode(bytes) -> RasterImage."""

    def __init__(self, device: int = 0, quantization_matrix=None):
        self.device = device
        self.q = np.ones(32, np.int32) if quantization_matrix is None else np.asarray(quantization_matrix, np.int32)

    def decode(self, data: bytes) -> RasterImage:
        try:
            ci = serialize.decode(data)
            md = ci.metadata
            plan = _plan(md.width, md.height, md.colorspace.num_channels, capi.FRI_BASE_DEPTH, 1, self.device)
            return RasterImage(md, plan.frv_decode(data, self.q))
        except (capi.FriError, StageError) as e:
            raise StageError("Failed to decode: " + str(e)) from e
