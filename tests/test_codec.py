"""Host codec behind the transform: context model, rANS, `frif` container and
the host predictor.  CPU tests run the pure-host entry points on coefficients from the oracle; the -m gpu tests
run the whole pipeline (device transform + prediction, host entropy coder) through fri_frv_encode / _decode."""
import struct

import numpy as np
import pytest

from frave_b200 import capi
from oracle import c_oracle as O
from oracle import fri_entropy_np as EN
from oracle import fri_predict_np as PR
from tests.conftest import smallest_layer_q, smooth_image, uniform_image

ONES = np.ones(32, np.int32)


def _coefs(plan, img, q=ONES):
    coef, some = O.extract_tiles(img, plan.centers())
    return O.quantize(coef, some, q)


@pytest.mark.parametrize("shape,smooth", [((48, 64, 1), True), ((131, 77, 3), True), ((131, 77, 3), False), ((100, 37, 3), True)],
                         ids=["48x64x1", "131x77x3", "131x77x3-noise", "37x100x3"])
def test_host_predictor_and_container_round_trip(shape, smooth):
    h, w, c = shape
    img = (smooth_image if smooth else uniform_image)(h, w, c, seed=h)
    with capi.Plan(w, h, c, device=-1) as plan:
        coefs = _coefs(plan, img, smallest_layer_q(2))
        some = plan.masks()
        order = plan.emission_order().astype(np.int64)
        src = order[some.reshape(-1)[order]]
        vp, wp = plan.fit_parameters(coefs)
        assert vp.shape == wp.shape == (c, 3, 6) and np.isfinite(vp).all() and np.isfinite(wp).all()
        b, p, s, hist, over = plan.predict_host(coefs, vp, wp)
        wb, wpv, ws, wh, wo = PR.predict(plan.centers(), coefs, some, src, vp, wp)  # the dict-based restatement
        assert np.array_equal(b, wb) and np.array_equal(p, wpv) and np.array_equal(hist, wh) and over == wo
        assert np.array_equal(s, np.minimum(ws, 0xffff).astype(np.uint16))
        assert over == 0
        data = plan.frv_pack(vp, wp, b, s, hist)
        # an independently written Python restatement of the context model + rANS + container gives the same bytes
        assert data == EN.encode(h, w, 1 if c == 1 else 2, vp, wp, b, s, hist)
        assert capi.frv_info(data) == (w, h, c)
        assert np.array_equal(plan.frv_unpack(data), coefs)  # serial decode with the host predictor
        assert data[:4] == b"frif" and data[-2:] == b"\xff\xdf"
        assert struct.unpack("<II", data[4:12]) == (h, w)


def test_context_tables_are_rebuilt_identically_from_the_header():
    """The container stores max_freq_bits and the off-distribution symbols only (serialize.rs:93-105); the decoder
    rebuilds the same Laplace-model tables (serialize.rs:216-236).  Checked through a round trip that uses a symbol
    far outside the model's support."""
    h, w, c = 48, 64, 1
    img = uniform_image(h, w, c, seed=3)
    with capi.Plan(w, h, c, device=-1) as plan:
        coefs = _coefs(plan, img)
        vp = np.zeros((c, 3, 6), np.float32)
        wp = np.zeros((c, 3, 6), np.float32)  # width 0 -> bucket 0 (narrowest model) for every high-frequency coefficient
        b, p, s, hist, over = plan.predict_host(coefs, vp, wp)
        assert over == 0 and (b[0, 2 * plan.n_tiles:] == 0).all()
        assert s.max() > 300  # raw residues of a noise image: far in the tail of a width-2.5 Laplace model
        data = plan.frv_pack(vp, wp, b, s, hist)
        assert data == EN.encode(h, w, 1, vp, wp, b, s, hist)
        assert np.array_equal(plan.frv_unpack(data), coefs)
        ctx = EN.context_from_counts(hist[0][0], 0)
        assert len(ctx.off_distribution_values) > 0 and sum(ctx.freqs) == 1 << ctx.max_freq_bits


@pytest.mark.parametrize("seed", [0, 1])
def test_rans_bytes_for_arbitrary_symbol_streams(seed):
    """The division-free coding step and the decoder's coarse slot table against the Python restatement's plain
    integer arithmetic, on symbol / bucket streams that do not come from an image: wide range of frequencies
    (2^17 symbols per context at most), rare symbols far in the tail."""
    h, w, c = 300, 420, 1
    rng = np.random.default_rng(seed)
    with capi.Plan(w, h, c, device=-1) as plan:
        n = plan.emission_count()
        b = rng.integers(0, 10, size=(c, n)).astype(np.uint8)
        scale = np.array([0.7, 1.5, 3, 5, 8, 12, 20, 40, 80, 160])[b]
        s = np.minimum(np.abs(rng.laplace(0, scale)).astype(np.int64), 1023).astype(np.uint16)
        if seed:
            b[:] = 3  # one context only: the other nine stay at their unused default
        hist = np.zeros((c, 10, 1024), np.uint32)
        np.add.at(hist[0], (b[0], s[0]), 1)
        vp = rng.normal(size=(c, 3, 6)).astype(np.float32)
        wp = rng.normal(size=(c, 3, 6)).astype(np.float32)
        data = plan.frv_pack(vp, wp, b, s, hist)
        assert data == EN.encode(h, w, 1, vp, wp, b, s, hist)


def test_symbols_outside_the_alphabet_are_refused():
    h, w, c = 48, 64, 1
    with capi.Plan(w, h, c, device=-1) as plan:
        n = plan.emission_count()
        b = np.zeros((c, n), np.uint8)
        s = np.zeros((c, n), np.uint16)
        s[0, 5] = 1024
        hist = np.zeros((c, 10, 1024), np.uint32)
        hist[0, 0, 0] = n
        with pytest.raises(capi.FriError) as ei:
            plan.frv_pack(np.zeros((c, 3, 6), np.float32), np.zeros((c, 3, 6), np.float32), b, s, hist)
        assert ei.value.code == capi.FRI_E_UNSUPPORTED and "alphabet" in str(ei.value)


def test_malformed_containers_are_rejected():
    h, w, c = 48, 64, 1
    img = smooth_image(h, w, c, seed=9)
    with capi.Plan(w, h, c, device=-1) as plan, capi.Plan(w + 1, h, c, device=-1) as other:
        coefs = _coefs(plan, img)
        vp, wp = plan.fit_parameters(coefs)
        b, p, s, hist, _ = plan.predict_host(coefs, vp, wp)
        data = plan.frv_pack(vp, wp, b, s, hist)
        for bad in (b"frig" + data[4:], data[:40], data[:-2] + b"\xff\x00", data[:16] + b"\x00\x00" + data[18:]):
            with pytest.raises(capi.FriError) as ei:
                plan.frv_unpack(bad)
            assert ei.value.code == capi.FRI_E_INVALID
        with pytest.raises(capi.FriError) as ei:
            other.frv_unpack(data)
        assert "plan is for" in str(ei.value)
        with pytest.raises(capi.FriError):
            capi.frv_info(b"nope")
        # truncating the entropy-coded payload (DAT length left intact is caught by the bounds check, a shorter
        # declared length by the decoder running dry)
        i = data.index(b"\xff\xb4")
        n = struct.unpack("<Q", data[i + 2:i + 10])[0]
        cut = data[:i + 2] + struct.pack("<Q", n // 2) + data[i + 10:i + 10 + n // 2] + data[i + 10 + n:]
        with pytest.raises(capi.FriError):
            plan.frv_unpack(cut)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,smooth", [((270, 480, 3), True), ((512, 512, 1), True), ((1080, 1920, 3), True), ((131, 77, 3), False)],
                         ids=["270x480x3", "512x512x1", "1080p", "131x77x3-noise"])
def test_frv_encode_decode_on_the_device(shape, smooth):
    """fri_frv_encode / fri_frv_decode: pixels -> `frif` bytes -> pixels.  Lossless at the reference's all-ones
    matrix; with the smallest-layer quantizer the decoded pixels equal the oracle's reconstruction; the device
    prediction stage agrees with the host predictor and the bytes with the Python restatement."""
    h, w, c = shape
    img = (smooth_image if smooth else uniform_image)(h, w, c, seed=w)
    with capi.Plan(w, h, c) as plan:
        data = plan.frv_encode(img)
        assert capi.frv_info(data) == (w, h, c)
        rec = plan.frv_decode(data)
        covered = plan.pixels_covered == w * h
        if covered:
            assert np.array_equal(rec, img)
        coefs = plan.encode(img)[0]
        assert np.array_equal(plan.frv_unpack(data), coefs)
        # bytes: host predictor + Python entropy coder on the same coefficients and the fit the library made
        vp, wp = plan.fit_parameters(coefs)
        b, p, s, hist, over = plan.predict_host(coefs, vp, wp)
        assert over == 0
        assert data == plan.frv_pack(vp, wp, b, s, hist)
        if h * w <= 270 * 480:
            assert data == EN.encode(h, w, 1 if c == 1 else 2, vp, wp, b, s, hist)
        # quantized: q[8] = q[9] = 3; the decoder divides again (quantization.rs:37)
        q = smallest_layer_q(3)
        dq = plan.frv_encode(img, q)
        assert len(dq) < len(data)
        cq = plan.encode(img, q)[0]
        some = np.broadcast_to(plan.masks()[:, None, :], plan.coef_shape)
        want = O.extract_values(plan.centers(), O.quantize(cq, some, q), some, h, w)
        assert np.array_equal(plan.frv_decode(dq, q), want)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,smooth", [((270, 480, 3), True), ((512, 512, 1), True), ((300, 700, 3), False), ((1080, 1920, 3), True)],
                         ids=["270x480x3", "512x512x1", "700x300x3-noise", "1080p"])
def test_device_fit_is_bit_identical_to_the_host_fit(shape, smooth):
    """fri_fit_device sums the normal equations on the device in exact integers (any order), the host fit sums
    the same integers: the solved parameters must agree bit for bit, with and without a quantizer."""
    import torch
    h, w, c = shape
    img = (smooth_image if smooth else uniform_image)(h, w, c, seed=h + 1)
    with capi.Plan(w, h, c) as plan:
        for q in (ONES, smallest_layer_q(3)):
            coefs = plan.encode(img, q)[0]
            d = torch.from_numpy(coefs).cuda()
            vp, wp = plan.fit_device(d.data_ptr())
            assert plan.last_launches == 2
            hv, hw = plan.fit_parameters(coefs)
            assert np.array_equal(vp.view(np.uint32), hv.view(np.uint32))
            assert np.array_equal(wp.view(np.uint32), hw.view(np.uint32))
            assert np.isfinite(vp).all() and np.isfinite(wp).all() and np.abs(vp).max() > 0


def test_stage_mirror_of_the_entropy_stages_on_cpu():
    """stages.prediction / entropy_coding / serialize keep the reference's names and hand-offs; on a host-only plan
    (device = -1) everything behind the quantizer runs, fed with oracle coefficients."""
    from frave_b200 import stages
    h, w, c = 96, 120, 3
    img = smooth_image(h, w, c, seed=4)
    with capi.Plan(w, h, c, device=-1) as plan:
        coefs = _coefs(plan, img)
        wi = stages.WaveletImage(stages.ImageMetadata(h, w, stages.ColorSpace.YCbCr), plan.centers(), coefs, plan.masks(), 9,
                                 np.ones(32, np.int32), 1, -1, quantized=True)
    ctx = stages.prediction.encode(wi)
    ci = stages.entropy_coding.encode(wi, ctx)
    data = stages.serialize.encode(ci)
    back = stages.serialize.decode(data)
    assert (back.metadata.height, back.metadata.width, back.metadata.colorspace) == (h, w, stages.ColorSpace.YCbCr)
    wi2 = stages.entropy_coding.decode(back, device=-1)
    assert np.array_equal(wi2.coefficients, coefs) and np.array_equal(wi2.centers, wi.centers)
    with pytest.raises(stages.StageError):
        stages.serialize.decode(b"not a frif file")


@pytest.mark.gpu
def test_fri_encoder_and_decoder_drivers():
    """The reference's two public entry points (encoder.rs:87-109, decoder.rs:48-59) on the device."""
    from frave_b200 import stages
    h, w = 270, 480
    rgb = smooth_image(h, w, 3, seed=8)
    data = stages.FRIEncoder().encode(rgb.tobytes(), h, w, stages.ColorSpace.RGB)
    out = stages.FRIDecoder().decode(data)
    assert out.metadata.colorspace is stages.ColorSpace.RGB and out.data.shape == (h, w, 3)
    with capi.Plan(w, h, 3) as plan:  # 480x270: the reference's BFS leaves five pixels uncovered, they decode to 0
        assert int((out.data != rgb).any(axis=2).sum()) == w * h - plan.pixels_covered
    luma = smooth_image(64, 96, 1, seed=9)
    back = stages.FRIDecoder().decode(stages.FRIEncoder().encode(luma, 64, 96, stages.ColorSpace.Luma))
    assert np.array_equal(back.data, luma) and back.metadata.colorspace is stages.ColorSpace.Luma
    with pytest.raises(stages.StageError) as ei:
        stages.FRIDecoder().decode(data[:100])
    assert "Failed to decode" in str(ei.value)
