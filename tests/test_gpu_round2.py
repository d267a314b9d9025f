"""GPU parity suite, second part (-m gpu): full-image comparisons at BASELINE.json's sizes, deep trees at
16384^2, the 10-bit packed transport, and the boundary's concurrency rules.  Everything goes through the C
ABI and is compared with the CPU oracle (oracle/fri_oracle.c) on the same seeded inputs."""
import threading

import numpy as np
import pytest

from frave_b200 import capi
from oracle import c_oracle as O
from tests.conftest import smallest_layer_q, smooth_image, uniform_image
from tests.test_gpu_parity import ONES, oracle_decode, oracle_encode, random_q, some_of

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------
# VERDICT r1 "Next round" 1(a)/(b): every tile of the exact bench / BASELINE configurations, both directions
# ------------------------------------------------------------------------------------------------
def _full_image_parity(w, h, c, dtype, q, seed, smooth=False):
    img = (smooth_image if smooth else uniform_image)(h, w, c, seed=seed, dtype=dtype)
    with capi.Plan(w, h, c, sample_bytes=np.dtype(dtype).itemsize) as plan:
        got = plan.encode(img, q)[0]
        want, some = oracle_encode(plan, img, ONES if q is None else q)
        assert np.array_equal(got, want), "coefficients differ from the oracle"
        rec = plan.decode(got, q)[0]
        assert np.array_equal(rec, oracle_decode(plan, got, some, ONES if q is None else q)), "pixels differ from the oracle"
        if np.dtype(dtype).itemsize == 1:  # the 16-bit transport of the same calls
            got16 = plan.encode(img, q, dtype=np.int16)[0]
            assert np.array_equal(got16, want)
            assert np.array_equal(plan.decode(got16, q)[0], rec)
        return plan.n_tiles


def test_bench_config_every_tile_4096_rgb_q4():
    """bench.py's own configuration (BASELINE.json configs[1]): 4096x4096x3 u8, q[8] = q[9] = 4."""
    assert _full_image_parity(4096, 4096, 3, np.uint8, smallest_layer_q(4), seed=2) == 33289


def test_4k_frame_every_tile():
    """One frame of BASELINE.json configs[2] (3840x2160x3), reference matrix and the smallest-layer quantizer."""
    assert _full_image_parity(3840, 2160, 3, np.uint8, None, seed=3) == 16541
    _full_image_parity(3840, 2160, 3, np.uint8, smallest_layer_q(7), seed=4, smooth=True)


def test_16384_u16_depth9_every_tile():
    """BASELINE.json configs[3] at the reference's depth: 16384x16384x1 u16, 526369 tiles."""
    assert _full_image_parity(16384, 16384, 1, np.uint16, smallest_layer_q(3), seed=4) == 526369


def test_512_gray_config0_every_tile():
    """BASELINE.json configs[0] shape: 512x512x1 u8 (smooth input S, seed 1)."""
    assert _full_image_parity(512, 512, 1, np.uint8, None, seed=1, smooth=True) == 578


# ------------------------------------------------------------------------------------------------
# 1(c): deep trees at 16384^2 — sampled fractals against the oracle, plus the oracle's uncovered-pixel set
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("depth", [16, 20, 24])
def test_deep_tree_16384_every_fractal(depth):
    """BASELINE.json configs[3] with the deep-tree extension: every coefficient of every fractal and every
    reconstructed pixel against the oracle, and the set of pixels the reference's BFS leaves uncovered."""
    w = h = 16384
    q = random_q(depth, 6)
    img = uniform_image(h, w, 1, seed=depth, dtype=np.uint16)
    with capi.Plan(w, h, 1, depth=depth, sample_bytes=2) as plan:
        centers = plan.centers()
        got = plan.encode(img, q)[0]
        assert plan.last_launches in (2, 3)  # [zero fill of base tiles outside the image +] base kernel + coarse levels
        want, some = O.extract_tiles(img, centers, depth=depth, nthreads=8)
        want = O.quantize(want, some, q, depth=depth)
        assert np.array_equal(got, want)
        assert np.array_equal(plan.masks(), some[:, 0, :])
        rec = plan.decode(got, q)[0]
        dq = O.quantize(want, some, q, depth=depth)
        del want
        assert np.array_equal(rec, O.extract_values(centers, dq, some, h, w, depth=depth, dtype=np.uint16, nthreads=8))
        # the oracle's covered-pixel set: a low-pass root of 1 with zero residues decodes to 1 on every owned pixel
        dq[...] = 0
        dq[:, :, 0] = 1
        owned = O.extract_values(centers, dq, some, h, w, depth=depth, dtype=np.uint16, nthreads=8)[:, :, 0] == 1
        del dq, some
        assert int(owned.sum()) == plan.pixels_covered
        # q == 1: the identity on covered pixels, zero elsewhere (from_wavelet's zero raster, wavelet_transform.rs:309-317)
        rec = plan.decode(plan.encode(img))[0]
        assert np.array_equal(rec[owned], img[owned])
        assert not rec[~owned].any()
        if depth == 20:  # the BFS does not reach every pixel of 16384^2 at this depth (profiles/r1_other_configs.jsonl)
            assert 0 < plan.pixels_covered < w * h
        else:
            assert plan.pixels_covered == w * h


# ------------------------------------------------------------------------------------------------
# 10-bit packed transport (fri_*_emit10)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(48, 64, 1), (131, 77, 3), (270, 480, 3), (1080, 1920, 3)], ids=lambda s: "x".join(map(str, s)))
def test_packed10_transport(shape):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    h, w, c = shape
    frames = np.stack([uniform_image(h, w, c, seed=300 + i) for i in range(3)])
    q = smallest_layer_q(2)
    with capi.Plan(w, h, c) as plan:
        cnt, nb = plan.emission_count(), plan.emission_packed_bytes()
        assert nb == 80 * ((cnt + 63) // 64)
        streams = plan.encode_emit(frames, q)                   # int32 streams (checked against the oracle elsewhere)
        packed = plan.encode_emit10(frames, q)
        assert packed.shape == (3, c, nb)
        assert np.array_equal(packed, capi.pack10(streams))     # host restatement of the format, padding = symbol 0
        assert np.array_equal(capi.unpack10(packed, cnt), streams)
        some = some_of(plan)
        coefs = plan.encode(frames, q)
        want_px = np.stack([oracle_decode(plan, coefs[i], some, q) for i in range(3)])
        assert np.array_equal(plan.decode_emit10(packed, q), want_px)
        # device-level entry points
        d_coefs = torch.from_numpy(coefs).to(dev)
        d_packed = torch.zeros((3, c, nb), dtype=torch.uint8, device=dev)
        plan.emit_device10(d_coefs.data_ptr(), 3, d_packed.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_packed.cpu().numpy(), packed)
        back = torch.full_like(d_coefs, 77)
        plan.unemit_device10(d_packed.data_ptr(), 3, back.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(back.cpu().numpy(), coefs)
        # arbitrary symbols (a decoder's input need not come from our encoder): every 10-bit value
        rng = np.random.Generator(np.random.PCG64(w))
        anys = rng.integers(-512, 512, size=streams.shape, dtype=np.int32)
        assert np.array_equal(plan.decode_emit10(capi.pack10(anys), q), plan.decode_emit(anys, q))


def test_packed10_saturates_instead_of_wrapping():
    """A coefficient outside the 1024-symbol alphabet cannot come from an 8-bit image; a q-less u8 encode stays
    within +-255.  Feed the gather directly with out-of-range values: they saturate at -512 / +511."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    with capi.Plan(96, 70, 3) as plan:
        cnt = plan.emission_count()
        gen = torch.Generator(device=dev).manual_seed(1)
        coefs = torch.randint(-3000, 3001, plan.coef_shape, generator=gen, device=dev, dtype=torch.int32)
        d_i32 = torch.empty((1, 3, cnt), dtype=torch.int32, device=dev)
        d_p = torch.zeros((1, 3, plan.emission_packed_bytes()), dtype=torch.uint8, device=dev)
        plan.emit_device(coefs.data_ptr(), 1, d_i32.data_ptr())
        plan.emit_device10(coefs.data_ptr(), 1, d_p.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(capi.unpack10(d_p.cpu().numpy(), cnt), np.clip(d_i32.cpu().numpy(), -512, 511))


# ------------------------------------------------------------------------------------------------
# boundary: one handle, both directions, asynchronous mode; pageable buffers; concurrent device calls
# ------------------------------------------------------------------------------------------------
def test_one_handle_alternating_directions_in_async_mode():
    """a decode enqueued right after an encode on the SAME handle reuses the encode's device slots;
    the copy-in of the second call must wait for the first call's device-to-host copy (and vice versa)."""
    h, w, c = 1080, 1920, 3
    q = smallest_layer_q(4)
    with capi.Plan(w, h, c) as plan:
        some = some_of(plan)
        px = capi.PinnedBuffer((2, h, w, c), np.uint8)
        cf = capi.PinnedBuffer((2,) + plan.coef_shape, np.int32)
        cf16 = capi.PinnedBuffer((2,) + plan.coef_shape, np.int16)
        cf_in = capi.PinnedBuffer((2,) + plan.coef_shape, np.int32)
        out = capi.PinnedBuffer((2, h, w, c), np.uint8)
        st_out = capi.PinnedBuffer((2, c, plan.emission_count()), np.int16)
        imgs = np.stack([uniform_image(h, w, c, seed=70 + i) for i in range(2)])
        others = np.stack([oracle_encode(plan, uniform_image(h, w, c, seed=80 + i), q)[0] for i in range(2)])
        want_cf = np.stack([oracle_encode(plan, imgs[i], q)[0] for i in range(2)])
        want_px = np.stack([oracle_decode(plan, others[i], some, q) for i in range(2)])
        px.array[...] = imgs
        cf_in.array[...] = others
        plan.set_async(True)
        for _ in range(3):
            cf.array[...] = 0
            out.array[...] = 0
            plan.encode(px.array, q, out=cf.array)          # i32 blocks, both slots
            plan.decode(cf_in.array, q, out=out.array)      # other direction, same slots
            plan.encode(px.array, q, out=cf16.array)        # 16-bit element size on the same slots
            plan.encode_emit(px.array, q, out=st_out.array)  # emission staging
            plan.sync()
            assert np.array_equal(cf.array, want_cf)
            assert np.array_equal(out.array, want_px)
            assert np.array_equal(cf16.array, want_cf)
        plan.set_async(False)
        for b in (px, cf, cf16, cf_in, out, st_out):
            b.free()


def test_async_mode_rejects_pageable_buffers():
    h, w, c = 64, 96, 3
    with capi.Plan(w, h, c) as plan:
        img = uniform_image(h, w, c, seed=1)[None]
        ok = plan.encode(img)  # pageable numpy memory is fine in the default, synchronous mode
        plan.set_async(True)
        with pytest.raises(capi.FriError) as ei:
            plan.encode(img)
        assert ei.value.code == capi.FRI_E_INVALID and "page-locked" in str(ei.value)
        with pytest.raises(capi.FriError):
            plan.decode(ok)
        pinned_in = capi.PinnedBuffer(img.shape, np.uint8)
        pinned_out = capi.PinnedBuffer(ok.shape, np.int32)
        pinned_in.array[...] = img
        plan.encode(pinned_in.array, out=pinned_out.array)
        plan.sync()
        assert np.array_equal(pinned_out.array, ok)
        plan.set_async(False)
        pinned_in.free()
        pinned_out.free()


def test_concurrent_deep_tree_device_calls_on_two_streams():
    """depth > 9 device calls used one scratch buffer per plan; now the low-pass scratch is allocated
    per call (stream-ordered), so an encode and a decode in flight on two streams do not disturb each other."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    h, w, c, depth = 700, 900, 3, 13
    q = random_q(5, 7)
    imgs = [uniform_image(h, w, c, seed=s) for s in (1, 2)]
    with capi.Plan(w, h, c, depth=depth) as plan:
        some = some_of(plan)
        want = [oracle_encode(plan, im, q)[0] for im in imgs]
        want_px = oracle_decode(plan, want[1], some, q)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        d_px = torch.from_numpy(imgs[0]).to(dev)
        d_co = torch.empty(plan.coef_shape, dtype=torch.int32, device=dev)
        d_co2 = torch.from_numpy(want[1]).to(dev)
        d_out = torch.empty((h, w, c), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        for _ in range(20):
            plan.encode_device(d_px.data_ptr(), 1, d_co.data_ptr(), q, s1.cuda_stream)
            plan.decode_device(d_co2.data_ptr(), 1, d_out.data_ptr(), q, False, s2.cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_co.cpu().numpy(), want[0])
        assert np.array_equal(d_out.cpu().numpy(), want_px)


def test_back_to_back_dependent_launches_on_one_stream():
    """Programmatic dependent launch must not break stream order: encode -> decode -> encode -> ... of the SAME
    buffers on one stream, no synchronisation in between, is still the oracle's result."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    h, w, c = 2160, 3840, 3
    img = uniform_image(h, w, c, seed=11)
    with capi.Plan(w, h, c) as plan:
        d_px = torch.from_numpy(img).to(dev)
        d_co = torch.zeros(plan.coef_shape, dtype=torch.int32, device=dev)
        d_out = torch.zeros_like(d_px)
        for _ in range(8):  # lossless chain: pixels -> coefs -> pixels' -> coefs' ...
            plan.encode_device(d_px.data_ptr(), 1, d_co.data_ptr(), None)
            plan.decode_device(d_co.data_ptr(), 1, d_out.data_ptr(), None)
            plan.encode_device(d_out.data_ptr(), 1, d_co.data_ptr(), None)
            plan.decode_device(d_co.data_ptr(), 1, d_px.data_ptr(), None)
        torch.cuda.synchronize()
        assert np.array_equal(d_px.cpu().numpy(), img)
        assert np.array_equal(d_out.cpu().numpy(), img)
        want, _ = oracle_encode(plan, img)
        assert np.array_equal(d_co.cpu().numpy(), want)


def test_encoder_and_decoder_threads_packed_transport():
    """The e2e driving pattern of bench.py: an encoder thread and a decoder thread, one handle each, 10-bit
    packed emission streams on the host side."""
    h, w, c = 1080, 1920, 3
    q = smallest_layer_q(4)
    with capi.Plan(w, h, c) as eplan, capi.Plan(w, h, c) as dplan:
        some = some_of(eplan)
        img = uniform_image(h, w, c, seed=5)[None]
        streams = eplan.encode_emit(img, q)
        packed_want = capi.pack10(streams)
        px_want = oracle_decode(eplan, eplan.encode(img, q)[0], some, q)
        res = {}

        def enc():
            for _ in range(5):
                res["packed"] = eplan.encode_emit10(img, q)

        def dec():
            for _ in range(5):
                res["px"] = dplan.decode_emit10(packed_want, q)

        th = [threading.Thread(target=enc), threading.Thread(target=dec)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert np.array_equal(res["packed"], packed_want)
        assert np.array_equal(res["px"][0], px_want)


# ------------------------------------------------------------------------------------------------
# next-2: prediction + context bucketing on the device, against oracle/fri_predict_np.py
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,smooth", [((48, 64, 1), False), ((131, 77, 3), False), ((131, 77, 3), True), ((270, 480, 1), True),
                                          ((37, 100, 3), False)],
                         ids=["48x64x1", "131x77x3", "131x77x3-smooth", "270x480x1-smooth", "37x100x3"])
def test_prediction_and_context_buckets(shape, smooth):
    """fri_predict_device: (bucket, prediction as i32, zig-zag symbol) of every emitted coefficient and the
    per-context histograms, bit for bit against the dict-based restatement of prediction.rs / context_modeling.rs,
    for fixed predictor parameters (the lstsq fit stays on the host)."""
    from oracle import fri_predict_np as PR
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    h, w, c = shape
    img = (smooth_image if smooth else uniform_image)(h, w, c, seed=h + w)
    q = smallest_layer_q(3)
    rng = np.random.Generator(np.random.PCG64(w))
    for trial in range(2):
        # trial 0: plausible parameters (a smoothing predictor); trial 1: arbitrary ones incl. negative widths
        if trial == 0:
            vp = np.tile(np.array([0.4, 0.1, 0.1, 0.2, 0.1, 0.1], np.float32), (c, 3, 1)) + rng.normal(0, 0.02, (c, 3, 6)).astype(np.float32)
            wp = np.abs(rng.normal(0.5, 0.3, (c, 3, 6))).astype(np.float32)
        else:
            vp = rng.normal(0, 1.5, (c, 3, 6)).astype(np.float32)
            wp = rng.normal(0, 2.0, (c, 3, 6)).astype(np.float32)
        with capi.Plan(w, h, c) as plan:
            coefs = plan.encode(img, q)[0]
            some = plan.masks()
            order = plan.emission_order().astype(np.int64)
            src = order[some.reshape(-1)[order]]
            cnt = plan.emission_count()
            assert len(src) == cnt
            want_b, want_p, want_s, want_h, want_o = PR.predict(plan.centers(), coefs, some, src, vp, wp)
            d_coefs = torch.from_numpy(coefs).to(dev)
            d_b = torch.full((1, c, cnt), 255, dtype=torch.uint8, device=dev)
            d_p = torch.zeros((1, c, cnt), dtype=torch.int32, device=dev)
            d_s = torch.zeros((1, c, cnt), dtype=torch.int16, device=dev)
            d_h = torch.full((1, c, 10, 1024), 7, dtype=torch.int32, device=dev)  # the call zeroes it
            d_o = torch.full((1,), 99, dtype=torch.int32, device=dev)
            plan.predict_device(d_coefs.data_ptr(), 1, vp, wp, d_b.data_ptr(), d_p.data_ptr(), d_s.data_ptr(), d_h.data_ptr(),
                                d_o.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(d_b.cpu().numpy()[0], want_b)
            assert np.array_equal(d_p.cpu().numpy()[0], want_p)
            assert np.array_equal(d_s.cpu().numpy()[0].view(np.uint16), np.minimum(want_s, 0xffff).astype(np.uint16))
            assert np.array_equal(d_h.cpu().numpy()[0].view(np.uint32), want_h)
            assert int(d_o.item()) == want_o
            if trial == 0 and smooth:
                assert want_o == 0  # a sane predictor on a smooth image stays inside the 1024-symbol alphabet
            # the symbols + predictions give the coefficients back: value = unpack_signed(symbol) + prediction
            if want_o == 0:
                back = capi.unpack10(capi.pack10(np.zeros(1, np.int32)), 1)  # (format helper sanity)
                assert back[0] == 0
                sym = want_s.astype(np.int64)
                res = np.where(sym % 2 == 0, sym // 2, -((sym + 1) // 2))
                streams = plan.encode_emit(img, q)[0]
                assert np.array_equal(res + want_p, streams)


def test_independent_calls_hint_overlaps_kernels_without_changing_results():
    """fri_plan_set_independent_calls: a stream of frames in disjoint buffers, no dependency wait between the
    launches — every frame still equals the oracle; with the hint off again, dependent chains work as before."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    h, w, c, n = 1080, 1920, 3, 6
    q = smallest_layer_q(4)
    imgs = [uniform_image(h, w, c, seed=500 + i) for i in range(n)]
    with capi.Plan(w, h, c) as plan:
        some = some_of(plan)
        want = [oracle_encode(plan, im, q)[0] for im in imgs]
        d_px = [torch.from_numpy(im).to(dev) for im in imgs]
        d_co = [torch.zeros(plan.coef_shape, dtype=torch.int32, device=dev) for _ in range(n)]
        d_in = [torch.from_numpy(x).to(dev) for x in want]
        d_out = [torch.zeros((h, w, c), dtype=torch.uint8, device=dev) for _ in range(n)]
        torch.cuda.synchronize()
        plan.set_independent_calls(True)
        for rep in range(3):
            for i in range(n):  # encode frame i and decode frame i's reference coefficients: all buffers distinct
                plan.encode_device(d_px[i].data_ptr(), 1, d_co[i].data_ptr(), q)
                plan.decode_device(d_in[i].data_ptr(), 1, d_out[i].data_ptr(), q)
        torch.cuda.synchronize()
        plan.set_independent_calls(False)
        for i in range(n):
            assert np.array_equal(d_co[i].cpu().numpy(), want[i])
            assert np.array_equal(d_out[i].cpu().numpy(), oracle_decode(plan, want[i], some, q))
        plan.encode_device(d_px[0].data_ptr(), 1, d_co[1].data_ptr(), None)   # dependent chain, hint off
        plan.decode_device(d_co[1].data_ptr(), 1, d_out[1].data_ptr(), None)
        torch.cuda.synchronize()
        assert np.array_equal(d_out[1].cpu().numpy(), imgs[0])


@pytest.mark.parametrize("bits", [9, 10])
@pytest.mark.parametrize("shape", [(131, 77, 3), (512, 512, 1), (1080, 1920, 3)], ids=lambda s: "x".join(map(str, s)))
def test_packed_transport_at_9_and_10_bits(shape, bits):
    """fri_*_packed: the packed layout at 9 bits (everything an 8-bit image produces: |k| <= 255) and 10 bits."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    h, w, c = shape
    frames = np.stack([uniform_image(h, w, c, seed=700 + i) for i in range(2)])
    q = smallest_layer_q(2)
    with capi.Plan(w, h, c) as plan:
        cnt, nb = plan.emission_count(), plan.emission_packed_size(bits)
        assert nb == 8 * bits * ((cnt + 63) // 64)
        streams = plan.encode_emit(frames, None)  # q == 1: residues span the whole +-255 range
        assert np.abs(streams).max() <= 255
        packed = plan.encode_emit_packed(frames, None, bits)
        assert np.array_equal(packed, capi.pack_bits(streams, bits))
        assert np.array_equal(capi.unpack_bits(packed, cnt, bits), streams)  # nothing saturates
        assert np.array_equal(plan.decode_emit_packed(packed, None, bits), frames if plan.pixels_covered == w * h else plan.decode_emit(streams))
        sq = plan.encode_emit(frames, q)
        assert np.array_equal(plan.decode_emit_packed(plan.encode_emit_packed(frames, q, bits), q, bits), plan.decode_emit(sq, q))
        d_coefs = torch.from_numpy(plan.encode(frames)).to(dev)
        d_p = torch.zeros((2, c, nb), dtype=torch.uint8, device=dev)
        plan.emit_device_packed(d_coefs.data_ptr(), 2, bits, d_p.data_ptr())
        back = torch.full_like(d_coefs, 3)
        plan.unemit_device_packed(d_p.data_ptr(), 2, bits, back.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_p.cpu().numpy(), packed) and torch.equal(back, d_coefs)
        with pytest.raises(capi.FriError) as ei:
            plan.encode_emit_packed(frames, None, 8)
        assert ei.value.code == capi.FRI_E_INVALID


# ------------------------------------------------------------------------------------------------
# one image split over several GPUs by ranges of tile groups (here: the parts run one after the
# other on one GPU, each with its own band buffers, exactly as N ranks would hold them)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,dtype,n_parts", [((1080, 1920, 3), np.uint8, 2), ((1080, 1920, 3), np.uint8, 8), ((517, 333, 3), np.uint8, 3),
                                                 ((700, 500, 1), np.uint16, 4), ((2048, 2048, 1), np.uint8, 5)],
                         ids=["1080p/2", "1080p/8", "333x517x3/3", "500x700 u16/4", "2048x2048x1/5"])
def test_image_split_by_tile_group_ranges(shape, dtype, n_parts):
    import torch
    from frave_b200 import sharding
    h, w, c = shape
    sb = np.dtype(dtype).itemsize
    img = uniform_image(h, w, c, seed=h + n_parts, dtype=dtype)
    q = smallest_layer_q(3)
    tdt = torch.uint8 if sb == 1 else torch.int16
    with capi.Plan(w, h, c, sample_bytes=sb) as plan:
        want = plan.encode(img, q)[0]
        full = plan.decode(want, q)[0]
        merged = torch.zeros((h, w, c), dtype=torch.int32, device="cuda")
        seen_tiles = 0
        for r in range(n_parts):
            p = sharding.shard_image(plan, r, n_parts)
            t0, t1, r0, r1 = p["tile_begin"], p["tile_end"], p["row_begin"], p["row_end"]
            band = torch.from_numpy(img[r0:r1].view(np.uint8 if sb == 1 else np.int16).copy()).cuda()  # only the rows the part touches
            coefs = torch.full((t1 - t0, c, 512), 123456, dtype=torch.int32, device="cuda")
            plan.encode_device_part(band.data_ptr(), coefs.data_ptr(), r, n_parts, q)
            assert plan.last_launches == 1
            assert np.array_equal(coefs.cpu().numpy(), want[t0:t1]), f"part {r}: coefficients differ from the whole-frame call"
            out = torch.zeros((r1 - r0, w, c), dtype=tdt, device="cuda")
            plan.decode_device_part(coefs.data_ptr(), out.data_ptr(), r, n_parts, q)
            o32 = out.to(torch.int32) & (0xFFFF if sb == 2 else 0xFF)
            assert not ((merged[r0:r1] != 0) & (o32 != 0)).any(), "two parts wrote the same pixel"
            merged[r0:r1] += o32  # the exchange step: overlap rows merge by addition
            seen_tiles += t1 - t0
        assert seen_tiles == plan.n_tiles
        assert np.array_equal(merged.cpu().numpy().astype(dtype), full)
        # the same exchange done by the kernel's own stores (what N ranks do over peer-mapped bands): every part
        # decodes its tiles into its own band, then re-runs its groups along each cut into the neighbour's band;
        # no zeroing, no merge — bands start as garbage and must end up complete
        if plan.pixels_covered == w * h:
            margin = plan.launch_info()["region_h"]
            parts = [sharding.shard_image(plan, r, n_parts) for r in range(n_parts)]
            dcoefs = torch.from_numpy(want).cuda()
            bands = [torch.full((p["row_end"] - p["row_begin"] + 2 * margin, w, c), 77, dtype=tdt, device="cuda") for p in parts]
            stride = w * c * sb
            for r, p in enumerate(parts):
                own = bands[r].data_ptr() + margin * stride
                plan.decode_device_part(dcoefs[p["tile_begin"]:].data_ptr(), own, r, n_parts, q)
                for push in sharding.halo_pushes(plan, r, n_parts):
                    peer = parts[push["peer"]]
                    assert peer["row_begin"] - margin <= push["span_begin"] and push["span_end"] <= peer["row_end"] + margin
                    plan.decode_device_groups(dcoefs[p["tile_begin"]:].data_ptr(), p["tile_begin"], bands[push["peer"]].data_ptr(),
                                              peer["row_begin"] - margin, push["first"], push["last"], q)
            for r, p in enumerate(parts):
                got = bands[r][margin:margin + p["row_end"] - p["row_begin"]].cpu().numpy().view(dtype)
                assert np.array_equal(got, full[p["row_begin"]:p["row_end"]]), f"band {r} is incomplete after the pushes"
        with pytest.raises(capi.FriError):
            plan.encode_device_part(band.data_ptr(), coefs.data_ptr(), n_parts, n_parts, q)
