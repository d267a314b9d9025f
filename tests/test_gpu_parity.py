"""GPU parity suite (-m gpu): libfri_cuda, called through its C ABI, against the CPU oracle on
the same seeded inputs — bit-exact int32 coefficients and bit-exact reconstructed pixels — plus
size-independent properties at BASELINE.json's full sizes."""
import os

import numpy as np
import pytest

from frave_b200 import capi, stages
from oracle import c_oracle as O
from tests.conftest import smallest_layer_q, smooth_image, uniform_image

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ONES = np.ones(32, np.int32)


def oracle_encode(plan, img, q=ONES, nthreads=8):
    """Oracle coefficients in the plan's tile order; None -> 0 exactly like the GPU output."""
    coef, some = O.extract_tiles(img, plan.centers(), depth=plan.depth, nthreads=nthreads)
    return O.quantize(coef, some, q, depth=plan.depth), some


def oracle_decode(plan, coefs, some, q=ONES, multiply=False, nthreads=8):
    dq = O.quantize(coefs, some, q, depth=plan.depth, multiply=multiply)
    return O.extract_values(plan.centers(), dq, some, plan.height, plan.width, depth=plan.depth,
                            dtype=plan.pixel_dtype, nthreads=nthreads)


def some_of(plan):
    m = plan.masks()
    return np.broadcast_to(m[:, None, :], plan.coef_shape).copy()


def random_q(seed, hi=64):
    return np.random.Generator(np.random.PCG64(seed)).integers(1, hi + 1, size=32).astype(np.int32)


SHAPES = [(1, 1, 1), (10, 10, 3), (48, 64, 1), (37, 100, 3), (300, 7, 3), (131, 77, 3), (512, 512, 1), (257, 1031, 3),
          (64, 4096, 1), (33, 65, 1)]


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16], ids=["u8", "u16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_encode_bit_exact(shape, dtype):
    h, w, c = shape
    img = uniform_image(h, w, c, seed=h * 7 + w, dtype=dtype)
    with capi.Plan(w, h, c, sample_bytes=img.itemsize) as plan:
        for q in (None, smallest_layer_q(5), random_q(h + w), np.full(32, 70000, np.int32)):
            got = plan.encode(img, q)[0]
            want, some = oracle_encode(plan, img, ONES if q is None else q)
            assert np.array_equal(got, want), f"q={None if q is None else q[:10]}"
            assert not got[~some].any()  # None comes out as 0
        assert plan.last_launches == 1


@pytest.mark.parametrize("bulk", ["0", "1"])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16], ids=["u8", "u16"])
@pytest.mark.parametrize("shape", [(131, 77, 3), (512, 512, 1), (257, 1031, 3), (700, 900, 3), (1080, 1920, 3)],
                         ids=lambda s: "x".join(map(str, s)))
def test_encoder_staging_paths_are_bit_exact(monkeypatch, shape, dtype, bulk):
    """The encoder stages interior groups either with 16-byte cp.async driven by the chunk list or with one
    bulk copy (TMA) per row, chosen by launch size; FRI_STAGE_BULK forces either on every shape."""
    monkeypatch.setenv("FRI_STAGE_BULK", bulk)
    h, w, c = shape
    frames = np.stack([uniform_image(h, w, c, seed=h + i, dtype=dtype) for i in range(2)])
    with capi.Plan(w, h, c, sample_bytes=frames.itemsize) as plan:
        for q in (None, smallest_layer_q(6), random_q(w, hi=20)):
            got = plan.encode(frames, q)
            for f in range(2):
                want, _ = oracle_encode(plan, frames[f], ONES if q is None else q)
                assert np.array_equal(got[f], want)
        if frames.itemsize == 1:
            got16 = plan.encode(frames, smallest_layer_q(6), dtype=np.int16)
            assert np.array_equal(got16[1], oracle_encode(plan, frames[1], smallest_layer_q(6))[0])


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16], ids=["u8", "u16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_decode_bit_exact(shape, dtype):
    h, w, c = shape
    rng = np.random.Generator(np.random.PCG64(h * 11 + w))
    with capi.Plan(w, h, c, sample_bytes=np.dtype(dtype).itemsize) as plan:
        some = some_of(plan)
        hi = 300 if dtype == np.uint8 else 70000
        coefs = rng.integers(-hi, hi + 1, size=plan.coef_shape).astype(np.int32)
        coefs[:, :, 0] = rng.integers(-50, hi + 50, size=plan.coef_shape[:2])  # DC mostly in range, clamp exercised
        for q, multiply in ((None, False), (random_q(w, 5), False), (random_q(h, 3), True)):
            got = plan.decode(coefs, q, multiply=multiply)[0]
            want = oracle_decode(plan, coefs, some, ONES if q is None else q, multiply)
            assert np.array_equal(got, want)
        # coefficients the reference holds as None are ignored, whatever the buffer contains
        noisy = coefs.copy()
        noisy[~some] = rng.integers(-2**31, 2**31 - 1, size=int((~some).sum()))
        assert np.array_equal(plan.decode(noisy)[0], oracle_decode(plan, coefs, some))
        # wrapping i32 arithmetic (release-mode Rust) on extreme coefficients
        wild = rng.integers(-2**31, 2**31 - 1, size=plan.coef_shape).astype(np.int32)
        assert np.array_equal(plan.decode(wild, random_q(3, 4))[0], oracle_decode(plan, wild, some, random_q(3, 4)))


@pytest.mark.parametrize("name", ["u8_rgb_70x96_q5", "u8_luma_48x64_q1", "u16_luma_65x33_q3"])
def test_golden_fixtures(name):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    img = fx["pixels"]
    h, w, c = img.shape
    with capi.Plan(w, h, c, sample_bytes=img.itemsize) as plan:
        order = {tuple(x): i for i, x in enumerate(fx["centers"].tolist())}
        idx = np.array([order[tuple(x)] for x in plan.centers().tolist()])
        got = plan.encode(img, fx["q"])[0]
        assert np.array_equal(got, fx["coef"][idx])
        assert np.array_equal(plan.mask_words().view(np.uint8), fx["some"][idx])
        assert np.array_equal(plan.decode(got, fx["q"])[0], fx["recon"])


@pytest.mark.parametrize("depth", [10, 11, 12, 13, 15, 16, 17, 19])
@pytest.mark.parametrize("dtype,c", [(np.uint8, 1), (np.uint8, 3), (np.uint16, 1)], ids=["u8x1", "u8x3", "u16x1"])
def test_deep_tree_extension(depth, dtype, c):
    # odd sub_bits = depth - 9 and even ones: the coarse inverse kernel ping-pongs between two shared-memory
    # buffers of different sizes and must end in the large one whatever the parity (depths 13/15/17/19)
    h, w = (150, 260) if depth < 15 else (420, 610)
    img = uniform_image(h, w, c, seed=depth, dtype=dtype)
    with capi.Plan(w, h, c, depth=depth, sample_bytes=img.itemsize) as plan:
        q = random_q(depth, 9)
        got = plan.encode(img, q)[0]
        want, some = oracle_encode(plan, img, q)
        assert np.array_equal(got, want)
        assert plan.last_launches in (2, 3)  # [zero fill of base tiles outside the image +] base kernel + coarse levels
        assert np.array_equal(plan.decode(got, q)[0], oracle_decode(plan, got, some, q))
        lossless = plan.encode(img)[0]
        rec = plan.decode(lossless)[0]
        assert np.array_equal(rec, oracle_decode(plan, lossless, some))
        if plan.pixels_covered == w * h:  # the BFS does not reach every pixel for every (size, depth)
            assert np.array_equal(rec, img)


def test_batch_equals_per_frame_and_unaligned_frames():
    # 5 frames of 37x100x3: frame_bytes = 11100 (not a multiple of 16), row stride 300
    h, w, c, n = 37, 100, 3, 5
    frames = np.stack([uniform_image(h, w, c, seed=50 + i) for i in range(n)])
    with capi.Plan(w, h, c) as plan:
        q = smallest_layer_q(3)
        batch = plan.encode(frames, q)
        for i in range(n):
            assert np.array_equal(batch[i], oracle_encode(plan, frames[i], q)[0])
        rec = plan.decode(plan.encode(frames))
        assert np.array_equal(rec, frames)


@pytest.mark.parametrize("bands", [1, 3, 8])
@pytest.mark.parametrize("shape", [(270, 480, 3), (700, 900, 3), (513, 1000, 1)], ids=lambda s: "x".join(map(str, s)))
def test_banded_host_pipeline(monkeypatch, shape, bands):
    """The host-buffer entry points stream a frame through the device in bands of groups (copies in,
    kernels, copies out on three event-chained streams); any band count and several frames in
    flight must give the same bytes."""
    monkeypatch.setenv("FRI_BANDS", str(bands))
    h, w, c = shape
    frames = np.stack([uniform_image(h, w, c, seed=200 + i) for i in range(5)])
    q = smallest_layer_q(6)
    with capi.Plan(w, h, c) as plan:
        some = some_of(plan)
        coefs = plan.encode(frames, q)
        rec = plan.decode(coefs, q)
        for i in range(5):
            want, _ = oracle_encode(plan, frames[i], q)
            assert np.array_equal(coefs[i], want)
            assert np.array_equal(rec[i], oracle_decode(plan, want, some, q))
        assert np.array_equal(plan.decode(plan.encode(frames)), frames) or plan.pixels_covered < w * h


@pytest.mark.parametrize("bands", [1, 4])
@pytest.mark.parametrize("shape", [(270, 480, 3), (513, 1000, 1), (1080, 1920, 3)], ids=lambda s: "x".join(map(str, s)))
def test_16bit_transport_equals_32bit(monkeypatch, shape, bands):
    """fri_encode_tq16 / fri_decode_tq16: same coefficients and pixels as the i32 entry points (and
    the oracle), the host side of the copies being int16."""
    monkeypatch.setenv("FRI_BANDS", str(bands))
    h, w, c = shape
    frames = np.stack([uniform_image(h, w, c, seed=900 + i) for i in range(3)])
    with capi.Plan(w, h, c) as plan:
        some = some_of(plan)
        for q in (None, smallest_layer_q(3), random_q(h, hi=9)):
            c16 = plan.encode(frames, q, dtype=np.int16)
            assert c16.dtype == np.int16
            for i in range(3):
                want, _ = oracle_encode(plan, frames[i], ONES if q is None else q)
                assert np.array_equal(c16[i], want)
            rec = plan.decode(c16, q)
            for i in range(3):
                assert np.array_equal(rec[i], oracle_decode(plan, c16[i].astype(np.int32), some, ONES if q is None else q))
        # arbitrary i16 input (not produced by an encoder), both dequantizers
        rng = np.random.Generator(np.random.PCG64(h))
        anyc = rng.integers(-32768, 32768, size=(1,) + plan.coef_shape, dtype=np.int16)
        q = random_q(w, hi=5)
        for mul in (False, True):
            assert np.array_equal(plan.decode(anyc, q, multiply=mul)[0],
                                  oracle_decode(plan, anyc[0].astype(np.int32), some, q, multiply=mul))


def test_encoder_and_decoder_threads_with_one_handle_each():
    """The ABI is re-entrant across handles: an encoder thread and a decoder thread,
    one plan each on the same device, run concurrently (this is how bench.py's e2e drives the
    library) and both keep producing the oracle's bytes."""
    import threading

    h, w, c = 540, 960, 3
    frames = [uniform_image(h, w, c, seed=700 + i) for i in range(4)]
    q = smallest_layer_q(4)
    with capi.Plan(w, h, c) as eplan, capi.Plan(w, h, c) as dplan:
        some = some_of(eplan)
        want = [oracle_encode(eplan, f, q)[0] for f in frames]
        want_px = [oracle_decode(dplan, wc, some, q) for wc in want]
        eplan.set_bands(1)
        dplan.set_bands(3)
        errors = []

        def enc():
            try:
                for it in range(12):
                    i = it % 4
                    dt = np.int16 if it & 1 else np.int32
                    assert np.array_equal(eplan.encode(frames[i], q, dtype=dt)[0], want[i]), f"encode {it}"
            except Exception as exc:  # noqa: BLE001
                errors.append(exc)

        def dec():
            try:
                for it in range(12):
                    i = (it + 1) % 4
                    cf = want[i].astype(np.int16) if it & 1 else want[i]
                    assert np.array_equal(dplan.decode(cf, q)[0], want_px[i]), f"decode {it}"
            except Exception as exc:  # noqa: BLE001
                errors.append(exc)

        th = [threading.Thread(target=enc), threading.Thread(target=dec)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errors, errors


@pytest.mark.parametrize("seed", [20261018, 7, 424242])
def test_random_shapes_fuzz(seed):
    """Seeded fuzz over image shapes, channel counts, sample sizes, frame counts and quantization matrices
    (all three quantization classes): encode and decode bit for bit against the oracle, the 16-bit
    transport and — where the reference's scan exists — the emitted streams against the block path."""
    rng = np.random.Generator(np.random.PCG64(seed))
    checked_emit = 0
    for case in range(36):
        h = int(rng.integers(1, 420)) if case % 3 else int(rng.integers(1, 40))
        w = int(rng.integers(1, 640)) if case % 4 else int(rng.integers(1, 40))
        c = int(rng.choice([1, 3]))
        dtype = np.uint16 if case % 5 == 4 else np.uint8
        n = int(rng.integers(1, 4))
        frames = np.stack([uniform_image(h, w, c, seed=seed % 1000 + 5000 + 10 * case + i, dtype=dtype) for i in range(n)])
        kind = case % 3
        q = None if kind == 0 else smallest_layer_q(int(rng.integers(2, 65))) if kind == 1 else random_q(case, hi=int(rng.integers(2, 300)))
        qq = ONES if q is None else q
        tag = f"case {case}: {h}x{w}x{c} {np.dtype(dtype).name} n={n} kind={kind}"
        with capi.Plan(w, h, c, sample_bytes=frames.itemsize) as plan:
            some = some_of(plan)
            got = plan.encode(frames, q)
            rec = plan.decode(got, q, multiply=bool(case & 1))
            for f in range(n):
                want, _ = oracle_encode(plan, frames[f], qq)
                assert np.array_equal(got[f], want), tag
                assert np.array_equal(rec[f], oracle_decode(plan, want, some, qq, multiply=bool(case & 1))), tag
            if dtype == np.uint8:
                g16 = plan.encode(frames, q, dtype=np.int16)
                assert np.array_equal(g16, got), tag
                assert np.array_equal(plan.decode(g16, q, multiply=bool(case & 1)), rec), tag
                try:
                    cnt = plan.emission_count()
                except capi.FriError as e:  # sizes on which the reference's own scan asserts
                    assert e.code == capi.FRI_E_UNSUPPORTED, tag
                    continue
                order = plan.emission_order().astype(np.int64)
                src = order[plan.masks().reshape(-1)[order]]
                assert len(src) == cnt, tag
                streams = plan.encode_emit(frames, q)
                for f in range(n):
                    for ch in range(c):
                        assert np.array_equal(streams[f, ch], got[f][:, ch, :].reshape(-1)[src]), tag
                assert np.array_equal(plan.decode_emit(streams.astype(np.int16), q, multiply=bool(case & 1)), rec), tag
                checked_emit += 1
    assert checked_emit >= 10


def test_asynchronous_mode_of_the_host_entry_points():
    """fri_plan_set_async / fri_plan_sync: calls return after enqueueing on pinned buffers; one thread
    drives an encoder and a decoder handle at once and gets the oracle's bytes after the sync."""
    h, w, c = 540, 960, 3
    q = smallest_layer_q(4)
    with capi.Plan(w, h, c) as eplan, capi.Plan(w, h, c) as dplan:
        some = some_of(eplan)
        px = capi.PinnedBuffer((1, h, w, c), np.uint8)
        cf = capi.PinnedBuffer((1,) + eplan.coef_shape, np.int16)
        cf_in = capi.PinnedBuffer((1,) + eplan.coef_shape, np.int16)
        out = capi.PinnedBuffer((1, h, w, c), np.uint8)
        eplan.set_async(True)
        dplan.set_async(True)
        eplan.sync()  # nothing enqueued yet: a no-op
        for it in range(4):
            img = uniform_image(h, w, c, seed=800 + it)
            other = oracle_encode(eplan, uniform_image(h, w, c, seed=900 + it), q)[0]
            px.array[0] = img
            cf_in.array[0] = other
            eplan.encode(px.array, q, out=cf.array)
            dplan.decode(cf_in.array, q, out=out.array)
            eplan.sync()
            dplan.sync()
            assert np.array_equal(cf.array[0], oracle_encode(eplan, img, q)[0])
            assert np.array_equal(out.array[0], oracle_decode(dplan, other, some, q))
        eplan.set_async(False)
        assert np.array_equal(eplan.encode(px.array, q)[0], cf.array[0])  # back to synchronous calls
        for b in (px, cf, cf_in, out):
            b.free()


def test_more_frames_than_one_grid_dimension():
    """n_frames above the 65535 gridDim.y limit is split over several launches."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    h, w, c, n = 5, 7, 1, 65535 + 9
    with capi.Plan(w, h, c) as plan:
        gen = torch.Generator(device=dev).manual_seed(11)
        px = torch.randint(0, 256, (n, h, w, c), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)
        coefs = torch.empty((n,) + plan.coef_shape, dtype=torch.int32, device=dev)
        out = torch.empty_like(px)
        q = smallest_layer_q(3)
        plan.encode_device(px.data_ptr(), n, coefs.data_ptr(), q)
        assert plan.last_launches == 2
        plan.decode_device(coefs.data_ptr(), n, out.data_ptr(), q)
        torch.cuda.synchronize()
        some = some_of(plan)
        for f in (0, 65534, 65535, n - 1):
            want, _ = oracle_encode(plan, px[f].cpu().numpy(), q)
            assert np.array_equal(coefs[f].cpu().numpy(), want)
            assert np.array_equal(out[f].cpu().numpy(), oracle_decode(plan, want, some, q))
        plan.encode_device(px.data_ptr(), n, coefs.data_ptr())
        plan.decode_device(coefs.data_ptr(), n, out.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(px, out) or plan.pixels_covered < w * h


@pytest.mark.parametrize("shape", [(37, 100, 3), (257, 1031, 3), (512, 512, 1), (1080, 1920, 3)], ids=lambda s: "x".join(map(str, s)))
def test_int16_device_arrays(shape):
    """fri_encode_tq_device16 / fri_decode_tq_device16: the kernels on int16 coefficient arrays (3 B per
    sample) against the int32 kernels and the oracle, every quantization class, both dequantizers."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    h, w, c = shape
    frames = np.stack([uniform_image(h, w, c, seed=300 + i) for i in range(2)])
    with capi.Plan(w, h, c) as plan:
        some = some_of(plan)
        px = torch.from_numpy(frames).to(dev)
        c16 = torch.empty((2,) + plan.coef_shape, dtype=torch.int16, device=dev)
        out = torch.empty_like(px)
        for q in (None, smallest_layer_q(4), smallest_layer_q(7), random_q(h, hi=12)):
            qq = ONES if q is None else q
            c16.fill_(-3)
            plan.encode_device(px.data_ptr(), 2, c16.data_ptr(), q, half=True)
            assert plan.last_launches == 1
            torch.cuda.synchronize()
            got = c16.cpu().numpy()
            for f in range(2):
                want, _ = oracle_encode(plan, frames[f], qq)
                assert np.array_equal(got[f], want)
            plan.decode_device(c16.data_ptr(), 2, out.data_ptr(), q, half=True)
            torch.cuda.synchronize()
            for f in range(2):
                assert np.array_equal(out[f].cpu().numpy(), oracle_decode(plan, got[f].astype(np.int32), some, qq))
        rng = np.random.Generator(np.random.PCG64(w))
        anyc = rng.integers(-32768, 32768, size=(2,) + plan.coef_shape, dtype=np.int16)
        c16.copy_(torch.from_numpy(anyc))
        q = random_q(h + 1, hi=6)
        for mul in (False, True):
            plan.decode_device(c16.data_ptr(), 2, out.data_ptr(), q, multiply=mul, half=True)
            torch.cuda.synchronize()
            for f in range(2):
                assert np.array_equal(out[f].cpu().numpy(), oracle_decode(plan, anyc[f].astype(np.int32), some, q, multiply=mul))
    with capi.Plan(64, 48, 1, depth=12) as deep:  # int16 arrays: depth 9 only
        with pytest.raises(capi.FriError) as e:
            deep.encode_device(px.data_ptr(), 1, c16.data_ptr(), half=True)
        assert e.value.code == capi.FRI_E_UNSUPPORTED


def test_int16_and_emission_kernels_stay_inside_their_buffers():
    """Guard bands around every output of the int16 / emission kernels stay untouched (the pool has no
    compute-sanitizer): outputs are carved out of larger sentinel-filled allocations."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    h, w, c, n = 131, 517, 3, 3
    G = 4096  # guard elements on each side (keeps 16-byte alignment for every element size)

    def guarded(numel, dtype, fill):
        buf = torch.full((numel + 2 * G,), fill, dtype=dtype, device=dev)
        return buf, buf[G:G + numel]

    def intact(buf, fill):
        return bool((buf[:G] == fill).all()) and bool((buf[-G:] == fill).all())

    with capi.Plan(w, h, c) as plan:
        px = torch.from_numpy(np.stack([uniform_image(h, w, c, seed=60 + i) for i in range(n)])).to(dev)
        ncoef = n * plan.coefs_per_frame
        cnt = plan.emission_count()
        q = smallest_layer_q(3)
        b16, c16 = guarded(ncoef, torch.int16, 0x5A5A)
        plan.encode_device(px.data_ptr(), n, c16.data_ptr(), q, half=True)
        bpx, opx = guarded(px.numel(), torch.uint8, 0xA5)
        plan.decode_device(c16.data_ptr(), n, opx.data_ptr(), q, half=True)
        b32, c32 = guarded(ncoef, torch.int32, 0x5A5A5A5A)
        plan.encode_device(px.data_ptr(), n, c32.data_ptr(), q)
        be, e32 = guarded(n * c * cnt, torch.int32, 0x11111111)
        plan.emit_device(c32.data_ptr(), n, e32.data_ptr())
        be16, e16 = guarded(n * c * cnt, torch.int16, 0x1111)
        plan.emit_device(c32.data_ptr(), n, e16.data_ptr(), half=True)
        bu, u32 = guarded(ncoef, torch.int32, 0x22222222)
        plan.unemit_device(e32.data_ptr(), n, u32.data_ptr())
        bu2, u32b = guarded(ncoef, torch.int32, 0x22222222)
        plan.unemit_device(e16.data_ptr(), n, u32b.data_ptr(), half=True)
        torch.cuda.synchronize()
        assert intact(b16, 0x5A5A) and intact(bpx, 0xA5) and intact(b32, 0x5A5A5A5A)
        assert intact(be, 0x11111111) and intact(be16, 0x1111) and intact(bu, 0x22222222) and intact(bu2, 0x22222222)
        assert torch.equal(c16.to(torch.int32), c32) and torch.equal(e16.to(torch.int32), e32)
        assert torch.equal(u32, c32) and torch.equal(u32b, c32)
        want = torch.empty_like(px)
        plan.decode_device(c32.data_ptr(), n, want.data_ptr(), q)
        torch.cuda.synchronize()
        assert torch.equal(opx.view(n, h, w, c), want)


def test_16bit_transport_rejects_16bit_samples():
    with capi.Plan(64, 48, 1, sample_bytes=2) as plan:
        with pytest.raises(capi.FriError) as e:
            plan.encode(np.zeros((48, 64, 1), np.uint16), dtype=np.int16)
        assert e.value.code == capi.FRI_E_UNSUPPORTED


def test_device_entry_points_with_misaligned_pixel_pointer():
    torch = pytest.importorskip("torch")
    h, w, c, n = 61, 93, 3, 3
    frames = np.stack([uniform_image(h, w, c, seed=70 + i) for i in range(n)])
    with capi.Plan(w, h, c) as plan:
        dev = torch.device("cuda", 0)
        for shift in (0, 1, 7):
            raw = torch.zeros(frames.size + 64, dtype=torch.uint8, device=dev)
            raw[shift:shift + frames.size] = torch.from_numpy(frames.reshape(-1)).to(dev)
            d_coefs = torch.empty((n,) + plan.coef_shape, dtype=torch.int32, device=dev)
            plan.encode_device(raw.data_ptr() + shift, n, d_coefs.data_ptr())
            torch.cuda.synchronize()
            got = d_coefs.cpu().numpy()
            for i in range(n):
                assert np.array_equal(got[i], oracle_encode(plan, frames[i])[0])
            out = torch.full((frames.size + 64,), 0xAB, dtype=torch.uint8, device=dev)
            plan.decode_device(d_coefs.data_ptr(), n, out.data_ptr() + shift)
            torch.cuda.synchronize()
            o = out.cpu().numpy()
            assert np.array_equal(o[shift:shift + frames.size].reshape(frames.shape), frames)
            assert (o[:shift] == 0xAB).all() and (o[shift + frames.size:] == 0xAB).all()  # no stray writes


def test_uncovered_pixels_are_zeroed_like_from_wavelet():
    # 7x300: the reference's BFS does not reach every pixel (pixels_covered < W*H)
    h, w, c = 300, 7, 3
    img = uniform_image(h, w, c, seed=9)
    with capi.Plan(w, h, c) as plan:
        assert plan.pixels_covered < w * h
        coefs = plan.encode(img)
        want = oracle_decode(plan, coefs[0], some_of(plan))
        out = np.full((1, h, w, c), 0x5A, np.uint8)
        plan.decode(coefs, out=out)
        assert np.array_equal(out[0], want)


def test_quantization_sweep_1080p():
    """BASELINE.json configs[4]: smallest-layer divisor 1..64 on 1920x1080 RGB, coefficient for
    coefficient against the oracle (container bytes are out of scope: SURVEY.md §8(c))."""
    h, w, c = 1080, 1920, 3
    img = smooth_image(h, w, c, seed=5)
    with capi.Plan(w, h, c) as plan:
        base, some = O.extract_tiles(img, plan.centers(), nthreads=8)
        for d in range(1, 65):
            for both in (True, False):
                q = smallest_layer_q(d, both)
                got = plan.encode(img, q)[0]
                assert np.array_equal(got, O.quantize(base, some, q)), (d, both)
        q = smallest_layer_q(16)
        coefs = plan.encode(img, q)
        assert np.array_equal(plan.decode(coefs, q)[0], oracle_decode(plan, coefs[0], some, q))
        assert np.array_equal(plan.decode(coefs, q, multiply=True)[0], oracle_decode(plan, coefs[0], some, q, True))


def _device_roundtrip(torch, w, h, c, n_frames, dtype, seed, sample_tiles=64):
    """Full-size check on device-resident data: encode -> decode is the identity at q == 1, the
    buffers' guard bands stay untouched, and a random sample of tiles matches the oracle."""
    tdtype = torch.uint8 if dtype == np.uint8 else torch.uint16
    dev = torch.device("cuda", 0)
    with capi.Plan(w, h, c, sample_bytes=np.dtype(dtype).itemsize) as plan:
        gen = torch.Generator(device=dev).manual_seed(seed)
        hi = 256 if dtype == np.uint8 else 65536
        px = torch.randint(0, hi, (n_frames, h, w, c), generator=gen, device=dev, dtype=torch.int32).to(tdtype)
        coefs = torch.empty((n_frames,) + plan.coef_shape, dtype=torch.int32, device=dev)
        out = torch.empty_like(px)
        plan.encode_device(px.data_ptr(), n_frames, coefs.data_ptr())
        plan.decode_device(coefs.data_ptr(), n_frames, out.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(px.view(torch.uint8), out.view(torch.uint8))
        rng = np.random.Generator(np.random.PCG64(seed))
        f = int(rng.integers(0, n_frames))
        tiles = np.sort(rng.choice(plan.n_tiles, size=min(sample_tiles, plan.n_tiles), replace=False))
        img = px[f].cpu().numpy() if dtype == np.uint8 else px[f].view(torch.int16).cpu().numpy().view(np.uint16)
        want, _ = O.extract_tiles(img, plan.centers()[tiles], nthreads=8)
        got = coefs[f][torch.from_numpy(tiles).to(dev)].cpu().numpy()
        assert np.array_equal(got, want)
        # the DC of a fully covered tile is bounded by the sample range; checksum of checksums
        assert int(coefs[:, :, :, 0].min()) >= 0 and int(coefs[:, :, :, 0].max()) < hi


def test_full_size_4096_rgb_roundtrip():
    torch = pytest.importorskip("torch")
    _device_roundtrip(torch, 4096, 4096, 3, 1, np.uint8, seed=2)


def test_full_size_4k_batch_roundtrip():
    torch = pytest.importorskip("torch")
    _device_roundtrip(torch, 3840, 2160, 3, 8, np.uint8, seed=3)


def test_full_size_16384_u16_roundtrip():
    torch = pytest.importorskip("torch")
    _device_roundtrip(torch, 16384, 16384, 1, 1, np.uint16, seed=4)


@pytest.mark.parametrize("shape", [(48, 64, 1), (131, 77, 3), (270, 480, 3), (512, 512, 1)], ids=lambda s: "x".join(map(str, s)))
def test_emitted_streams_follow_the_reference_scan_order(shape):
    """fri_encode_tq_emit / fri_emit_device: quantized Some coefficients of every channel in the
    order entropy_coding.rs:283-329 pushes them, against the oracle (coefficients from
    oracle/fri_oracle.c, order from oracle/fri_order_np.py)."""
    from oracle import fri_order_np as R
    torch = pytest.importorskip("torch")
    h, w, c = shape
    frames = np.stack([uniform_image(h, w, c, seed=90 + i) for i in range(2)])
    q = smallest_layer_q(3)
    with capi.Plan(w, h, c) as plan:
        order = R.emission_order(plan.centers(), w, h)
        src = order[:, 0] * 512 + order[:, 1]
        some = plan.masks().reshape(-1)
        src = src[some[src]]
        assert len(src) == plan.emission_count()
        got = plan.encode_emit(frames, q)
        assert plan.last_launches == 2 * 2  # per frame: transform + quant kernel, emission gather
        for f in range(2):
            want, _ = oracle_encode(plan, frames[f], q)
            for ch in range(c):
                assert np.array_equal(got[f, ch], want[:, ch, :].reshape(-1)[src])
        dev = torch.device("cuda", 0)
        d_coefs = torch.from_numpy(plan.encode(frames, q)).to(dev)
        d_out = torch.empty((2, c, plan.emission_count()), dtype=torch.int32, device=dev)
        plan.emit_device(d_coefs.data_ptr(), 2, d_out.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), got)
        # int16 streams (fri_encode_tq_emit16 / fri_emit_device16)
        got16 = plan.encode_emit(frames, q, dtype=np.int16)
        assert got16.dtype == np.int16 and np.array_equal(got16, got)
        d_out16 = torch.empty((2, c, plan.emission_count()), dtype=torch.int16, device=dev)
        plan.emit_device(d_coefs.data_ptr(), 2, d_out16.data_ptr(), half=True)
        torch.cuda.synchronize()
        assert np.array_equal(d_out16.cpu().numpy(), got)


@pytest.mark.parametrize("shape", [(48, 64, 1), (131, 77, 3), (270, 480, 3)], ids=lambda s: "x".join(map(str, s)))
def test_decoder_side_of_the_emission_order(shape):
    """fri_unemit_device* / fri_decode_tq_emit*: streams in emission order -> dense blocks (None slots 0)
    -> pixels; the mirror of the emitter, against the i32 block path and the oracle."""
    torch = pytest.importorskip("torch")
    h, w, c = shape
    frames = np.stack([uniform_image(h, w, c, seed=40 + i) for i in range(3)])
    q = smallest_layer_q(5)
    with capi.Plan(w, h, c) as plan:
        some = some_of(plan)
        coefs = plan.encode(frames, q)
        streams = plan.encode_emit(frames, q)
        want_px = np.stack([oracle_decode(plan, coefs[i], some, q) for i in range(3)])
        assert np.array_equal(plan.decode_emit(streams, q), want_px)
        assert plan.last_launches == 2 * 3  # per frame: un-emit, dequant + inverse transform
        assert np.array_equal(plan.decode_emit(streams.astype(np.int16), q), want_px)
        dev = torch.device("cuda", 0)
        d_streams = torch.from_numpy(streams).to(dev)
        d_coefs = torch.full((3,) + plan.coef_shape, 12345, dtype=torch.int32, device=dev)  # None slots must be zeroed
        plan.unemit_device(d_streams.data_ptr(), 3, d_coefs.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_coefs.cpu().numpy(), coefs)  # encode leaves None slots at 0
        d16 = d_streams.to(torch.int16)
        d_coefs.fill_(-7)
        plan.unemit_device(d16.data_ptr(), 3, d_coefs.data_ptr(), half=True)
        torch.cuda.synchronize()
        assert np.array_equal(d_coefs.cpu().numpy(), coefs)
        # arbitrary stream content (not an encoder's), multiply dequantizer
        rng = np.random.Generator(np.random.PCG64(h))
        anys = rng.integers(-2000, 2001, size=streams.shape, dtype=np.int32)
        dense = np.zeros((3,) + plan.coef_shape, np.int32)
        order = plan.emission_order().astype(np.int64)
        src = order[plan.masks().reshape(-1)[order]]
        for f in range(3):
            for ch in range(c):
                blk = dense[f, :, ch, :].reshape(-1).copy()
                blk[src] = anys[f, ch]
                dense[f, :, ch, :] = blk.reshape(plan.n_tiles, 512)
        for f in range(3):
            assert np.array_equal(plan.decode_emit(anys[f], q, multiply=True)[0], oracle_decode(plan, dense[f], some, q, multiply=True))


def test_emission_gather_full_size():
    """The group-staged gather at BASELINE.json's full size against torch indexing with the plan's own
    order (the order itself is checked against the dict-based restatement on small shapes)."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    w, h, c = 4096, 4096, 3
    with capi.Plan(w, h, c) as plan:
        order = torch.from_numpy(plan.emission_order().astype(np.int64))
        some = torch.from_numpy(plan.masks().reshape(-1))
        src = order[some[order]].to(dev)
        cnt = plan.emission_count()
        assert len(src) == cnt
        gen = torch.Generator(device=dev).manual_seed(7)
        coefs = torch.randint(-255, 256, plan.coef_shape, generator=gen, device=dev, dtype=torch.int32)
        out = torch.empty((1, c, cnt), dtype=torch.int32, device=dev)
        plan.emit_device(coefs.data_ptr(), 1, out.data_ptr())
        torch.cuda.synchronize()
        for ch in range(c):
            assert torch.equal(out[0, ch], coefs[:, ch, :].reshape(-1)[src])
        back = torch.full_like(coefs, 99)
        plan.unemit_device(out.data_ptr(), 1, back.data_ptr())
        torch.cuda.synchronize()
        keep = some.reshape(plan.n_tiles, 1, 512).to(dev)
        assert torch.equal(back, torch.where(keep, coefs, torch.zeros_like(coefs)))


def test_stage_interface_mirrors_reference_pipeline():
    img = smooth_image(120, 200, 3, seed=1)
    raster = stages.RasterImage.from_array(img, stages.ColorSpace.RGB)
    opts = stages.EncoderOpts(quantization_matrix=smallest_layer_q(4))
    wi = stages.quantization.encode(stages.wavelet_transform.encode(raster, opts))  # encoder.rs:26-33
    centers, coef, some = O.from_raster(img)
    want = O.quantize(coef, some, opts.quantization_matrix)
    order = {tuple(x): i for i, x in enumerate(centers.tolist())}
    idx = np.array([order[tuple(x)] for x in wi.centers.tolist()])
    assert np.array_equal(wi.coefficients, want[idx])
    assert np.array_equal(wi.some, some[idx][:, 0, :])
    back = stages.wavelet_transform.decode(stages.quantization.decode(wi))  # decoder.rs:27-34
    ref = O.extract_values(centers, O.quantize(want, some, opts.quantization_matrix), some, 120, 200)
    assert np.array_equal(back.data, ref)
    lattice = wi.fractal_lattice()
    assert set(lattice) == set(order)


def test_bad_arguments_on_device():
    with capi.Plan(32, 32, 1) as plan:
        img = np.zeros((32, 32, 1), np.uint8)
        q = ONES.copy()
        q[4] = 0
        with pytest.raises(capi.FriError) as e:
            plan.encode(img, q)
        assert e.value.code == capi.FRI_E_INVALID
        with pytest.raises(ValueError):
            plan.encode(np.zeros((32, 32, 3), np.uint8))
    with pytest.raises(capi.FriError) as e:
        capi.Plan(32, 32, 1, device=capi.device_count())
    assert e.value.code == capi.FRI_E_CUDA
