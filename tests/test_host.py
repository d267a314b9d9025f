"""CPU suite: the C-ABI library loads and exports every declared symbol, the host-side plan
reproduces the reference's lattice (checked against the oracle), error behaviour without a
device, and the frame sharding used for N > 1 (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from frave_b200 import capi, sharding, stages
from oracle import c_oracle as O
from oracle import fri_oracle_np as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fri_cuda.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(fri_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found"
    L = ctypes.CDLL(capi.lib_path()) if os.path.exists(capi.lib_path()) else capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} is declared in include/fri_cuda.h but not exported"
    assert declared == set(capi.SYMBOL_NAMES)
    assert "sm_100a" in capi.version()


def test_header_is_valid_c_and_links(tmp_path):
    """include/fri_cuda.h compiles as plain C99 and a C program linked against the library resolves every
    prototype (tests/c_abi_check.c also exercises the host-only paths)."""
    src = os.path.join(ROOT, "tests", "c_abi_check.c")
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fri_cuda.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(fri_[a-z0-9_]+)\s*\(", header))
    used = set(re.findall(r"\(fn\)(fri_[a-z0-9_]+)", open(src).read()))
    assert used == declared, declared ^ used
    capi.lib()
    libdir, libname = os.path.split(capi.lib_path())
    exe = str(tmp_path / "c_abi_check")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", libdir, "-l:" + libname, "-Wl,-rpath," + libdir]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "c abi ok" in run.stdout


def test_kernels_are_compiled_for_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", capi.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


@pytest.mark.parametrize("shape", [(1, 1, 1), (10, 10, 3), (48, 64, 1), (37, 100, 3), (300, 7, 3), (512, 512, 1),
                                   (240, 427, 3)])
def test_plan_matches_reference_lattice(shape):
    h, w, c = shape
    with capi.Plan(w, h, c, device=-1) as p:
        img = np.zeros((h, w, c), np.uint8)
        centers, _, some = O.from_raster(img)
        assert p.n_built == len(O.fractal_divide(w, h))
        assert p.n_tiles == len(centers)
        pc = p.centers()
        order = {tuple(x): i for i, x in enumerate(centers.tolist())}
        assert len(order) == len(pc)
        idx = np.array([order[tuple(x)] for x in pc.tolist()], dtype=np.int64)
        assert np.array_equal(p.masks(), some[idx][:, 0, :])
        off = N.leaf_offsets(9)
        x, y = pc[:, 0:1] + off[None, :, 0], pc[:, 1:2] + off[None, :, 1]
        inside = (x >= 0) & (y >= 0) & (x < w) & (y < h)
        assert p.pixels_covered == int(inside.sum())
        assert p.n_full == int(inside.all(axis=1).sum())
        assert p.coefs_per_frame == p.n_tiles * c * 512
        info = p.launch_info()
        assert info["smem_bytes"] <= 227 * 1024 and info["n_base_tiles"] == p.n_tiles


def test_survey_tile_counts():
    # built / retained / fully-inside
    for (w, h), want in {(512, 512): (617, 578, 448), (1920, 1080): (4317, 4221, 3881)}.items():
        with capi.Plan(w, h, 3, device=-1) as p:
            assert (p.n_built, p.n_tiles, p.n_full) == want
            assert p.pixels_covered == w * h


def test_deep_plan_is_a_union_of_base_tiles():
    # depth-D fractal = 2^(D-9) base tiles (wavelet_transform.rs:47-53); check against the oracle's BFS
    for depth in (10, 12):
        w, h = 300, 200
        with capi.Plan(w, h, 1, depth=depth, device=-1) as p:
            centers, _, some = O.from_raster(np.zeros((h, w, 1), np.uint8), depth=depth)
            assert p.n_built == len(O.fractal_divide(w, h, depth))
            assert p.n_tiles == len(centers)
            pc = p.centers()
            order = {tuple(x): i for i, x in enumerate(centers.tolist())}
            idx = np.array([order[tuple(x)] for x in pc.tolist()], dtype=np.int64)
            assert np.array_equal(p.masks(), some[idx][:, 0, :])
            # the kernels process the base tiles that hold at least one in-image pixel; the rest (all `None`) are
            # zero-filled on encode and skipped on decode
            off9 = N.leaf_offsets(9)
            sub = N.leaf_offsets(depth)[::512]  # centres of the 2^(depth-9) base tiles relative to the fractal centre
            bc = (pc[:, None, :] + sub[None, :, :]).reshape(-1, 2)
            x, y = bc[:, 0:1] + off9[None, :, 0], bc[:, 1:2] + off9[None, :, 1]
            present = ((x >= 0) & (y >= 0) & (x < w) & (y < h)).any(axis=1)
            assert p.launch_info()["n_base_tiles"] == int(present.sum()) <= p.n_tiles << (depth - 9)


@pytest.mark.parametrize("shape", [(10, 10), (64, 48), (100, 37), (131, 77), (480, 270), (300, 400)])
def test_emission_order_matches_the_dict_based_restatement(shape):
    """the plan's arithmetic scan against oracle/fri_order_np.py (a literal
    restatement of sort_lattice / scan_level with hash maps) — same order, same Some count."""
    from oracle import fri_order_np as R
    w, h = shape
    with capi.Plan(w, h, 3, device=-1) as p:
        centers = p.centers()
        want = R.emission_order(centers, w, h)
        got = p.emission_order()
        assert np.array_equal(got, want[:, 0] * 512 + want[:, 1])
        assert sorted(got.tolist()) == list(range(p.n_tiles * 512))  # every (tile, coefficient) exactly once
        assert p.emission_count() == int(p.masks().sum())
        # entropy_coding.rs:283-329: first every DC, then every root residue, both in level-0 order
        n = p.n_tiles
        assert (got[:n] % 512 == 0).all() and (got[n:2 * n] % 512 == 1).all()
        assert np.array_equal(got[:n] // 512, got[n:2 * n] // 512)
        for level in range(1, 9):
            blk = got[n << level:n << (level + 1)] % 512
            assert blk.min() == 1 << level and blk.max() == (2 << level) - 1


def test_emission_order_mirrors_the_reference_assertion():
    """For some sizes the reference's scan misses a node and its assert_eq! at
    wavelet_transform.rs:701 would panic; both restatements must agree on that too."""
    from oracle import fri_order_np as R
    w, h = 257, 300
    with capi.Plan(w, h, 3, device=-1) as p:
        with pytest.raises(AssertionError):
            R.emission_order(p.centers(), w, h)
        with pytest.raises(capi.FriError) as e:
            p.emission_order()
        assert e.value.code == capi.FRI_E_UNSUPPORTED and "701" in str(e.value)
    with capi.Plan(300, 200, 1, depth=10, device=-1) as p:
        with pytest.raises(capi.FriError) as e:
            p.emission_order()
        assert e.value.code == capi.FRI_E_UNSUPPORTED


def test_invalid_arguments_and_no_cpu_fallback():
    for args in [(0, 5, 1), (5, 0, 1), (5, 5, 2), (5, 5, 4)]:
        with pytest.raises(capi.FriError) as e:
            capi.Plan(*args, device=-1)
        assert e.value.code == capi.FRI_E_INVALID
    with pytest.raises(capi.FriError):
        capi.Plan(8, 8, 1, depth=8, device=-1)
    with pytest.raises(capi.FriError):
        capi.Plan(8, 8, 1, sample_bytes=3, device=-1)
    with capi.Plan(16, 16, 1, device=-1) as p:
        with pytest.raises(capi.FriError) as e:
            p.encode(np.zeros((16, 16, 1), np.uint8))
        assert e.value.code == capi.FRI_E_CUDA and "no CPU fallback" in str(e.value)
        with pytest.raises(capi.FriError):
            p.decode(np.zeros((1,) + p.coef_shape, np.int32))
        with pytest.raises(capi.FriError) as e:
            p.encode(np.zeros((16, 16, 1), np.uint8), dtype=np.int16)  # the 16-bit transport has no fallback either
        assert e.value.code == capi.FRI_E_CUDA
        p.set_bands(1)
        p.set_bands(0)
        with pytest.raises(capi.FriError) as e:
            p.set_bands(9)
        assert e.value.code == capi.FRI_E_INVALID
    if capi.device_count() == 0:
        with pytest.raises(capi.FriError) as e:
            capi.Plan(16, 16, 1, device=0)
        assert e.value.code == capi.FRI_E_CUDA
        with pytest.raises(stages.StageError):
            stages.wavelet_transform.encode(stages.RasterImage.from_array(np.zeros((16, 16, 3), np.uint8)))


def test_kernel_division_routine_is_exact_truncating_division():
    """quantization.rs:19/:37 is Rust `i32 / i32` (truncation toward zero); the kernels use a
    multiply-high routine instead of a hardware divide — check it against C semantics."""
    L = capi.lib()
    rng = np.random.Generator(np.random.PCG64(0))
    edge = [0, 1, -1, 2, -2, 254, 255, 256, -255, -256, 65535, -65535, 65536, 2**31 - 1, -(2**31), -(2**31) + 1]
    qs = list(range(1, 70)) + [127, 128, 255, 256, 257, 641, 1000, 4095, 4096, 32768, 65535, 65536, 65537,
                                (1 << 20) + 7, 2**31 - 1, 2**30, 715827883]
    for q in qs:
        vals = edge + [int(v) for v in rng.integers(-2**31, 2**31 - 1, 300)] + list(range(-600, 600, 7)) + \
            [k * q + o for k in (-3, -1, 1, 2, 1000) for o in (-1, 0, 1) if -2**31 <= k * q + o < 2**31]
        for v in vals:
            want = abs(v) // q * (1 if v >= 0 else -1)
            assert L.fri_quant_divide(v, q) == want, (v, q)
            assert L.fri_quant_divide_magic(v, q) == want, (v, q)


def test_encoder_narrow_range_division_is_exact_on_its_whole_domain():
    """The encoder divides residues of 8/16-bit samples (|n| <= 65535) with a cheaper magic: check
    every numerator of that range against C semantics for a spread of divisors."""
    L = capi.lib()
    fn = np.vectorize(lambda v, q: L.fri_quant_divide_small(int(v), int(q)), otypes=[np.int64])
    n = np.concatenate([np.arange(-65535, 65536, 1)])
    for q in [2, 3, 4, 5, 7, 8, 64, 255, 256, 1000, 65535, 65536, 65537, 131072, 2**31 - 1]:
        sub = n if q in (3, 4, 7) else n[::17]
        want = np.sign(sub) * (np.abs(sub) // q)
        assert np.array_equal(fn(sub, q), want), q
    rng = np.random.Generator(np.random.PCG64(3))
    for q in rng.integers(2, 70000, 40).tolist():
        sub = rng.integers(-65535, 65536, 4000)
        assert np.array_equal(fn(sub, q), np.sign(sub) * (np.abs(sub) // q)), q


def test_shard_frames_partitions_the_batch():
    for n in (0, 1, 7, 16, 255, 256):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                sh = sharding.shard_frames(n, r, world)
                seen += list(sh)
                for f in sh:
                    assert sharding.frame_owner(f, n, world) == r
            assert seen == list(range(n))
            sizes = [len(sharding.shard_frames(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_image_parts_partition_the_plan():
    """fri_plan_part: tile ranges of the parts are disjoint and cover the plan, group ranges likewise, every pixel a
    part's tiles own lies inside the part's band of rows; depth > 9 and bad arguments are refused."""
    for (w, h, c), worlds in (((320, 240, 3), (1, 2, 3, 5)), ((1920, 1080, 1), (2, 8)), ((77, 131, 3), (1, 2))):
        img = np.full((h, w, c), 255, np.uint8)
        with capi.Plan(w, h, c, device=-1) as plan:
            centers = plan.centers()
            n_groups = plan.launch_info()["n_groups"]
            for world in worlds:
                if world > n_groups:
                    continue
                parts = [sharding.shard_image(plan, r, world) for r in range(world)]
                assert parts[0]["tile_begin"] == 0 and parts[-1]["tile_end"] == plan.n_tiles
                assert parts[0]["group_begin"] == 0 and parts[-1]["group_end"] == n_groups
                for a, b in zip(parts, parts[1:]):
                    assert a["tile_end"] == b["tile_begin"] and a["group_end"] == b["group_begin"]
                for r, p in enumerate(parts):
                    assert 0 <= p["row_begin"] <= p["row_end"] <= h
                    coef, some = O.extract_tiles(img, centers[p["tile_begin"]:p["tile_end"]])
                    own = O.extract_values(centers[p["tile_begin"]:p["tile_end"]], coef, some, h, w)
                    assert not own[:p["row_begin"]].any() and not own[p["row_end"]:].any()
                    for peer, lo, hi in sharding.overlaps(plan, r, world):
                        assert peer != r and p["row_begin"] <= lo < hi <= p["row_end"]
                    margin = plan.launch_info()["region_h"]
                    for push in sharding.halo_pushes(plan, r, world):  # the groups along the cut, for the peer-memory exchange
                        assert p["group_begin"] <= push["first"] < push["last"] <= p["group_end"]
                        assert push["span_begin"] < push["row_hi"] and push["span_end"] > push["row_lo"]
                        if abs(push["peer"] - r) == 1 and world <= 3:  # neighbours: the span fits the peer's band plus one region height
                            assert parts[push["peer"]]["row_begin"] - margin <= push["span_begin"]
                            assert push["span_end"] <= parts[push["peer"]]["row_end"] + margin
            with pytest.raises(capi.FriError):
                plan.part(2, 2)
            with pytest.raises(capi.FriError):
                plan.part(0, 0)
    with capi.Plan(300, 200, 1, depth=12, device=-1) as deep:
        with pytest.raises(capi.FriError) as ei:
            deep.part(0, 2)
        assert ei.value.code == capi.FRI_E_UNSUPPORTED
        with pytest.raises(capi.FriError) as ei:
            deep.groups_in_rows(0, 1, 0, 10)
        assert ei.value.code == capi.FRI_E_UNSUPPORTED
    with capi.Plan(320, 240, 3, device=-1) as plan:
        n_groups = plan.launch_info()["n_groups"]
        whole = plan.groups_in_rows(0, n_groups, 0, 240)
        assert (whole["first"], whole["last"], whole["span_begin"], whole["span_end"]) == (0, n_groups, 0, 240)
        none = plan.groups_in_rows(0, n_groups, 240, 240)
        assert none["first"] == none["last"]
        for bad in ((5, 2), (0, n_groups + 1)):
            with pytest.raises(capi.FriError) as ei:
                plan.groups_in_rows(bad[0], bad[1], 0, 10)
            assert ei.value.code == capi.FRI_E_INVALID
            with pytest.raises(capi.FriError) as ei:
                plan.decode_device_groups(16, 0, 16, 0, bad[0], bad[1])
            assert ei.value.code == capi.FRI_E_INVALID
        with pytest.raises(capi.FriError) as ei:  # a valid range still needs a device: no CPU fallback
            plan.decode_device_groups(16, 0, 16, 0, 0, 1)
        assert ei.value.code == capi.FRI_E_CUDA


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from frave_b200 import capi, sharding
from oracle import c_oracle as O
import bench
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
# the product's host side of the N > 1 path: every rank builds the same plan (replicated, no exchange), takes its
# contiguous block of the batch, and the per-rank frame counts bench.py reports add up to the batch
with capi.Plan(3840, 2160, 3, device=-1) as plan:
    meta = torch.tensor([plan.n_tiles, plan.coefs_per_frame, plan.n_built], dtype=torch.int64)
gathered = [torch.zeros_like(meta) for _ in range(world)]
dist.all_gather(gathered, meta)
assert all(torch.equal(g, meta) for g in gathered) and meta[0].item() == 16541  # SURVEY.md §8(a)
mine256 = sharding.shard_frames(256, rank, world)
cnt = torch.tensor([len(mine256)], dtype=torch.int64)
dist.all_reduce(cnt)
assert cnt.item() == 256 and len(mine256) == 128
bench.W, bench.H, bench.C = 3840, 2160, 3
cfg = bench.workload_config(world, "batch256")
assert cfg["frames_per_gpu"] == [128, 128] and cfg["global_frames"] == 256
try:  # and there is no CPU path to fall back to on a rank without a device
    capi.Plan(64, 64, 1, device=-1).encode(np.zeros((64, 64, 1), np.uint8))
    raise SystemExit("compute call succeeded without a device")
except capi.FriError as e:
    assert e.code == capi.FRI_E_CUDA
n_frames, h, w, c = 5, 40, 56, 3
frames = [np.random.Generator(np.random.PCG64(100 + f)).integers(0, 256, (h, w, c), dtype=np.uint8) for f in range(n_frames)]
mine = sharding.shard_frames(n_frames, rank, world)
# each rank transforms only its own frames (no data-path collective); a checksum of checksums is
# all-reduced for the test only, like the timing reduction in bench.py
local = 0
for f in mine:
    _, coef, _ = O.from_raster(frames[f])
    local += int(coef.astype(np.int64).sum()) * (f + 1)
t = torch.tensor([local, len(mine)], dtype=torch.int64)
dist.all_reduce(t)
want = sum(int(O.from_raster(frames[f])[1].astype(np.int64).sum()) * (f + 1) for f in range(n_frames))
assert t[0].item() == want and t[1].item() == n_frames, (t, want)
# ---- one image split by tile-group ranges (SURVEY.md §8(e)): every rank transforms its own tiles, decodes only the
# pixels they own, and the overlap rows are exchanged and merged by addition — the host logic of
# bench.py --workload image16k, with the oracle standing in for the kernels
ih, iw, ic = 240, 320, 3
img = np.random.Generator(np.random.PCG64(7)).integers(1, 256, (ih, iw, ic), dtype=np.uint8)  # no zero pixel: ownership is visible
with capi.Plan(iw, ih, ic, device=-1) as plan:
    part = sharding.shard_image(plan, rank, world)
    shared = sharding.overlaps(plan, rank, world)
    centers = plan.centers()
    covered = plan.pixels_covered == iw * ih
t0, t1, r0, r1 = part["tile_begin"], part["tile_end"], part["row_begin"], part["row_end"]
coef, some = O.extract_tiles(img, centers[t0:t1])
mine_px = O.extract_values(centers[t0:t1], coef, some, ih, iw)      # only the pixels my tiles own, zeros elsewhere
assert not mine_px[:r0].any() and not mine_px[r1:].any()            # ... all inside my band of rows
band = torch.from_numpy(mine_px[r0:r1].astype(np.int32))
assert shared == [(1 - rank, max(r0, shared[0][1]), min(r1, shared[0][2]))] and len(shared) == 1
peer, lo, hi = shared[0]
theirs = torch.zeros((hi - lo, iw, ic), dtype=torch.int32)
ops = [dist.P2POp(dist.isend, band[lo - r0:hi - r0].contiguous(), peer), dist.P2POp(dist.irecv, theirs, peer)]
for w_ in dist.batch_isend_irecv(ops):
    w_.wait()
assert not ((band[lo - r0:hi - r0] != 0) & (theirs != 0)).any()     # a pixel has exactly one owner
band[lo - r0:hi - r0] += theirs
if covered:
    assert np.array_equal(band.numpy().astype(np.uint8), img[r0:r1])
tiles = torch.tensor([t1 - t0], dtype=torch.int64)
dist.all_reduce(tiles)
assert tiles.item() == len(centers)
el = torch.tensor([1.0 + rank])
dist.all_reduce(el, op=dist.ReduceOp.MAX)
assert el.item() == 2.0
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_sharding_over_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    for r, p in enumerate(procs):
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert f"rank {r} ok" in out


def _c_prototypes(text):
    """name -> argument count of every fri_* prototype in a C header (comments stripped)."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(fri_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_rust_extern_block_matches_header():
    """rust/libfri-cuda/src/lib.rs cannot be compiled here (no cargo): its extern "C" block is diffed against
    include/fri_cuda.h instead — every prototype bound, none invented, same argument counts."""
    want = _c_prototypes(open(os.path.join(ROOT, "include", "fri_cuda.h")).read())
    src = open(os.path.join(ROOT, "rust", "libfri-cuda", "src", "lib.rs")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', src, flags=re.S).group(1)
    got = {}
    for m in re.finditer(r"\bfn (fri_[a-z0-9_]+)\s*\(([^)]*)\)", block):
        args = m.group(2).strip()
        got[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    assert set(got) == set(want), set(got) ^ set(want)
    assert got == want, {k: (got[k], want[k]) for k in want if got[k] != want[k]}
    # the safe wrappers check slice lengths before crossing the boundary and keep async mode unsafe
    assert "want_len(" in src and "pub unsafe fn set_async" in src and "struct PinnedBuf" in src
