#!/usr/bin/env python
"""bench.py — throughput of the fractal transform + quantization hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload frame|batch256]

One step = one pass of the hot path over one batch of synthetic input: the fused forward transform +
quantization (encode) followed by the fused dequantization + inverse transform (decode).

Workloads
  frame     (default) BASELINE.json configs[1]: one synthetic 4096x4096 8-bit RGB image per GPU and step
            (frames are independent, so N GPUs = N frames, weak scaling, no data-path collective);
  batch256  BASELINE.json configs[2]: a batch of 256 synthetic 3840x2160 RGB frames, sharded over the ranks with
            frave_b200.sharding.shard_frames (256 / N frames per GPU and step, strong scaling).

Prints ONE JSON line (rank 0).  `value` is device-resident MPix/s over exactly --steps steps, `e2e` the same
metric through the host-buffer C-ABI entry points (pinned host memory, copies inside the timed region),
`roofline` the dominant kernel against the measured HBM copy bandwidth, `cpu_baseline` the CPU oracle (single
thread, like the reference) on the same image.

Per-kernel durations do not depend on --steps: each kernel is launched `KERNEL_REPS` (>= 200) times back to
back on the launching stream between two CUDA events (the average launch duration a stream of frames sees,
programmatic dependent launch included), and the isolated, event-bracketed single-launch median is reported
beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "MPix/s fractal transform+quant (enc/dec)"
UNIT = "MPix/s"
W, H, C = 4096, 4096, 3
FRAMES = 1                  # frames per GPU per step (one batched launch per direction)
BATCH_FRAMES = 256          # --workload batch256
PREHEAT_S = 1.0
KERNEL_REPS = 200           # back-to-back launches per kernel for the roofline numbers, whatever --steps is
SMALLEST_LAYER_DIVISOR = 4  # q[8] = q[9] = 4: "dividing the smallest layer of fractals" (README.md:12)
BYTES_PER_SAMPLE = 5        # u8 pixel + i32 coefficient, either direction
N_SETS = int(os.environ.get("FRI_BENCH_SETS", "4"))  # rotating buffer sets: 4 x (50 + 201 + 50 MB) = 1.2 GB >> 126 MB of L2


def quant_matrix() -> np.ndarray:
    q = np.ones(32, np.int32)
    q[8] = q[9] = SMALLEST_LAYER_DIVISOR
    return q


def workload_config(n_gpus: int, workload: str) -> dict:
    if workload == "batch256":
        per = [len(_shard(BATCH_FRAMES, r, n_gpus)) for r in range(n_gpus)]
        return {
            "workload": f"batch of {BATCH_FRAMES} synthetic {W}x{H}x{C} u8 frames (BASELINE.json configs[2]) sharded over "
                        f"{n_gpus} GPU(s) with sharding.shard_frames; step = fused transform+quant encode then fused "
                        f"dequant+inverse decode of the whole batch, one batched launch per direction and GPU",
            "frames_per_gpu": per, "global_frames": BATCH_FRAMES, "depth": 9,
            "quant": f"q[8]=q[9]={SMALLEST_LAYER_DIVISOR}, other layers 1; decode divides again like quantization.rs:37",
            "l2": f"{max(per) * W * H * C * 5 / 1e9:.1f} GB touched per launch and GPU, far beyond the 126 MB L2",
            "parallelism": f"frames sharded over {n_gpus} GPU(s), no collective",
        }
    return {
        "workload": f"{W}x{H}x{C} u8 synthetic image{' (BASELINE.json configs[1])' if (W, H, C, FRAMES) == (4096, 4096, 3, 1) else ''}"
                    f"; step = fused transform+quant encode then fused dequant+inverse decode of {FRAMES} frame(s) per GPU",
        "frames_per_gpu": FRAMES,
        "global_frames": n_gpus * FRAMES,
        "depth": 9,
        "quant": f"q[8]=q[9]={SMALLEST_LAYER_DIVISOR}, other layers 1; decode divides again like quantization.rs:37",
        "l2": f"inputs rotate over {N_SETS} buffer sets ({N_SETS * FRAMES * W * H * C * 6 / 1e9:.1f} GB) so every timed "
              f"launch reads cold HBM",
        "parallelism": f"frames sharded over {n_gpus} GPU(s), no collective",
    }


def _shard(n_frames: int, rank: int, world: int) -> range:
    from frave_b200 import sharding
    return sharding.shard_frames(n_frames, rank, world)


def measured_peak_gbs() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.tmp.close()
        os.unlink(self.tmp.name)
        # samples taken while the GPU was busy: the load phase (pre-heat + warm-up + timed steps)
        busy = sorted(sm)[len(sm) // 4:] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def synthetic_image(seed: int) -> np.ndarray:
    return np.random.Generator(np.random.PCG64(seed)).integers(0, 256, size=(H, W, C), dtype=np.uint8)


def cpu_pass(O, img, centers, some, q, nthreads, tiles=None, coef_buf=None, out_buf=None) -> float:
    """One oracle pass (transform+quant, then dividing dequant+inverse) over `tiles`; seconds."""
    cen = centers if tiles is None else centers[tiles]
    sm = None if some is None else (some if tiles is None else some[tiles])
    t0 = time.perf_counter()
    coef = O.encode_tiles(img, cen, q, nthreads=nthreads, out=coef_buf)
    O.decode_tiles(cen, coef, sm, q, H, W, nthreads=nthreads, out=out_buf)
    return time.perf_counter() - t0


def oracle_lattice(O) -> tuple[np.ndarray, float]:
    """The reference's tile list from the ORACLE alone (no libfri_cuda involved): fractal_divide's BFS
    (wavelet_transform.rs:450-484) and the retain rule (:415-416, a tile survives iff one of its 512 leaves is
    inside the image).  Returns (retained centres, seconds the BFS took)."""
    from oracle import fri_oracle_np as N
    t0 = time.perf_counter()
    built = O.fractal_divide(W, H)
    t_bfs = time.perf_counter() - t0
    off = N.leaf_offsets(9)
    keep = np.zeros(len(built), bool)
    for lo in range(0, len(built), 8192):
        c = built[lo:lo + 8192]
        x, y = c[:, 0:1] + off[None, :, 0], c[:, 1:2] + off[None, :, 1]
        keep[lo:lo + 8192] = ((x >= 0) & (y >= 0) & (x < W) & (y < H)).any(axis=1)
    return np.ascontiguousarray(built[keep]), t_bfs


def run_reference(args, rank: int) -> None:
    """--impl reference: the reference's CPU path.  The reference is pure Rust and cannot be built here (no
    cargo/rustc), so this times the C oracle restating it (oracle/ only — nothing of frave_b200 is loaded), on
    all host threads."""
    if rank != 0:
        return
    from oracle import c_oracle as O

    threads = host_threads()
    q = quant_matrix()
    img = synthetic_image(2)
    centers, t_bfs = oracle_lattice(O)
    n_tiles = len(centers)
    # calibrate, then bound the per-step sample so that the whole run stays within ~90 s
    probe = np.arange(min(2048, n_tiles))
    t_probe = cpu_pass(O, img, centers, None, q, threads, probe)
    per_tile = t_probe / len(probe)
    budget = 90.0 / max(1, args.steps + args.warmup)
    sample = int(max(256, min(n_tiles, budget / per_tile)))
    tiles = np.linspace(0, n_tiles - 1, sample).astype(np.int64)
    coef_buf = np.empty((sample, C, 512), np.int32)
    out_buf = np.zeros((H, W, C), np.uint8)
    for _ in range(args.warmup):
        cpu_pass(O, img, centers, None, q, threads, tiles, coef_buf, out_buf)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_pass(O, img, centers, None, q, threads, tiles, coef_buf, out_buf)
    px_per_step = W * H * sample / n_tiles  # the sampled share of the image's pixels
    mpix = px_per_step * args.steps / t / 1e6
    desc = f"{sample} of {n_tiles} tiles ({px_per_step / 1e6:.2f} MPix) of the {W}x{H}x{C} image per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32", "data": "synthetic", "config": workload_config(args.gpus, "frame"),
        "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": mpix, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "threads": threads,
        "lattice_build_ms": 1e3 * t_bfs,
        "lattice_note": "fractal_divide BFS of the C oracle for this image size, NOT inside the timed region; the real "
                        "reference rebuilds the lattice (plus 511 HashMap inserts per tile) on every encode and every "
                        "decode (wavelet_transform.rs:405-410, :392-403), so the timed arithmetic-only figure flatters it",
        "note": "reference is Rust and unbuildable here (no cargo); timed the C oracle port of its hot path, "
                "tiles split over all host threads; tile list from the oracle's own BFS + retain (no libfri_cuda)",
    }
    print(json.dumps(line), flush=True)


def b2b(fn, reps: int, torch) -> float:
    """Average duration (ms) of `reps` back-to-back calls of fn(i) between two CUDA events on the current stream."""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def isolated(fn, reps: int, torch) -> float:
    """Median duration (ms) of single launches, each bracketed by its own pair of CUDA events."""
    evs = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs)


def run_b200(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    from frave_b200 import capi

    if capi.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: libfri_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    q = quant_matrix()
    batch = args.workload == "batch256"
    frames = len(_shard(BATCH_FRAMES, rank, world)) if batch else FRAMES
    n_sets = 1 if batch else N_SETS  # a 256-frame shard is tens of GB per launch: no rotation needed
    t0 = time.perf_counter()
    plan = capi.Plan(W, H, C, device=local_rank)
    plan_build_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    emission_count = plan.emission_count()  # builds the emission order (the *_emit* entry points need it)
    emission_build_ms = 1e3 * (time.perf_counter() - t0)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- buffers: rotating sets, synthetic pixels generated per rank
    img0 = synthetic_image(2 + rank)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    px = []
    for _ in range(n_sets):
        t = torch.empty((frames, H, W, C), dtype=torch.uint8, device=dev)
        for f0 in range(0, frames, 8):  # generated in slices: randint works in int64
            n = min(8, frames - f0)
            t[f0:f0 + n] = torch.randint(0, 256, (n, H, W, C), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)
        px.append(t)
    px[0][0] = torch.from_numpy(img0).to(dev)
    coefs = [torch.empty((frames,) + plan.coef_shape, dtype=torch.int32, device=dev) for _ in range(n_sets)]
    outs = [torch.empty((frames, H, W, C), dtype=torch.uint8, device=dev) for _ in range(n_sets)]

    # sanity (untimed): encode -> decode is the identity at q == 1
    plan.encode_device(px[0].data_ptr(), frames, coefs[0].data_ptr(), None, stream)
    plan.decode_device(coefs[0].data_ptr(), frames, outs[0].data_ptr(), None, False, stream)
    torch.cuda.synchronize()
    if not torch.equal(px[0], outs[0]) and not os.environ.get("FRI_BENCH_NO_SANITY"):  # (timing-only hack builds)
        raise RuntimeError("sanity check failed: encode -> decode is not lossless at q == 1")
    for s in range(n_sets):
        plan.encode_device(px[s].data_ptr(), frames, coefs[s].data_ptr(), q, stream)
    torch.cuda.synchronize()

    launches = 0

    def enc(i: int) -> None:
        nonlocal launches
        a = i % n_sets
        plan.encode_device(px[a].data_ptr(), frames, coefs[a].data_ptr(), q, stream)
        launches += plan.last_launches

    def dec(i: int) -> None:
        nonlocal launches
        b = (i + n_sets // 2) % n_sets
        plan.decode_device(coefs[b].data_ptr(), frames, outs[b].data_ptr(), q, False, stream)
        launches += plan.last_launches

    def step(i: int) -> None:
        enc(i)
        dec(i)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_end = time.perf_counter() + PREHEAT_S  # pre-heat so clocks are sampled under the same load
    i = 0
    while time.perf_counter() < t_end:
        for _ in range(max(1, 50 // frames)):
            step(i)
            i += 1
        torch.cuda.synchronize()
    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    # ---- the timed region: exactly --steps steps, device events on the launching stream, nothing in between
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    start.record()
    for k in range(args.steps):
        step(k)
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed_ms = start.elapsed_time(end)
    timed_launches = launches

    # ---- per-kernel durations, independent of --steps: KERNEL_REPS back-to-back launches of one kernel between
    # two events (what a stream of frames sees), and the isolated single-launch median
    reps = max(8, KERNEL_REPS // frames) if not batch else 4
    enc_ms, dec_ms = b2b(enc, reps, torch), b2b(dec, reps, torch)
    enc_iso, dec_iso = isolated(enc, min(reps, 200), torch), isolated(dec, min(reps, 200), torch)
    # the same launches with the independent-calls hint (fri_plan_set_independent_calls): consecutive launches touch
    # different buffer sets here, so their kernels may overlap — what a stream of independent frames can opt into
    o_enc = o_dec = o_pair = 0.0
    if not batch and n_sets >= 4:
        plan.set_independent_calls(True)
        o_enc, o_dec = b2b(enc, reps, torch), b2b(dec, reps, torch)
        o_pair = b2b(step, reps, torch)
        plan.set_independent_calls(False)
    clocks = sampler.stop() if sampler else None

    # ---- the same step on int16 coefficient arrays (fri_*_tq_device16): 3 B per sample instead of 5,
    # reported as its own variant with its own byte count; the headline stays int32
    v_enc = v_dec = 0.0
    if not batch:
        c16 = [torch.empty((frames,) + plan.coef_shape, dtype=torch.int16, device=dev) for _ in range(n_sets)]
        for s_ in range(n_sets):
            plan.encode_device(px[s_].data_ptr(), frames, c16[s_].data_ptr(), q, stream, half=True)
        v_enc = b2b(lambda i: plan.encode_device(px[i % n_sets].data_ptr(), frames, c16[i % n_sets].data_ptr(), q, stream, half=True),
                    reps, torch)
        v_dec = b2b(lambda i: plan.decode_device(c16[i % n_sets].data_ptr(), frames, outs[i % n_sets].data_ptr(), q, False, stream,
                                                 half=True), reps, torch)
        del c16

    # ---- steady state: BASELINE.json configs[2] per-GPU share at 8 GPUs (32 batched 4K frames in one
    # launch per direction).  Reported beside the headline because a single 4096^2 frame is a ~50 us
    # launch whose ramp-up and last partial wave cost 15-25 %.
    batched = None
    if not args.no_batched and not batch:
        bw, bh, bf = 3840, 2160, 32
        bplan = capi.Plan(bw, bh, C, device=local_rank)
        bpx = torch.randint(0, 256, (bf, bh, bw, C), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)
        bco = torch.empty((bf,) + bplan.coef_shape, dtype=torch.int32, device=dev)
        bout = torch.empty_like(bpx)
        b_enc = b2b(lambda i: bplan.encode_device(bpx.data_ptr(), bf, bco.data_ptr(), q, stream), 8, torch)
        b_dec = b2b(lambda i: bplan.decode_device(bco.data_ptr(), bf, bout.data_ptr(), q, False, stream), 8, torch)
        batched = (bw, bh, bf, b_enc, b_dec)
        del bpx, bco, bout
        bplan.close()

    # ---- end to end through the host-buffer C ABI: pinned host memory, H2D and D2H in the timed region.
    # Every e2e step moves one frame through encode and one through decode, each with its own host->device and
    # device->host copies.  Driving patterns:
    #   serial  = one host thread calls encode, then decode (the reference's single-threaded shape);
    #   duplex  = an encoder thread and a decoder thread, one plan handle each (the ABI is re-entrant across
    #             handles), so one call's device->host copy overlaps the other's host->device copy (PCIe is
    #             full duplex);
    #   async   = one thread, both handles in asynchronous mode, both synced every step;
    # and host-side coefficient formats: i32 blocks (fri_*_tq), i16 blocks (fri_*_tq16), p10 = emission-ordered
    # streams in the 10-bit packed transport (fri_*_tq_emit10: what the host entropy coder consumes, 1.25 B per
    # coefficient).
    e2e_steps = 24  # per variant and repetition, whatever --steps is (a step is ~3 ms: short runs are noisy)
    e2e, e2e_bytes = {}, {}
    if not args.no_e2e:
        px_h = capi.PinnedBuffer((1, H, W, C), np.uint8)
        out_h = capi.PinnedBuffer((1, H, W, C), np.uint8)
        px_h.array[0] = img0
        dplan = capi.Plan(W, H, C, device=local_rank)  # the decoder thread's handle
        dplan.emission_count()

        def duplex(enc_call, dec_call) -> float:
            return statistics.median(duplex_once(enc_call, dec_call) for _ in range(3))

        def duplex_once(enc_call, dec_call) -> float:
            errors = []

            def loop(call):
                try:
                    for _ in range(e2e_steps):
                        call()
                except Exception as exc:  # surfaced after the join
                    errors.append(exc)

            if world > 1:
                dist.barrier()
            th = [threading.Thread(target=loop, args=(enc_call,)), threading.Thread(target=loop, args=(dec_call,))]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
            if errors:
                raise errors[0]
            return dt

        for cdt, tag in ((np.int32, "i32"), (np.int16, "i16")):
            cf_enc = capi.PinnedBuffer((1,) + plan.coef_shape, cdt)
            cf_dec = capi.PinnedBuffer((1,) + plan.coef_shape, cdt)
            plan.encode(px_h.array, q, out=cf_enc.array)  # warm-up (allocates the plans' device slots)
            cf_dec.array[...] = cf_enc.array
            dplan.decode(cf_dec.array, q, out=out_h.array)
            plan.decode(cf_dec.array, q, out=out_h.array)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                plan.encode(px_h.array, q, out=cf_enc.array)
                plan.decode(cf_enc.array, q, out=out_h.array)
            e2e["serial_" + tag] = time.perf_counter() - t0
            plan.set_bands(1)   # concurrent callers: the other thread's call supplies the overlap (fri_plan_set_bands)
            dplan.set_bands(1)
            e2e["duplex_" + tag] = duplex(lambda: plan.encode(px_h.array, q, out=cf_enc.array),
                                          lambda: dplan.decode(cf_dec.array, q, out=out_h.array))
            # the same overlap from ONE host thread: asynchronous mode, both handles enqueued, then both synced
            plan.set_async(True)
            dplan.set_async(True)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                plan.encode(px_h.array, q, out=cf_enc.array)
                dplan.decode(cf_dec.array, q, out=out_h.array)
                plan.sync()
                dplan.sync()
            e2e["async_" + tag] = time.perf_counter() - t0
            plan.set_async(False)
            dplan.set_async(False)
            plan.set_bands(0)
            dplan.set_bands(0)
            e2e_bytes[tag] = (px_h.array.nbytes + cf_dec.array.nbytes, cf_enc.array.nbytes + out_h.array.nbytes)
            cf_enc.free(); cf_dec.free()
        # 9-bit packed emission streams (everything an 8-bit image produces fits: |k| <= 255), duplex and quad
        nb9 = plan.emission_packed_size(9)
        p9 = [capi.PinnedBuffer((1, C, nb9), np.uint8) for _ in range(2)]
        plan.encode_emit_packed(px_h.array, q, 9, out=p9[0].array)
        p9[1].array[...] = p9[0].array
        dplan.decode_emit_packed(p9[1].array, q, 9, out=out_h.array)
        if not np.array_equal(out_h.array, plan.decode(plan.encode(px_h.array, q), q)):
            raise RuntimeError("9-bit packed transport: decode of the packed streams differs from the block path")
        e2e["duplex_p9"] = duplex(lambda: plan.encode_emit_packed(px_h.array, q, 9, out=p9[0].array),
                                  lambda: dplan.decode_emit_packed(p9[1].array, q, 9, out=out_h.array))
        e2e_bytes["p9"] = (px_h.array.nbytes + p9[1].array.nbytes, p9[0].array.nbytes + out_h.array.nbytes)
        for b_ in p9:
            b_.free()
        # 10-bit packed emission streams
        nb = plan.emission_packed_bytes()
        pk_enc = capi.PinnedBuffer((1, C, nb), np.uint8)
        pk_dec = capi.PinnedBuffer((1, C, nb), np.uint8)
        plan.encode_emit10(px_h.array, q, out=pk_enc.array)
        pk_dec.array[...] = pk_enc.array
        dplan.decode_emit10(pk_dec.array, q, out=out_h.array)
        if not np.array_equal(out_h.array, plan.decode(plan.encode(px_h.array, q), q)):
            raise RuntimeError("packed transport: decode of the packed streams differs from the block path")
        if world > 1:
            dist.barrier()  # every rank copies at the same time: the host's aggregate PCIe rate is the shared resource
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.encode_emit10(px_h.array, q, out=pk_enc.array)
            plan.decode_emit10(pk_enc.array, q, out=out_h.array)
        e2e["serial_p10"] = time.perf_counter() - t0
        e2e["duplex_p10"] = duplex(lambda: plan.encode_emit10(px_h.array, q, out=pk_enc.array),
                                   lambda: dplan.decode_emit10(pk_dec.array, q, out=out_h.array))
        e2e_bytes["p10"] = (px_h.array.nbytes + pk_dec.array.nbytes, pk_enc.array.nbytes + out_h.array.nbytes)
        # quad: TWO encoder threads and TWO decoder threads, one handle and one set of pinned buffers each — while one
        # call of a direction runs its kernels, the other's copies keep the link busy.  Half the steps per thread:
        # the same number of frames as the duplex run.
        plan2, dplan2 = capi.Plan(W, H, C, device=local_rank), capi.Plan(W, H, C, device=local_rank)
        px_h2, out_h2 = capi.PinnedBuffer((1, H, W, C), np.uint8), capi.PinnedBuffer((1, H, W, C), np.uint8)
        pk_enc2, pk_dec2 = capi.PinnedBuffer((1, C, nb), np.uint8), capi.PinnedBuffer((1, C, nb), np.uint8)
        px_h2.array[...] = px_h.array
        pk_dec2.array[...] = pk_dec.array
        plan2.encode_emit10(px_h2.array, q, out=pk_enc2.array)
        dplan2.decode_emit10(pk_dec2.array, q, out=out_h2.array)
        q9 = [capi.PinnedBuffer((1, C, nb9), np.uint8) for _ in range(4)]  # enc, enc2, dec, dec2
        plan.encode_emit_packed(px_h.array, q, 9, out=q9[0].array)
        q9[2].array[...] = q9[0].array
        q9[3].array[...] = q9[0].array

        def quad_once(bits: int = 10) -> float:
            errors = []

            def loop(call):
                try:
                    for _ in range(e2e_steps // 2):
                        call()
                except Exception as exc:
                    errors.append(exc)

            if bits == 10:
                calls = [lambda: plan.encode_emit10(px_h.array, q, out=pk_enc.array),
                         lambda: plan2.encode_emit10(px_h2.array, q, out=pk_enc2.array),
                         lambda: dplan.decode_emit10(pk_dec.array, q, out=out_h.array),
                         lambda: dplan2.decode_emit10(pk_dec2.array, q, out=out_h2.array)]
            else:
                calls = [lambda: plan.encode_emit_packed(px_h.array, q, 9, out=q9[0].array),
                         lambda: plan2.encode_emit_packed(px_h2.array, q, 9, out=q9[1].array),
                         lambda: dplan.decode_emit_packed(q9[2].array, q, 9, out=out_h.array),
                         lambda: dplan2.decode_emit_packed(q9[3].array, q, 9, out=out_h2.array)]
            if world > 1:
                dist.barrier()
            th = [threading.Thread(target=loop, args=(c,)) for c in calls]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
            if errors:
                raise errors[0]
            return dt

        e2e["quad_p10"] = statistics.median(quad_once() for _ in range(3))
        e2e["quad_p9"] = statistics.median(quad_once(9) for _ in range(3))
        if not np.array_equal(q9[1].array, q9[0].array):
            raise RuntimeError("quad e2e: the two encoder handles disagree (9-bit streams)")
        for b_ in q9:
            b_.free()
        if not np.array_equal(pk_enc2.array, pk_enc.array) or not np.array_equal(out_h2.array, out_h.array):
            raise RuntimeError("quad e2e: the two handles of a direction disagree")
        for b in (px_h2, out_h2, pk_enc2, pk_dec2):
            b.free()
        plan2.close(); dplan2.close()
        pk_enc.free(); pk_dec.free()
        dplan.close()
        px_h.free(); out_h.free()
    e2e_keys = sorted(e2e)

    vals = [elapsed_ms, enc_ms, dec_ms, enc_iso, dec_iso, v_enc, v_dec] + ([batched[3], batched[4]] if batched else [0.0, 0.0])
    vals += [o_enc, o_dec, o_pair]
    vals += [e2e[k] for k in e2e_keys]
    if world > 1:
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(x) for x in t.tolist()]
        dist.barrier()
    elapsed_ms, enc_ms, dec_ms, enc_iso, dec_iso, v_enc, v_dec, b0, b1, o_enc, o_dec, o_pair = vals[:12]
    e2e = dict(zip(e2e_keys, vals[12:]))
    if batched:
        batched = batched[:3] + (b0, b1)

    if rank == 0:
        total_frames = BATCH_FRAMES if batch else FRAMES * world
        pix_step = W * H * total_frames  # pixels through encode+decode per step, all GPUs
        value = pix_step * args.steps / (elapsed_ms * 1e-3) / 1e6
        peak, peak_src = measured_peak_gbs()
        alg_bytes = W * H * C * frames * BYTES_PER_SAMPLE  # per launch on one GPU (rank 0's shard)

        def roof(ms: float, iso_ms: float, kernel: str) -> dict:
            ach = alg_bytes / (ms * 1e-3) / 1e9
            return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": ncu_traffic(kernel) if not batch else None, "algorithmic_bytes": alg_bytes, "avg_launch_ms": ms,
                    "timing": f"{reps} back-to-back launches between two CUDA events on the launching stream, inputs rotating "
                              f"over {n_sets} buffer set(s); independent of --steps",
                    "isolated_launch_ms": iso_ms, "isolated_frac": alg_bytes / (iso_ms * 1e-3) / 1e9 / peak,
                    "frac_of_nominal_8TBs": ach / 8000.0, "peak_source": peak_src}

        r_enc = roof(enc_ms, enc_iso, f"fri_encode_kernel<{C},u8>")
        r_dec = roof(dec_ms, dec_iso, f"fri_decode_kernel<{C},u8>")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong" if batch else "weak",
            "vs_baseline": None, "dtype": "i32", "data": "synthetic", "config": workload_config(world, args.workload),
            "encode_mpix_s": pix_step / (enc_ms * 1e-3) / 1e6, "decode_mpix_s": pix_step / (dec_ms * 1e-3) / 1e6,
            "roofline": r_enc if enc_ms >= dec_ms else r_dec, "roofline_encode": r_enc, "roofline_decode": r_dec,
            "gpu_launches": timed_launches, "clocks": clocks, "launch": plan.launch_info(),
            "plan_build_ms": plan_build_ms, "emission_order_build_ms": emission_build_ms,
            "plan_note": "host work once per image size (lattice BFS, retain, chunk tables, upload; emission order = the "
                         "reference's sort_lattice scan, needed by the *_emit* entry points only); outside the timed region",
        }
        if e2e:
            best = min((k for k in e2e if k.startswith(("duplex", "async", "quad"))), key=lambda k: e2e[k])
            fmt = best.split("_")[1]
            api = {"i32": "fri_encode_tq + fri_decode_tq (int32 coefficient blocks on the host side)",
                   "i16": "fri_encode_tq16 + fri_decode_tq16 (int16 coefficient blocks on the host side)",
                   "p9": "fri_encode_tq_emit_packed + fri_decode_tq_emit_packed at 9 bits per symbol (emission-ordered streams; every "
                         "coefficient of an 8-bit image fits: |k| <= 255)",
                   "p10": "fri_encode_tq_emit10 + fri_decode_tq_emit10 (emission-ordered streams, 10-bit packed symbols on the "
                          "host side: what the reference's entropy coder consumes / produces)"}[fmt]
            how = {"quad": "two encoder threads and two decoder threads with one plan handle each",
                   "duplex": "encoder thread and decoder thread with one plan handle each",
                   "async": "one thread, both handles in asynchronous mode", "serial": "one thread"}[best.split("_")[0]]
            # one frame per GPU per e2e step
            line["e2e"] = {
                "value": W * H * world * e2e_steps / e2e[best] / 1e6, "unit": UNIT,
                "h2d_bytes_per_step": e2e_bytes[fmt][0], "d2h_bytes_per_step": e2e_bytes[fmt][1], "steps": e2e_steps,
                "api": f"{api}, pinned host buffers, {how}; one {W}x{H}x{C} frame per GPU and step",
                "variant": best,
                "variants_mpix_s": {k: W * H * world * e2e_steps / v / 1e6 for k, v in e2e.items()},
                "variants": "serial = one thread, encode then decode; duplex = encoder and decoder threads; quad = two of each; async = one thread, "
                            "both handles asynchronous; i32 / i16 = coefficient blocks (4 / 2 B per coefficient over PCIe), "
                            "p10 / p9 = emission-ordered streams packed at 10 / 9 bits per symbol (1.25 / 1.125 B per coefficient)",
                "limiter": "host<->device PCIe copies (both directions busy); kernels are a few percent of the step"}
        if o_pair > 0:
            line["stream_overlap"] = {
                "note": "same launches with fri_plan_set_independent_calls(1): the caller promises that consecutive calls on "
                        "the stream touch disjoint buffers (true here: rotating buffer sets), the kernels skip the "
                        "programmatic-dependency wait and fill each other's last wave; opt-in, NOT the headline",
                "value": pix_step / (o_pair * 1e-3) / 1e6, "unit": UNIT,
                "encode": {"achieved": alg_bytes / (o_enc * 1e-3) / 1e9, "frac": alg_bytes / (o_enc * 1e-3) / 1e9 / peak, "avg_launch_ms": o_enc},
                "decode": {"achieved": alg_bytes / (o_dec * 1e-3) / 1e9, "frac": alg_bytes / (o_dec * 1e-3) / 1e9 / peak, "avg_launch_ms": o_dec},
                "unit_bw": "GB/s"}
        if not batch:
            b16 = W * H * C * frames * 3  # u8 pixel + i16 coefficient
            line["int16_arrays"] = {
                "note": "same step through fri_encode_tq_device16 / fri_decode_tq_device16 (int16 coefficient arrays, "
                        "3 B per sample); a separate variant, not the headline",
                "value": pix_step / ((v_enc + v_dec) * 1e-3) / 1e6, "unit": UNIT,
                "encode": {"achieved": b16 / (v_enc * 1e-3) / 1e9, "frac": b16 / (v_enc * 1e-3) / 1e9 / peak, "avg_launch_ms": v_enc},
                "decode": {"achieved": b16 / (v_dec * 1e-3) / 1e9, "frac": b16 / (v_dec * 1e-3) / 1e9 / peak, "avg_launch_ms": v_dec},
                "algorithmic_bytes": b16, "unit_bw": "GB/s"}
        if batched:
            bw, bh, bf, b_enc, b_dec = batched
            bbytes = bw * bh * C * bf * BYTES_PER_SAMPLE
            line["batched"] = {
                "workload": f"{bf} x {bw}x{bh}x{C} u8 frames per GPU in one launch per direction (BASELINE.json configs[2] "
                            f"at 8 GPUs); {bbytes / 1e9:.1f} GB per launch, far beyond L2",
                "value": bw * bh * bf * world / ((b_enc + b_dec) * 1e-3) / 1e6, "unit": UNIT,
                "encode": {"achieved": bbytes / (b_enc * 1e-3) / 1e9, "frac": bbytes / (b_enc * 1e-3) / 1e9 / peak,
                           "avg_launch_ms": b_enc},
                "decode": {"achieved": bbytes / (b_dec * 1e-3) / 1e9, "frac": bbytes / (b_dec * 1e-3) / 1e9 / peak,
                           "avg_launch_ms": b_dec},
                "peak": peak, "unit_bw": "GB/s"}
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle as O
            centers = plan.centers()
            coef_buf = np.empty(plan.coef_shape, np.int32)
            out_buf = np.zeros((H, W, C), np.uint8)
            t_cpu, passes = 0.0, 0
            while t_cpu < 10.0:
                t_cpu += cpu_pass(O, img0, centers, None, q, 1, None, coef_buf, out_buf)
                passes += 1
            # the reference's real cost is dominated by the containers Fractal::new builds
            # per tile (hash maps, Vecs), which the arithmetic-only port leaves out; a cost model of those
            # (oracle/fri_oracle.c, fri_oracle_fractal_new_cost) on a tile sample, once per direction
            # (from_raster on encode, from_metadata on decode), gives a clearly labelled estimate
            n_sample = min(len(centers), 4000)
            t0 = time.perf_counter()
            O.fractal_new_cost(centers[:n_sample], C)
            t_new = (time.perf_counter() - t0) * len(centers) / n_sample
            line["cpu_baseline"] = {
                "value": W * H * passes / t_cpu / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"{passes} full pass(es) of the same {W}x{H}x{C} image (encode+decode), C oracle, 1 thread "
                          f"(the reference is single-threaded)",
                "reference_shaped_estimate": {
                    "value": W * H / (t_cpu / passes + 2 * t_new) / 1e6, "unit": UNIT,
                    "note": "ESTIMATE, not a measurement of the reference: arithmetic pass + 2 x a cost model of "
                            "Fractal::new's per-tile containers (SipHash-1-3 HashMap inserts, Vec allocations; "
                            f"{n_sample} tiles sampled, {t_new:.2f} s per image and direction)"}}
        if world == 1 and not args.no_codec and not batch and C in (1, 3):
            # the whole codec behind the transform through the reference-shaped call
            # FRIEncoder::encode / FRIDecoder::decode: pixels -> `frif` bytes -> pixels.  A smooth synthetic
            # image (the residual alphabet is 1024 symbols; uniform noise is not what the codec is for).
            yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
            rng = np.random.Generator(np.random.PCG64(11))
            sm = np.empty((H, W, C), np.uint8)
            for ch in range(C):
                v = 0.5 + 0.375 * np.sin(2 * np.pi * xx / W * 3 + ch) * np.cos(2 * np.pi * yy / H * 2 + 0.5 * ch)
                sm[:, :, ch] = np.clip(np.rint(v * 255 + rng.normal(0, 2, size=(H, W))), 0, 255).astype(np.uint8)
            ones = np.ones(32, np.int32)
            data = plan.frv_encode(sm, ones)  # warm-up: tables, pinned staging
            t_enc, t_dec = [], []
            for _ in range(3):
                t0 = time.perf_counter(); data = plan.frv_encode(sm, ones); t_enc.append(time.perf_counter() - t0)
            for _ in range(2):
                t0 = time.perf_counter(); back = plan.frv_decode(data, ones); t_dec.append(time.perf_counter() - t0)
            line["codec"] = {
                "note": "whole pipeline, NOT the headline: device transform + parameter-fit sums + prediction, host solve + rANS + "
                        "`frif` container (fri_frv_encode), serial host entropy decoding + device inverse (fri_frv_decode); "
                        "all-ones quantization matrix (the reference's), smooth synthetic image; parity of the container "
                        "bytes with the reference is unpinned",
                "encode_mpix_s": W * H / sorted(t_enc)[1] / 1e6, "decode_mpix_s": W * H / min(t_dec) / 1e6,
                "bits_per_pixel": 8 * len(data) / (W * H), "lossless": bool(np.array_equal(back, sm)) if plan.pixels_covered == W * H else None,
                "unit": UNIT}
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def run_image_split(args, rank: int, local_rank: int, world: int) -> None:
    """--workload image16k: ONE 16384x16384 16-bit single-channel image (BASELINE.json configs[3], depth 9) split over
    the GPUs by contiguous ranges of tile groups.  Strong scaling.  A step on every rank: transform the
    rank's tiles from its band of pixel rows, invert them into a band that starts from zero in the rows shared with
    a neighbour, then the path's one exchange step: the overlap rows (tiles straddle the cut) go to the neighbour
    and are merged by addition (NCCL send / recv).  Device-resident; the max over ranks is the step time."""
    import torch
    import torch.distributed as dist

    from frave_b200 import capi, sharding

    if capi.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: libfri_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    iw, ih = (16384, 16384) if not args.shape else tuple(int(v) for v in args.shape.lower().split("x")[:2])
    q = np.ones(32, np.int32)  # 16-bit extension: transform only (no container), the reference's all-ones matrix
    t0 = time.perf_counter()
    plan = capi.Plan(iw, ih, 1, sample_bytes=2, device=local_rank)
    plan_build_ms = 1e3 * (time.perf_counter() - t0)
    part = sharding.shard_image(plan, rank, world)
    shared = sharding.overlaps(plan, rank, world)
    t_lo, t_hi, r0, r1 = part["tile_begin"], part["tile_end"], part["row_begin"], part["row_end"]
    stream = torch.cuda.current_stream().cuda_stream
    # every rank generates the same image row by row block (seeded per block of rows) and keeps only its band
    n_sets = 2
    block_rows = 1024
    bands = []
    for s_ in range(n_sets):
        band = torch.empty((r1 - r0, iw), dtype=torch.int16, device=dev)
        for b0 in range(r0 // block_rows * block_rows, r1, block_rows):
            g_ = torch.Generator(device=dev).manual_seed(77 + 1000 * s_ + b0)
            rows = torch.randint(0, 65536, (block_rows, iw), generator=g_, device=dev, dtype=torch.int32).to(torch.int16)
            lo, hi = max(b0, r0), min(b0 + block_rows, r1)
            band[lo - r0:hi - r0] = rows[lo - b0:hi - b0]
        bands.append(band)
    coefs = [torch.empty((t_hi - t_lo, 1, 512), dtype=torch.int32, device=dev) for _ in range(n_sets)]
    outs = [torch.empty((r1 - r0, iw), dtype=torch.int16, device=dev) for _ in range(n_sets)]
    recv = {peer: torch.empty((hi - lo, iw), dtype=torch.int16, device=dev) for peer, lo, hi in shared}
    launches = 0

    def step_nccl(i: int) -> None:
        """exchange by NCCL: shared rows start from zero, point-to-point send / receive, merge by addition"""
        nonlocal launches
        a = i % n_sets
        plan.encode_device_part(bands[a].data_ptr(), coefs[a].data_ptr(), rank, world, q, stream)
        launches += plan.last_launches
        for _, lo, hi in shared:  # the rows a neighbour also writes into start from zero
            outs[a][lo - r0:hi - r0].zero_()
        plan.decode_device_part(coefs[a].data_ptr(), outs[a].data_ptr(), rank, world, q, False, stream)
        launches += plan.last_launches
        if shared:
            ops = []
            for peer, lo, hi in shared:
                ops.append(dist.P2POp(dist.isend, outs[a][lo - r0:hi - r0].view(torch.uint8), peer))  # (NCCL has no int16)
                ops.append(dist.P2POp(dist.irecv, recv[peer].view(torch.uint8), peer))
            for w_ in dist.batch_isend_irecv(ops):
                w_.wait()
            for peer, lo, hi in shared:
                outs[a][lo - r0:hi - r0] += recv[peer]  # a pixel has one owner: the other side holds zero there

    # ---- exchange by the transform kernel's own stores over peer memory: the bands live in symmetric memory (every
    # rank maps every other rank's band over NVLink); a rank re-runs its groups along each cut with the NEIGHBOUR's
    # band as the target (fri_decode_tq_device_groups), so the neighbour's shared rows are completed by P2P stores
    # of the kernel that computed them; no zeroing, no merge pass, one device-side barrier per step
    peer_err, step_peer, p_outs = None, None, None
    if world > 1 and args.exchange in ("auto", "peer"):
        try:
            import torch.distributed._symmetric_memory as symm_mem
            margin = plan.launch_info()["region_h"]
            parts = [sharding.shard_image(plan, r, world) for r in range(world)]
            pushes = sharding.halo_pushes(plan, rank, world)
            for ps in pushes:
                pp = parts[ps["peer"]]
                if not (pp["row_begin"] - margin <= ps["span_begin"] and ps["span_end"] <= pp["row_end"] + margin):
                    raise RuntimeError("margin too small for the groups along the cut")
            rows_max = max(p_["row_end"] - p_["row_begin"] for p_ in parts) + 2 * margin
            sym = symm_mem.empty((n_sets, rows_max, iw), dtype=torch.int16, device=dev)
            hdl = symm_mem.rendezvous(sym, dist.group.WORLD)
            set_bytes = rows_max * iw * 2
            stride = iw * 2
            p_outs = [sym[a_][margin:margin + r1 - r0] for a_ in range(n_sets)]

            def step_peer(i: int) -> None:
                nonlocal launches
                a = i % n_sets
                plan.encode_device_part(bands[a].data_ptr(), coefs[a].data_ptr(), rank, world, q, stream)
                launches += plan.last_launches
                for ps in pushes:  # first, so that the remote stores are in flight while the own tiles are decoded
                    target = int(hdl.buffer_ptrs[ps["peer"]]) + a * set_bytes  # the peer's band incl. its margin
                    plan.decode_device_groups(coefs[a].data_ptr(), t_lo, target, parts[ps["peer"]]["row_begin"] - margin,
                                              ps["first"], ps["last"], q, False, stream)
                    launches += plan.last_launches
                plan.decode_device_part(coefs[a].data_ptr(), p_outs[a].data_ptr(), rank, world, q, False, stream)
                launches += plan.last_launches
                hdl.barrier(channel=0)  # every rank's pushes have landed

            for a_ in range(n_sets):
                sym[a_].fill_(12345)  # no zeroing in this variant: stale data must be overwritten by an owner
            torch.cuda.synchronize()
            dist.barrier()  # (a neighbour's first push must not land before this fill)
            step_peer(0)
            torch.cuda.synchronize()
            if plan.pixels_covered == iw * ih and not torch.equal(p_outs[0], bands[0]):
                raise RuntimeError("peer exchange: assembled band differs from the input")
        except Exception as e:  # noqa: BLE001 — symmetric memory may be unavailable on a box; the NCCL variant stands
            peer_err, step_peer = f"{type(e).__name__}: {e}"[:300], None
            if args.exchange == "peer":
                raise
    ok = torch.tensor([1 if step_peer is not None else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks or none
    if ok.item() == 0:
        step_peer = None

    # sanity (untimed): the assembled band equals the input (lossless at the all-ones matrix)
    step_nccl(0)
    torch.cuda.synchronize()
    if plan.pixels_covered == iw * ih and not torch.equal(outs[0], bands[0]):
        raise RuntimeError("sanity check failed: the split encode -> decode -> exchange is not lossless")
    sampler = ClockSampler(local_rank) if rank == 0 else None

    def timed(step) -> tuple[float, int]:
        nonlocal launches
        for k in range(max(args.warmup, 3)):
            step(k)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        start.record()
        for k in range(args.steps):
            step(k)
        end.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return start.elapsed_time(end), launches

    variants = {}
    if args.exchange != "peer" or step_peer is None:
        variants["nccl"] = timed(step_nccl)
    if step_peer is not None:
        variants["peer"] = timed(step_peer)
    tv = torch.tensor([variants.get(k, (0.0, 0))[0] for k in ("nccl", "peer")], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
    for k, v in zip(("nccl", "peer"), tv.tolist()):
        if k in variants:
            variants[k] = (v, variants[k][1])
    best = min(variants, key=lambda k: variants[k][0])
    elapsed_ms, timed_launches = variants[best]

    def enc_only(i: int) -> None:
        plan.encode_device_part(bands[i % n_sets].data_ptr(), coefs[i % n_sets].data_ptr(), rank, world, q, stream)

    def dec_only(i: int) -> None:
        plan.decode_device_part(coefs[i % n_sets].data_ptr(), outs[i % n_sets].data_ptr(), rank, world, q, False, stream)

    reps = 40
    enc_ms, dec_ms = b2b(enc_only, reps, torch), b2b(dec_only, reps, torch)
    clocks = sampler.stop() if sampler else None
    vals = torch.tensor([enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    enc_ms, dec_ms = vals.tolist()
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        alg = (t_hi - t_lo) * 512 * 6  # this rank's samples x (2 B pixel + 4 B coefficient)

        def roof(ms: float, kernel: str) -> dict:
            ach = alg / (ms * 1e-3) / 1e9
            return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "algorithmic_bytes": alg, "avg_launch_ms": ms, "peak_source": peak_src,
                    "timing": f"{reps} back-to-back launches of rank 0's part between two CUDA events (max over ranks)"}

        r_enc, r_dec = roof(enc_ms, "fri_encode_kernel<1,u16>"), roof(dec_ms, "fri_decode_kernel<1,u16>")
        halo = sum(hi - lo for _, lo, hi in shared)
        line = {
            "metric": METRIC, "value": iw * ih * args.steps / (elapsed_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": {"workload": f"one {iw}x{ih}x1 u16 synthetic image (BASELINE.json configs[3], depth 9) split over {world} GPU(s) by "
                                   "contiguous ranges of tile groups; step = encode + decode of every rank's tiles + exchange of the "
                                   "overlap rows between neighbours (NCCL send/recv, merged by addition)",
                       "parts": world, "tiles_rank0": t_hi - t_lo, "rows_rank0": [r0, r1], "halo_rows_rank0": halo,
                       "exchange": {"nccl": "shared rows zeroed, NCCL send/recv, merged by addition",
                                    "peer": "bands in symmetric memory; every rank re-runs its groups along the cut with the neighbour's "
                                            "band as the target of the decode kernel (P2P stores over NVLink), one device-side "
                                            "barrier per step; no zeroing, no merge"}[best],
                       "quant": "all ones (16-bit extension: transform only)",
                       "l2": f"{n_sets} rotating buffer sets; rank 0's set is {(r1 - r0) * iw * 2 * 2 + (t_hi - t_lo) * 2048 >> 20} MB",
                       "parallelism": f"tile groups of one image sharded over {world} GPU(s); one point-to-point exchange of "
                                      f"{halo} overlap rows per rank and step"},
            "roofline": r_enc if enc_ms >= dec_ms else r_dec, "roofline_encode": r_enc, "roofline_decode": r_dec,
            "gpu_launches": timed_launches, "clocks": clocks, "launch": plan.launch_info(), "plan_build_ms": plan_build_ms,
            "exchanges_ms_per_step": {k: v[0] / args.steps for k, v in variants.items()}, "peer_exchange_error": peer_err,
            "e2e": None, "cpu_baseline": None,
            "note": "secondary workload (the driver's line is the default one): e2e / cpu_baseline are reported there"}
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="frame", choices=["frame", "batch256", "image16k"],
                    help="frame = BASELINE.json configs[1] (default); batch256 = configs[2], 256 4K frames sharded over the ranks; "
                         "image16k = configs[3] at depth 9, ONE 16384x16384 u16 image split over the ranks by tile-group ranges")
    ap.add_argument("--shape", default=None, help="WxHxC override for experiments (default: BASELINE.json configs[1])")
    ap.add_argument("--frames", type=int, default=1, help="frames per GPU per step (batched launch)")
    ap.add_argument("--batch-frames", type=int, default=256, help="--workload batch256: total frames (experiments only)")
    ap.add_argument("--preheat", type=float, default=1.0, help="seconds of untimed load before the warm-up steps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (experiments only)")
    ap.add_argument("--no-batched", action="store_true", help="skip the batched steady-state leg (experiments only)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (experiments only)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "peer"],
                    help="--workload image16k: how the overlap rows reach the neighbour (auto = time both, report the faster)")
    ap.add_argument("--no-codec", action="store_true", help="skip the whole-codec (frif container) leg (experiments only)")
    ap.add_argument("--divisor", type=int, default=None, help="smallest-layer divisor override (experiments only; default 4)")
    args = ap.parse_args()
    global W, H, C, FRAMES, PREHEAT_S, SMALLEST_LAYER_DIVISOR, BATCH_FRAMES
    if args.divisor is not None:
        SMALLEST_LAYER_DIVISOR = args.divisor
    if args.workload == "batch256":
        W, H, C = 3840, 2160, 3
        BATCH_FRAMES = args.batch_frames
        if args.steps == 500:
            args.steps = 5  # a step is 256 frames (25 GB of traffic per direction)
    if args.shape:
        W, H, C = (int(v) for v in args.shape.lower().split("x"))
    FRAMES, PREHEAT_S = args.frames, args.preheat
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch N > 1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    if args.workload == "image16k":
        run_image_split(args, rank, local_rank, world)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
