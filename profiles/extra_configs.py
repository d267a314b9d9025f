"""Times the other BASELINE.json configs (parity-test cases, not bench lines) for DESIGN.md:
config 4 (16384x16384 u16 luma, depth 9 and deep trees) and 1-channel / u16 variants."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from frave_b200 import capi

dev = torch.device("cuda", 0)


def run(w, h, c, depth, sb, frames=1, reps=10, q=None, half=False):
    plan = capi.Plan(w, h, c, depth=depth, sample_bytes=sb)
    tdt = torch.uint8 if sb == 1 else torch.int16
    px = torch.randint(0, 256 if sb == 1 else 32767, (frames, h, w, c), device=dev, dtype=torch.int32).to(tdt)
    co = torch.empty((frames,) + plan.coef_shape, dtype=torch.int16 if half else torch.int32, device=dev)
    out = torch.empty_like(px)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        plan.encode_device(px.data_ptr(), frames, co.data_ptr(), q, st, half)
        plan.decode_device(co.data_ptr(), frames, out.data_ptr(), q, False, st, half)
    torch.cuda.synchronize()
    ok = torch.equal(px, out) if q is None else None
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    te = td = 0.0
    for _ in range(reps):
        e[0].record(); plan.encode_device(px.data_ptr(), frames, co.data_ptr(), q, st, half)
        e[1].record(); plan.decode_device(co.data_ptr(), frames, out.data_ptr(), q, False, st, half)
        e[2].record(); torch.cuda.synchronize()
        te += e[0].elapsed_time(e[1]); td += e[1].elapsed_time(e[2])
    te /= reps; td /= reps
    samples = w * h * c * frames
    bps = sb + (2 if half else 4)
    # coefficient traffic counts the blocks actually moved (deep trees include out-of-image base tiles)
    moved = plan.coefs_per_frame * frames * (2 if half else 4) + samples * sb
    print(json.dumps({"shape": f"{w}x{h}x{c}", "coefs": "i16" if half else "i32", "depth": depth, "sample_bytes": sb, "frames": frames, "tiles": plan.n_tiles,
                      "lossless": ok, "every_pixel_covered": plan.pixels_covered == w * h, "enc_us": round(te * 1e3, 1), "dec_us": round(td * 1e3, 1),
                      "enc_GBps_alg": round(samples * bps / te / 1e6), "dec_GBps_alg": round(samples * bps / td / 1e6),
                      "enc_GBps_moved": round(moved / te / 1e6), "dec_GBps_moved": round(moved / td / 1e6),
                      "enc_MPix_s": round(w * h * frames / te / 1e3), "dec_MPix_s": round(w * h * frames / td / 1e3),
                      "launches": plan.last_launches}))
    plan.close()


EMIT_ONLY = "--emit-only" in sys.argv
if EMIT_ONLY:
    run = lambda *a, **k: None  # noqa: E731
run(512, 512, 1, 9, 1, frames=1, reps=50)
run(512, 512, 1, 9, 1, frames=256, reps=10)
run(4096, 4096, 1, 9, 1)
run(1920, 1080, 3, 9, 1, frames=16)
run(16384, 16384, 1, 9, 2, reps=5)
for d in (16, 20, 24):
    run(16384, 16384, 1, d, 2, reps=3)
run(3840, 2160, 3, 9, 1, frames=32, reps=5)
run(3840, 2160, 3, 9, 1, frames=32, reps=5, half=True)
run(4096, 4096, 3, 9, 1, reps=20, half=True)


def run_emit(w, h, c, frames=1, reps=10, half=False):
    """next-1: emission-order gather on device-resident coefficients."""
    plan = capi.Plan(w, h, c)
    cnt = plan.emission_count()
    co = torch.randint(-255, 256, (frames,) + plan.coef_shape, device=dev, dtype=torch.int32)
    out = torch.empty((frames, c, cnt), dtype=torch.int16 if half else torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        plan.emit_device(co.data_ptr(), frames, out.data_ptr(), st, half)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        plan.emit_device(co.data_ptr(), frames, out.data_ptr(), st, half)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    a.record()
    for _ in range(reps):
        plan.unemit_device(out.data_ptr(), frames, co.data_ptr(), st, half)
    b.record()
    torch.cuda.synchronize()
    ms_un = a.elapsed_time(b) / reps
    bpc = 6 if half else 8
    print(json.dumps({"emit": f"{w}x{h}x{c}", "out": "i16" if half else "i32", "frames": frames, "count_per_channel": cnt,
                      "us": round(ms * 1e3, 1), "unemit_us": round(ms_un * 1e3, 1), f"GBps_alg({bpc}B per coefficient)": round(bpc * cnt * c * frames / ms / 1e6),
                      "MPix_s": round(w * h * frames / ms / 1e3)}))
    plan.close()


if "--emit-only" in sys.argv:
    pass
run_emit(4096, 4096, 3)
run_emit(4096, 4096, 3, half=True)
run_emit(3840, 2160, 3, frames=8)
