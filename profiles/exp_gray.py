"""1-channel group-shape sweep (FRI_GROUP / FRI_TILES_PER_WARP are read when the plan is built)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from frave_b200 import capi

dev = torch.device("cuda", 0)


def run(w, h, c, sb, frames, reps, grp, tpw):
    os.environ["FRI_GROUP"] = grp
    os.environ["FRI_TILES_PER_WARP"] = str(tpw)
    plan = capi.Plan(w, h, c, sample_bytes=sb)
    tdt = torch.uint8 if sb == 1 else torch.int16
    px = torch.randint(0, 256 if sb == 1 else 32767, (frames, h, w, c), device=dev, dtype=torch.int32).to(tdt)
    co = torch.empty((frames,) + plan.coef_shape, dtype=torch.int32, device=dev)
    out = torch.empty_like(px)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.encode_device(px.data_ptr(), frames, co.data_ptr(), None, st)
        plan.decode_device(co.data_ptr(), frames, out.data_ptr(), None, False, st)
    torch.cuda.synchronize()
    assert torch.equal(px, out)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    te = td = 0.0
    for _ in range(reps):
        e[0].record(); plan.encode_device(px.data_ptr(), frames, co.data_ptr(), None, st)
        e[1].record(); plan.decode_device(co.data_ptr(), frames, out.data_ptr(), None, False, st)
        e[2].record(); torch.cuda.synchronize()
        te += e[0].elapsed_time(e[1]); td += e[1].elapsed_time(e[2])
    te /= reps; td /= reps
    s = w * h * c * frames * (sb + 4)
    info = plan.launch_info()
    print(f"{w}x{h}x{c} sb={sb} f={frames} group={grp} tpw={tpw} thr={info['threads']} smem={info['smem_bytes']} groups={info['n_groups']}: "
          f"enc {te*1e3:.1f} us {s/te/1e6:.0f} GB/s, dec {td*1e3:.1f} us {s/td/1e6:.0f} GB/s")
    plan.close()


for grp, tpw in (("8x4", 4), ("8x4", 2), ("4x4", 2), ("8x2", 2), ("4x4", 1), ("4x2", 1), ("8x2", 1)):
    try:
        run(4096, 4096, 1, 1, 1, 30, grp, tpw)
        run(512, 512, 1, 1, 256, 10, grp, tpw)
        run(16384, 16384, 1, 2, 1, 5, grp if grp != "8x4" else "4x2", tpw) if False else None
    except Exception as ex:
        print(grp, tpw, "FAILED", ex)
