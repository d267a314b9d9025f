#!/usr/bin/env python
"""Back-to-back launch timing of the transform kernels (no events between launches): R encode launches,
R decode launches, R encode+decode pairs, rotating over buffer sets larger than L2.  Used for same-box A/B
of launch-level changes (FRI_PDL=0/1, variant libraries via FRI_CUDA_LIB).

    python profiles/exp_b2b.py [--shape WxHxC] [--frames F] [--reps R] [--divisor D] [--half]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frave_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="4096x4096x3")
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--reps", type=int, default=300)
ap.add_argument("--divisor", type=int, default=4)
ap.add_argument("--sets", type=int, default=4)
ap.add_argument("--half", action="store_true")
ap.add_argument("--sample-bytes", type=int, default=1)
ap.add_argument("--depth", type=int, default=9)
ap.add_argument("--tag", default="")
a = ap.parse_args()
W, H, C = (int(v) for v in a.shape.split("x"))
dev = torch.device("cuda", 0)
q = np.ones(32, np.int32)
q[8] = q[9] = a.divisor
plan = capi.Plan(W, H, C, depth=a.depth, sample_bytes=a.sample_bytes, device=0)
st = torch.cuda.current_stream().cuda_stream
gen = torch.Generator(device=dev).manual_seed(5)
pdt = torch.uint8 if a.sample_bytes == 1 else torch.int16
px = [torch.randint(0, 256, (a.frames, H, W, C * a.sample_bytes), generator=gen, device=dev, dtype=torch.int32).to(torch.uint8)
      for _ in range(a.sets)]
cdt = torch.int16 if a.half else torch.int32
co = [torch.empty((a.frames,) + plan.coef_shape, dtype=cdt, device=dev) for _ in range(a.sets)]
out = [torch.empty_like(px[0]) for _ in range(a.sets)]
for s in range(a.sets):
    plan.encode_device(px[s].data_ptr(), a.frames, co[s].data_ptr(), q, st, half=a.half)
torch.cuda.synchronize()


def enc(i):
    plan.encode_device(px[i % a.sets].data_ptr(), a.frames, co[i % a.sets].data_ptr(), q, st, half=a.half)


def dec(i):
    plan.decode_device(co[i % a.sets].data_ptr(), a.frames, out[i % a.sets].data_ptr(), q, False, st, half=a.half)


def pair(i):
    enc(i)
    dec(i + a.sets // 2)


def timed(fn, reps):
    for i in range(20):
        fn(i)
    torch.cuda.synchronize()
    best = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / reps * 1e3)
    return sorted(best)[1]


def isolated(fn, reps):
    ts = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        ts.append((e0, e1))
    torch.cuda.synchronize()
    v = sorted(x.elapsed_time(y) * 1e3 for x, y in ts)
    return v[len(v) // 2]


bps = (a.sample_bytes + (2 if a.half else 4)) * W * H * C * a.frames
res = {"tag": a.tag, "pdl": os.environ.get("FRI_PDL", "1"), "lib": os.path.basename(capi.lib_path()), "shape": a.shape,
       "frames": a.frames}
for name, fn in (("enc", enc), ("dec", dec)):
    t = timed(fn, a.reps)
    ti = isolated(fn, min(a.reps, 200))
    res[name + "_us"] = round(t, 2)
    res[name + "_gbs"] = round(bps / t / 1e3, 1)
    res[name + "_iso_us"] = round(ti, 2)
t = timed(pair, a.reps)
res["pair_us"] = round(t, 2)
res["pair_mpix"] = round(W * H * a.frames / t, 1)
print(json.dumps(res))
