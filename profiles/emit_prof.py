"""Times the widening kernels (emission gather / un-gather, 10-bit pack, prediction + context bucketing) on
4096x4096x3, L2 flushed before every launch; also the target of `ncu -k regex:fri_(un)?emit|fri_predict`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from frave_b200 import capi

dev = torch.device("cuda", 0)
W = H = 4096
C = 3
plan = capi.Plan(W, H, C)
cnt, nb = plan.emission_count(), plan.emission_packed_bytes()
co = torch.randint(-255, 256, plan.coef_shape, device=dev, dtype=torch.int32)
out = torch.empty((1, C, cnt), dtype=torch.int32, device=dev)
out16 = torch.empty((1, C, cnt), dtype=torch.int16, device=dev)
pk = torch.empty((1, C, nb), dtype=torch.uint8, device=dev)
bk = torch.empty((1, C, cnt), dtype=torch.uint8, device=dev)
pr = torch.empty((1, C, cnt), dtype=torch.int32, device=dev)
sy = torch.empty((1, C, cnt), dtype=torch.int16, device=dev)
hi = torch.empty((1, C, 10, 1024), dtype=torch.int32, device=dev)
ov = torch.empty((1,), dtype=torch.int32, device=dev)
vp = np.tile(np.array([0.4, 0.1, 0.1, 0.2, 0.1, 0.1], np.float32), (C, 3, 1))
wp = np.full((C, 3, 6), 0.5, np.float32)
junk = torch.empty(1 << 28, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 1):
        junk.zero_()  # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return round(sorted(ts[1:])[len(ts[1:]) // 2], 1)


res = {"shape": f"{W}x{H}x{C}", "count_per_channel": cnt, "packed_bytes_per_channel": nb}
res["emit_i32_us"] = timed(lambda: plan.emit_device(co.data_ptr(), 1, out.data_ptr()))
res["emit_i16_us"] = timed(lambda: plan.emit_device(co.data_ptr(), 1, out16.data_ptr(), half=True))
res["emit_p10_us"] = timed(lambda: plan.emit_device10(co.data_ptr(), 1, pk.data_ptr()))
res["unemit_i32_us"] = timed(lambda: plan.unemit_device(out.data_ptr(), 1, co.data_ptr()))
res["unemit_p10_us"] = timed(lambda: plan.unemit_device10(pk.data_ptr(), 1, co.data_ptr()))
res["predict_us"] = timed(lambda: plan.predict_device(co.data_ptr(), 1, vp, wp, bk.data_ptr(), pr.data_ptr(), sy.data_ptr(),
                                                      hi.data_ptr(), ov.data_ptr()))
res["predict_MPix_s"] = round(W * H / res["predict_us"], 0)
print(json.dumps(res))
