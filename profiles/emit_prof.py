"""Runs the emission gather / un-gather a few times on 4096x4096x3 (target of `ncu -k regex:fri_(un)?emit`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frave_b200 import capi

dev = torch.device("cuda", 0)
plan = capi.Plan(4096, 4096, 3)
cnt = plan.emission_count()
co = torch.randint(-255, 256, plan.coef_shape, device=dev, dtype=torch.int32)
out = torch.empty((1, 3, cnt), dtype=torch.int32, device=dev)
junk = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
for _ in range(3):
    junk.zero_()  # flush L2
    plan.emit_device(co.data_ptr(), 1, out.data_ptr())
    junk.zero_()
    plan.unemit_device(out.data_ptr(), 1, co.data_ptr())
torch.cuda.synchronize()
print("ok")
