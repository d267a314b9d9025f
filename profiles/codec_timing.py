"""Whole-pipeline timing of fri_frv_encode / fri_frv_decode (device transform + prediction, host fit + rANS +
container) on smooth synthetic images; prints one JSON line per shape."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from frave_b200 import capi
from tests.conftest import smooth_image

for (h, w, c) in ((512, 512, 1), (1080, 1920, 3), (2160, 3840, 3), (4096, 4096, 3)):
    img = smooth_image(h, w, c, seed=1)
    with capi.Plan(w, h, c) as p:
        p.frv_encode(img)  # warm-up: emission order, tables, slots
        t = time.perf_counter(); data = p.frv_encode(img); te = time.perf_counter() - t
        t = time.perf_counter(); rec = p.frv_decode(data); td = time.perf_counter() - t
        print(json.dumps({"shape": f"{w}x{h}x{c}", "bytes": len(data), "bits_per_pixel": round(8 * len(data) / (h * w), 3),
                          "encode_s": round(te, 3), "decode_s": round(td, 3), "encode_MPix_s": round(h * w / te / 1e6, 1),
                          "decode_MPix_s": round(h * w / td / 1e6, 1), "lossless": bool(np.array_equal(rec, img))}), flush=True)
