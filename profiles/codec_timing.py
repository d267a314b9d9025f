"""Whole-pipeline timing of fri_frv_encode / fri_frv_decode (device transform + fit sums + prediction, host solve +
rANS + container) on smooth synthetic images, with the encoder's stages timed one by one through the separate
entry points; prints one JSON line per shape."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from frave_b200 import capi
from tests.conftest import smooth_image

for (h, w, c) in ((512, 512, 1), (1080, 1920, 3), (2160, 3840, 3), (4096, 4096, 3)):
    img = smooth_image(h, w, c, seed=1)
    with capi.Plan(w, h, c) as p:
        p.frv_encode(img)  # warm-up: emission order, tables, slots
        t = time.perf_counter(); data = p.frv_encode(img); te = time.perf_counter() - t
        t = time.perf_counter(); rec = p.frv_decode(data); td = time.perf_counter() - t
        # stage by stage (wall clock around synchronous calls; the device stages include their stream sync)
        def wall(f):
            torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
            return r, round((time.perf_counter() - t0) * 1e3, 2)
        q = np.ones(32, np.int32)
        d_px = torch.from_numpy(img).cuda()
        d_co = torch.empty(p.coef_shape, dtype=torch.int32, device="cuda")
        n = p.emission_count()
        d_b = torch.empty((c, n), dtype=torch.uint8, device="cuda")
        d_p = torch.empty((c, n), dtype=torch.int32, device="cuda")
        d_s = torch.empty((c, n), dtype=torch.int16, device="cuda")
        d_h = torch.empty((c, 10, 1024), dtype=torch.int32, device="cuda")
        _, ms_tr = wall(lambda: p.encode_device(d_px.data_ptr(), 1, d_co.data_ptr(), q, 0))
        (vp, wp), ms_fit = wall(lambda: p.fit_device(d_co.data_ptr(), 0))
        coefs_h, ms_d2h_dense = wall(lambda: d_co.cpu().numpy())
        (vp_h, wp_h), ms_fit_host = wall(lambda: p.fit_parameters(coefs_h))
        _, ms_pred = wall(lambda: p.predict_device(d_co.data_ptr(), 1, vp, wp, d_b.data_ptr(), d_p.data_ptr(), d_s.data_ptr(),
                                                   d_h.data_ptr(), 0, 0))
        (b, s_, hh), ms_d2h = wall(lambda: (d_b.cpu().numpy(), d_s.cpu().numpy().view(np.uint16), d_h.cpu().numpy().view(np.uint32)))
        packed, ms_pack = wall(lambda: p.frv_pack(vp, wp, b, s_, hh))
        stages = {"transform_ms": ms_tr, "fit_device_ms": ms_fit, "fit_host_ms": ms_fit_host, "dense_d2h_ms": ms_d2h_dense,
                  "predict_ms": ms_pred, "symbols_d2h_ms": ms_d2h, "rans_container_ms": ms_pack,
                  "fit_identical": bool(np.array_equal(vp, vp_h) and np.array_equal(wp, wp_h)), "bytes_identical": packed == data}
        print(json.dumps({"stages": stages, "shape": f"{w}x{h}x{c}", "bytes": len(data), "bits_per_pixel": round(8 * len(data) / (h * w), 3),
                          "encode_s": round(te, 3), "decode_s": round(td, 3), "encode_MPix_s": round(h * w / te / 1e6, 1),
                          "decode_MPix_s": round(h * w / td / 1e6, 1), "lossless": bool(np.array_equal(rec, img))}), flush=True)
