"""Random small shapes through the whole codec on the device: fri_frv_encode -> fri_frv_decode, device fit against host
fit, container bytes against the staged host pipeline.  Shapes whose emission order the reference itself cannot build
(FRI_E_UNSUPPORTED, wavelet_transform.rs:701) are counted, not failed.

    python profiles/fuzz_codec.py [n] [seed]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frave_b200 import capi
from tests.conftest import smooth_image, uniform_image

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ok = unsupported = 0
for i in range(n):
    w, h = (int(v) for v in rng.integers(1, 420, size=2))
    c = int(rng.choice([1, 3]))
    img = (smooth_image if rng.random() < 0.7 else uniform_image)(h, w, c, seed=i)
    q = np.ones(32, np.int32)
    if rng.random() < 0.5:
        q[8] = q[9] = int(rng.integers(2, 9))
    try:
        with capi.Plan(w, h, c) as p:
            data = p.frv_encode(img, q)
            rec = p.frv_decode(data, q)
            coefs = p.encode(img, q)[0]
            assert np.array_equal(p.frv_unpack(data), coefs), "unpack"
            d = torch.from_numpy(coefs).cuda()
            vp, wp = p.fit_device(d.data_ptr())
            hv, hw = p.fit_parameters(coefs)
            assert np.array_equal(vp.view(np.uint32), hv.view(np.uint32)) and np.array_equal(wp.view(np.uint32), hw.view(np.uint32)), "fit"
            b, _, s, hist, over = p.predict_host(coefs, vp, wp)
            assert over == 0 and data == p.frv_pack(vp, wp, b, s, hist), "bytes"
            assert np.array_equal(rec, p.decode(coefs[None], q)[0]), "decode"
            if (q == 1).all() and p.pixels_covered == w * h:
                assert np.array_equal(rec, img), "lossless"
        ok += 1
    except capi.FriError as e:
        if e.code == capi.FRI_E_UNSUPPORTED:
            unsupported += 1
        else:
            print("FAIL", w, h, c, e)
            raise
    except AssertionError as e:
        print("FAIL", w, h, c, q[8], e)
        raise
print(f"codec fuzz: {ok} shapes ok, {unsupported} unsupported by the reference's own scan, of {n}")
