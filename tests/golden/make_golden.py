"""Regenerates the golden fixtures in this directory.

The reference (pagmerek/frave) is Rust and cannot be built or imported in this environment, and
it ships no golden vectors of its own, so the fixtures are produced by the CPU oracle (oracle/
fri_oracle.c, cross-checked by oracle/fri_oracle_np.py) — they pin the oracle against silent
drift and give the GPU tests fixed known answers.  survey_kat.json is different: those digests
were produced by a third, throw-away restatement during the survey session
and are kept verbatim.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import c_oracle as O  # noqa: E402
from tests.conftest import smallest_layer_q, uniform_image  # noqa: E402

CASES = {
    # name: (h, w, c, dtype, seed, q)
    "u8_rgb_70x96_q5": (70, 96, 3, np.uint8, 11, smallest_layer_q(5)),
    "u8_luma_48x64_q1": (48, 64, 1, np.uint8, 12, np.ones(32, np.int32)),
    "u16_luma_65x33_q3": (65, 33, 1, np.uint16, 13, smallest_layer_q(3)),
}


def make(name):
    h, w, c, dtype, seed, q = CASES[name]
    img = uniform_image(h, w, c, seed, dtype)
    centers, coef, some = O.from_raster(img)
    qc = O.quantize(coef, some, q)
    dq = O.quantize(qc, some, q)  # the reference's decode divides again (quantization.rs:37)
    recon = O.extract_values(centers, dq, some, h, w, dtype=dtype)
    return dict(pixels=img, q=q, centers=centers, coef=qc, some=np.packbits(some[:, 0, :], axis=1, bitorder="little"),
                recon=recon)


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **make(name))
        print(name, os.path.getsize(os.path.join(HERE, name + ".npz")))
