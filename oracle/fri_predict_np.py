"""Prediction + context bucketing of the quantized coefficients, encode side — CPU
restatement.

TEST INFRASTRUCTURE ONLY, PARITY UNPINNED (see oracle/fri_oracle.h): a literal restatement of the reference's
host-side code with Python dicts and numpy float32 scalars, written to be compared with the device kernel
(fri_predict_kernel, frave_b200/csrc/fri_predict.cu).  Paths are relative to /root/reference/.

  crates/libfri/src/stages/wavelet_transform.rs:71-90     get_nearby_vectors (hard-wired vectors for depth 1..3)
  crates/libfri/src/stages/wavelet_transform.rs:97-177    get_left / get_right / get_down_left / get_down_right /
                                                          get_up_right / get_up_left (depth-2 special cases)
  crates/libfri/src/context_modeling.rs:25-77             get_neighbour_values (3 same-level + 3 parent-level)
  crates/libfri/src/stages/prediction.rs:55-68            assign_bucket
  crates/libfri/src/stages/prediction.rs:86-149           get_lf_context_bucket (MED-style predictor)
  crates/libfri/src/stages/prediction.rs:151-207          get_hf_context_bucket (6-tap f32 predictor + width bucket)
  crates/libfri/src/stages/prediction.rs:224-323          encode: per-context symbol histograms
  crates/libfri/src/utils.rs:34-40                        pack_signed

The 18 + 18 f32 predictor parameters per channel are INPUTS here: the reference fits them with
lstsq 0.6 / nalgebra 0.33 (f32 SVD, context_modeling.rs:144-202), neither of which is in /root/reference.

Faithful quirks:
  * the depth-2 special cases index `global_position_map[depth]` with depth = 9 - level = 2, i.e. they look the
    LEVEL-7 candidate position up in the LEVEL-2 map (wavelet_transform.rs:127-129, :143-145, :159-161, :175-177);
  * a neighbour the lattice does not hold, or a `None` coefficient, contributes 0 (`unwrap_or(0)`);
  * f32 sums are evaluated left to right without fusing; `as u32` / `as i32` saturate and map NaN to 0.
"""
from __future__ import annotations

import numpy as np

from .fri_order_np import BASE_FRAC_DEPTH, image_positions
from .fri_oracle_np import LITERALS

CONTEXT_AMOUNT = 10   # prediction.rs:15
ALPHABET_SIZE = 1024  # entropy_coding.rs:25
F = np.float32


def nearby_vectors(depth: int):
    """wavelet_transform.rs:71-90, including the hard-wired small depths."""
    if depth == 1:
        zl, zmd = (-1, 1), (0, 2)
    elif depth == 2:
        zl, zmd = (-2, 0), (0, -2)
    elif depth == 3:
        zl, zmd = (-3, -1), (-1, -3)
    else:
        zl = (int(LITERALS[depth][0]), int(LITERALS[depth][1]))
        zmd = (int(LITERALS[depth + 1][0]) + zl[0], int(LITERALS[depth + 1][1]) + zl[1])
    sub = lambda a, b: (a[0] - b[0], a[1] - b[1])
    neg = lambda a: (-a[0], -a[1])
    return [zl, sub(zl, zmd), neg(zmd), neg(zl), sub(zmd, zl), zmd]


def _add(a, b):
    return (a[0] + b[0], a[1] + b[1])


def get_left(c, depth, gpm):  # :97-104
    return _add(c, nearby_vectors(depth)[4])


def get_right(c, depth, gpm):  # :106-113
    return _add(c, nearby_vectors(depth)[1])


def get_down_left(c, depth, gpm):  # :115-129
    v = nearby_vectors(depth)
    if depth == 2 and _add(c, v[3]) not in gpm[depth] and _add(c, (1, 1)) in gpm[depth]:
        return _add(c, (1, 1))
    return _add(c, v[3])


def get_down_right(c, depth, gpm):  # :131-145
    v = nearby_vectors(depth)
    if depth == 2 and _add(c, v[3]) not in gpm[depth] and _add(c, (1, 1)) in gpm[depth]:
        return _add(_add(c, (1, 1)), v[1])
    return _add(c, v[2])


def get_up_right(c, depth, gpm):  # :147-161
    v = nearby_vectors(depth)
    if depth == 2 and _add(c, v[0]) not in gpm[depth] and _add(c, (-1, -1)) in gpm[depth]:
        return _add(c, (-1, -1))
    return _add(c, v[0])


def get_up_left(c, depth, gpm):  # :163-177
    v = nearby_vectors(depth)
    if depth == 2 and _add(c, v[0]) not in gpm[depth] and _add(c, (-1, -1)) in gpm[depth]:
        return _add(_add(c, (-1, -1)), v[4])
    return _add(c, v[5])


def build_maps(centers):
    """Per level: node position -> (tile index, heap index): global_position_map[level] composed with the owning
    fractal's position_map[level] (wavelet_transform.rs:434-448, :49)."""
    maps = [dict() for _ in range(BASE_FRAC_DEPTH)]
    for t, c in enumerate(centers):
        pos = image_positions(c)
        for level in range(BASE_FRAC_DEPTH):
            for p in range(1 << level, 1 << (level + 1)):
                maps[level][pos[p]] = (t, p)
    return maps


def assign_bucket(width) -> int:  # prediction.rs:55-68 (`width as u32`: saturating, NaN -> 0)
    w = float(width)
    u = 0 if not (w > 0.0) else min(int(w), 0xFFFFFFFF)
    for b, hi in enumerate((3, 5, 6, 8, 12, 16, 20, 25, 30)):
        if u < hi:
            return b
    return 9


def _as_i32(x) -> int:  # Rust `f32 as i32`
    x = float(x)
    if x != x:
        return 0
    return max(-(1 << 31), min((1 << 31) - 1, int(x)))


def neighbour_values(pos, level, coefs, maps, ch):  # context_modeling.rs:25-77
    d = BASE_FRAC_DEPTH - level
    out = []
    for getter in (get_left, get_up_left, get_up_right):
        hit = maps[level].get(getter(pos, d, maps))
        out.append(int(coefs[hit[0], ch, hit[1]]) if hit else 0)
    for getter in (get_right, get_down_left, get_down_right):
        hit = maps[level].get(getter(pos, d, maps))
        out.append(int(coefs[hit[0], ch, hit[1] // 2]) if hit else 0)
    return out


def hf_context_bucket(pos, level, coefs, maps, value_params, width_params, ch):  # prediction.rs:151-207
    layer = 2 if level < BASE_FRAC_DEPTH - 2 else (1 if level == BASE_FRAC_DEPTH - 2 else 0)
    vp = [F(x) for x in value_params[layer]]
    wp = [F(x) for x in width_params[layer]]
    v = neighbour_values(pos, level, coefs, maps, ch)
    with np.errstate(all="ignore"):
        width = wp[0]
        for k, (a, b) in enumerate(((0, 3), (1, 2), (4, 5), (1, 5), (2, 4))):
            width = F(width + F(wp[k + 1] * F(abs(v[a] - v[b]))))
        pred = F(F(v[0]) * vp[0])
        for k in range(1, 6):
            pred = F(pred + F(F(v[k]) * vp[k]))
    return assign_bucket(width), _as_i32(pred)


def lf_context_bucket(position, tile, centers, coefs, tile_of, ch):  # prediction.rs:86-149
    c = (int(centers[tile][0]), int(centers[tile][1]))
    v9 = nearby_vectors(BASE_FRAC_DEPTH)
    vals = []
    for vec in (v9[4], v9[5], v9[0]):  # get_left, get_up_left, get_up_right with an empty global map
        t = tile_of.get(_add(c, vec))  # get_containing_fractal: one of the six lattice neighbours, if retained
        vals.append(int(coefs[t, ch, position]) if t is not None else 0)
    bucket = assign_bucket(F(abs(vals[0] - vals[2])))
    hi, lo = max(vals[0], vals[2]), min(vals[0], vals[2])
    if vals[1] >= hi:
        pred = hi
    elif vals[1] <= lo:
        pred = lo
    else:
        pred = vals[0] + vals[2] - vals[1]
    return bucket, pred


def pack_signed(k: int) -> int:  # utils.rs:34-40
    return 2 * k if k >= 0 else -2 * k - 1


def predict(centers, coefs, some, emit_src, value_params, width_params):
    """For every channel: (bucket u8, prediction i32, symbol u32) of every `Some` coefficient in the order
    `emit_src` lists them (tile * 512 + heap index: the entropy coder's order), and the per-context histograms
    [C][10][1024] (prediction.rs:224-323).  coefs: quantized [n_tiles, C, 512], None slots 0.
    value_params / width_params: [C][3][6] f32.  A symbol outside the alphabet (the reference would panic at
    entropy_coding.rs:99) is recorded in `overflow`."""
    n_tiles, c, _ = coefs.shape
    maps = build_maps(centers)
    tile_of = {(int(x), int(y)): t for t, (x, y) in enumerate(centers)}
    rel = image_positions((0, 0))
    n = len(emit_src)
    bucket = np.zeros((c, n), np.uint8)
    pred = np.zeros((c, n), np.int32)
    sym = np.zeros((c, n), np.uint32)
    hist = np.zeros((c, CONTEXT_AMOUNT, ALPHABET_SIZE), np.uint32)
    overflow = 0
    for ch in range(c):
        for k, src in enumerate(emit_src):
            tile, heap = int(src) >> 9, int(src) & 511
            assert some[tile, heap]
            if heap < 2:
                b, p = lf_context_bucket(heap, tile, centers, coefs, tile_of, ch)
            else:
                level = heap.bit_length() - 1
                pos = (int(centers[tile][0]) + rel[heap][0], int(centers[tile][1]) + rel[heap][1])
                b, p = hf_context_bucket(pos, level, coefs, maps, value_params[ch], width_params[ch], ch)
            # residual in wrapping i32 like release-mode Rust, then pack_signed -> u32
            r = (int(coefs[tile, ch, heap]) - p + (1 << 31)) % (1 << 32) - (1 << 31)
            s = pack_signed(r) & 0xFFFFFFFF
            bucket[ch, k], pred[ch, k], sym[ch, k] = b, p, s
            if s < ALPHABET_SIZE:
                hist[ch, b, s] += 1
            else:
                overflow += 1
    return bucket, pred, sym, hist, overflow
