"""numpy restatement of frave's fractal transform + quantization hot path.

TEST INFRASTRUCTURE ONLY — never imported by the product (frave_b200/).  PARITY UNPINNED:
the Rust reference cannot be built here and ships no golden vectors; this module restates
the cited reference lines (paths relative to /root/reference/) and exists as a SECOND,
independently written restatement so that a mis-reading in oracle/fri_oracle.c shows up as a
disagreement (tests/test_oracle.py runs both against each other and against the known-answer
hashes of SURVEY.md §8(c)).

Style is deliberately different from the C file: everything is vectorised over tiles, and
Option<i32> is an (int64 value, bool some) pair of arrays.
"""
from __future__ import annotations

import hashlib
import struct
from collections import deque

import numpy as np

# crates/libfri/src/fractal.rs:51-86  (re = x, im = y)
LITERALS = np.array(
    [
        (0, 1), (-1, 1), (2, 0), (-3, -1), (5, -1), (1, 3), (-11, -1), (9, -5), (13, 7),
        (-31, 3), (5, -17), (57, 11), (-67, 23), (-47, -45), (181, -1), (-87, 91),
        (-275, -89), (449, -93), (101, 271), (-999, -85), (797, -457), (1201, 627),
        (-2795, 287), (393, -1541), (5197, 967), (-5983, 2115), (-4411, -4049),
        (16377, -181), (-7555, 8279), (-25199, -7917),
    ],
    dtype=np.int64,
)


def nearby_vectors(depth: int) -> np.ndarray:
    """wavelet_transform.rs:71-90"""
    if depth == 1:
        zl, zmd = np.array([-1, 1]), np.array([0, 2])
    elif depth == 2:
        zl, zmd = np.array([-2, 0]), np.array([0, -2])
    elif depth == 3:
        zl, zmd = np.array([-3, -1]), np.array([-1, -3])
    else:
        zl = LITERALS[depth]
        zmd = LITERALS[depth + 1] + zl
    return np.stack([zl, zl - zmd, -zmd, -zl, zmd - zl, zmd]).astype(np.int64)


def leaf_offsets(depth: int) -> np.ndarray:
    """Offsets of the 2^depth leaves (heap index 2^depth + k) from the tile centre.

    wavelet_transform.rs:47-53: going to the right child at tree level `level` adds
    LITERALS[depth-level-1]; bit j of k (LSB = deepest level) therefore selects LITERALS[j].
    """
    k = np.arange(1 << depth)
    off = np.zeros((1 << depth, 2), dtype=np.int64)
    for j in range(depth):
        off += ((k >> j) & 1)[:, None] * LITERALS[j][None, :]
    return off


def fractal_divide(width: int, height: int, depth: int) -> list[tuple[int, int]]:
    """wavelet_transform.rs:450-484 — literal queue/boundary simulation (set of built tiles)."""
    vec = [tuple(int(c) for c in v) for v in nearby_vectors(depth)]
    lattice: dict[tuple[int, int], None] = {}
    to_add = deque([(width // 2, height // 2)])
    in_queue = {to_add[0]: 1}
    boundary = []
    while to_add:
        pos = to_add.popleft()
        in_queue[pos] -= 1
        if pos[0] < 0 or pos[1] < 0 or pos[0] > width or pos[1] > height:
            boundary.append(pos)
            continue
        for v in vec:
            nb = (pos[0] + v[0], pos[1] + v[1])
            if nb not in lattice and in_queue.get(nb, 0) == 0:
                to_add.append(nb)
                in_queue[nb] = in_queue.get(nb, 0) + 1
        lattice[pos] = None
    for pos in boundary:
        lattice[pos] = None
    return list(lattice.keys())


def _trunc_div(a: np.ndarray, b) -> np.ndarray:
    """Rust i32 `/`: truncation toward zero."""
    a = np.asarray(a, dtype=np.int64)
    q = np.abs(a) // np.abs(b)
    return np.where((a < 0) != (np.asarray(b) < 0), -q, q)


def forward_tiles(img: np.ndarray, centers: np.ndarray, depth: int):
    """wavelet_transform.rs:179-225 for many tiles at once.

    img: (H, W, C) integer array.  centers: (n, 2) of (re, im).
    Returns coef (n, C, 2^depth) int64 and some (n, C, 2^depth) bool.
    """
    h, w, c = img.shape
    n = len(centers)
    size = 1 << depth
    off = leaf_offsets(depth)
    x = centers[:, 0:1] + off[None, :, 0]
    y = centers[:, 1:2] + off[None, :, 1]
    inside = (x >= 0) & (y >= 0) & (x < w) & (y < h)  # images.rs:90
    xc, yc = np.clip(x, 0, w - 1), np.clip(y, 0, h - 1)
    val = img[yc, xc, :].astype(np.int64)  # (n, size, C)
    val = np.where(inside[:, :, None], val, 0).transpose(0, 2, 1)  # (n, C, size)
    some = np.broadcast_to(inside[:, None, :], (n, c, size)).copy()

    coef = np.zeros((n, c, size), dtype=np.int64)
    csome = np.zeros((n, c, size), dtype=bool)
    low, lsome = val, some
    for level in range(depth - 1, -1, -1):
        l, r = low[:, :, 0::2], low[:, :, 1::2]
        ls, rs = lsome[:, :, 0::2], lsome[:, :, 1::2]
        # try_apply(left, right, l - r, 0)  (:211-212)
        d = np.where(ls, l, 0) - np.where(rs, r, 0)
        ds = ls | rs
        # try_apply(right, d, l + r/2, 0)   (:213-218)
        s = np.where(rs, r, 0) + _trunc_div(np.where(ds, d, 0), 2)
        ss = rs | ds
        lo = 1 << level
        coef[:, :, lo:2 * lo] = np.where(ds, d, 0)
        csome[:, :, lo:2 * lo] = ds
        low, lsome = np.where(ss, s, 0), ss
    coef[:, :, 0] = low[:, :, 0]  # :221
    csome[:, :, 0] = lsome[:, :, 0]
    return coef, csome


def from_raster(img: np.ndarray, depth: int = 9):
    """wavelet_transform.rs:405-416; tiles sorted by (im, re); retain over active channels
    (the reference's 3-slot predicate drops every tile of a 1-channel image — see the note in
    oracle/fri_oracle.c)."""
    h, w, _ = img.shape
    built = np.array(sorted(fractal_divide(w, h, depth), key=lambda p: (p[1], p[0])), dtype=np.int64)
    coef, some = forward_tiles(img, built, depth)
    keep = some[:, :, 0].all(axis=1)
    return built[keep], coef[keep], some[keep]


def quant_layers(depth: int) -> np.ndarray:
    """quantization.rs:13 — layer(i) = trailing_zeros(prev_power_two(i+1)) = floor(log2(i+1))."""
    i = np.arange(1 << depth) + 1
    return np.floor(np.log2(i)).astype(np.int64)


def quantize(coef: np.ndarray, some: np.ndarray, q, depth: int = 9, multiply: bool = False) -> np.ndarray:
    """quantization.rs:7-25 / :27-45 (both divide); multiply=True is the non-reference variant."""
    qv = np.asarray(q, dtype=np.int64)[quant_layers(depth)]
    out = coef * qv if multiply else _trunc_div(coef, qv)
    return np.where(some, out, coef)


def inverse_tiles(centers: np.ndarray, coef: np.ndarray, some: np.ndarray, depth: int, h: int, w: int,
                  maxval: int = 255) -> np.ndarray:
    """wavelet_transform.rs:358-381 + images.rs:103-111 into a zeroed (H, W, C) raster."""
    n, c, size = coef.shape
    low = coef[:, :, 0:1].astype(np.int64)  # low_pass_values[1] = coef[0].unwrap()
    alive = np.ones((n, c, 1), dtype=bool)  # has every ancestor been Some? (else stays 0 / unwritten)
    for level in range(depth):
        lo = 1 << level
        d = coef[:, :, lo:2 * lo]
        ok = some[:, :, lo:2 * lo]
        right = low - _trunc_div(d, 2)
        left = d + right
        nxt = np.zeros((n, c, 2 * lo), dtype=np.int64)
        nxt[:, :, 0::2] = np.where(ok, left, 0)
        nxt[:, :, 1::2] = np.where(ok, right, 0)
        al = np.zeros((n, c, 2 * lo), dtype=bool)
        al[:, :, 0::2] = ok
        al[:, :, 1::2] = ok
        # a skipped node leaves its children at 0, and they are still processed if Some:
        low, alive = nxt, al
    off = leaf_offsets(depth)
    x = centers[:, 0:1] + off[None, :, 0]
    y = centers[:, 1:2] + off[None, :, 1]
    inside = (x >= 0) & (y >= 0) & (x < w) & (y < h)
    out = np.zeros((h, w, c), dtype=np.int64)
    for ch in range(c):
        wr = inside & alive[:, ch, :]
        out[y[wr], x[wr], ch] = np.clip(low[:, ch, :][wr], 0, maxval)
    return out


def kat_hash(centers: np.ndarray, coef: np.ndarray, some: np.ndarray):
    """SURVEY.md §8(c) digest: tiles sorted by (im, re); <i32 re><i32 im>, then per channel per
    coefficient Some(v) -> <i32 v LE> 01, None -> FF FF FF 7F 00."""
    order = np.lexsort((centers[:, 0], centers[:, 1]))
    hsh = hashlib.sha256()
    for t in order:
        hsh.update(struct.pack("<ii", int(centers[t, 0]), int(centers[t, 1])))
        v = coef[t].astype("<i4")
        s = some[t].astype(bool)
        rec = np.zeros(v.shape + (5,), dtype=np.uint8)
        rec[..., :4] = v.view(np.uint8).reshape(v.shape + (4,))
        rec[..., 4] = 1
        rec[~s] = np.frombuffer(b"\xff\xff\xff\x7f\x00", dtype=np.uint8)
        hsh.update(rec.tobytes())
    return hsh.hexdigest(), int(some.sum()), int(coef[some].sum())


def survey_image(w: int, h: int, c: int) -> np.ndarray:
    """pix(x, y, ch) = (7x + 13y + (x*y mod 11) + 29ch) mod 256."""
    y, x, ch = np.meshgrid(np.arange(h), np.arange(w), np.arange(c), indexing="ij")
    return ((7 * x + 13 * y + (x * y) % 11 + 29 * ch) % 256).astype(np.uint8)
