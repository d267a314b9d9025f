"""Emission order of the coefficients — CPU restatement.

TEST INFRASTRUCTURE ONLY, PARITY UNPINNED (see oracle/fri_oracle.h): a literal restatement of
the reference's host-side ordering logic with Python dicts, written to be compared with the
plan's arithmetic implementation (frave_b200/csrc/fri_order.cpp).  Paths are relative to
/root/reference/.

  crates/libfri/src/stages/wavelet_transform.rs:434-448   get_global_position_map
  crates/libfri/src/stages/wavelet_transform.rs:490-503   is_pos_in_row_boundary
  crates/libfri/src/stages/wavelet_transform.rs:505-654   scan_level
  crates/libfri/src/stages/wavelet_transform.rs:657-705   sort_lattice
  crates/libfri/src/stages/entropy_coding.rs:283-329      the three scans that consume the order

The only thing the reference's own code pins here is the assertion at :701 — every level's scan
must visit every node of every retained fractal exactly once; `sort_lattice` below raises the
same way.
"""
from __future__ import annotations

import numpy as np

from .fri_oracle_np import LITERALS, nearby_vectors

BASE_FRAC_DEPTH = 9  # wavelet_transform.rs:39


def image_positions(center, depth: int = BASE_FRAC_DEPTH):
    """wavelet_transform.rs:42-54: positions[1] = centre; [2p] = [p]; [2p+1] = [p] + LITERALS[depth-level-1]."""
    pos = [None] * (1 << (depth + 1))
    pos[0] = pos[1] = (int(center[0]), int(center[1]))
    for level in range(depth):
        lx, ly = int(LITERALS[depth - level - 1][0]), int(LITERALS[depth - level - 1][1])
        for p in range(1 << level, 1 << (level + 1)):
            pos[2 * p] = pos[p]
            pos[2 * p + 1] = (pos[p][0] + lx, pos[p][1] + ly)
    return pos


def global_position_map(centers):
    """:434-448 — per level: node position -> centre of the fractal that owns it.  Also returns the
    per-fractal position_map (position -> heap index, :49) the entropy coder looks values up with."""
    gpm = [dict() for _ in range(BASE_FRAC_DEPTH)]
    heap = [dict() for _ in range(BASE_FRAC_DEPTH)]
    for c in centers:
        c = (int(c[0]), int(c[1]))
        pos = image_positions(c)
        for level in range(BASE_FRAC_DEPTH):
            for p in range(1 << level, 1 << (level + 1)):
                gpm[level][pos[p]] = c
                heap[level][(c, pos[p])] = p
    return gpm, heap


def _in_row_boundary(pos, row_dir, min_real, max_real, min_imag, max_imag):  # :490-503
    if abs(row_dir[0]) > abs(row_dir[1]):
        return min_imag <= pos[1] <= max_imag
    return min_real <= pos[0] <= max_real


def _add(a, b):
    return (a[0] + b[0], a[1] + b[1])


def scan_level(level, depth, center, gpm, min_real, max_real, min_imag, max_imag):
    """:505-654, statement for statement."""
    vec = [tuple(int(c) for c in v) for v in nearby_vectors(BASE_FRAC_DEPTH - level)]
    row_dir, rev_row_dir = vec[3], vec[0]
    col_dir, rev_col_dir = vec[1], vec[4]

    def in_box(p):
        return min_imag <= p[1] <= max_imag and min_real <= p[0] <= max_real

    first = center
    layer_seven_mod = 0
    if _add(center, rev_row_dir) not in gpm and _add(center, (-1, -1)) in gpm:  # :523-527
        layer_seven_mod = 1
    last_seen = first

    def step_back(first, mod):  # :532-541, :573-582
        if depth - level != 2:
            return _add(first, rev_row_dir), mod
        nxt = _add(first, rev_row_dir) if mod % 2 == 0 else _add(first, (-1, -1))
        return nxt, mod + 1

    while first in gpm:  # :530-542
        last_seen = first
        first, layer_seven_mod = step_back(first, layer_seven_mod)

    while True:  # :545-584 find first row
        fwd = bwd = first
        empty = True
        while (min_imag <= fwd[1] <= max_imag) or (min_imag <= bwd[1] <= max_imag) or \
                (min_real <= fwd[0] <= max_real) or (min_real <= bwd[0] <= max_real):
            fwd = _add(fwd, col_dir)
            bwd = _add(bwd, rev_col_dir)
            if fwd in gpm:
                last_seen, empty = fwd, False
                break
            if bwd in gpm:
                last_seen, empty = bwd, False
                break
        if empty:
            first = last_seen
            break
        first, layer_seven_mod = step_back(first, layer_seven_mod)

    while in_box(first):  # :587-596 scanning backwards find first column
        first = _add(first, rev_col_dir)
        if first in gpm:
            last_seen = first
    first = last_seen
    layer_seven_mod = 1

    plane = []
    while True:  # :601-652 fill plane in sorted order
        scan = first
        while True:
            if scan in gpm:
                plane.append(scan)
            if (scan[1] > max_imag or scan[1] < min_imag) or \
                    (col_dir[1] == 0 and (scan[0] > max_real or scan[0] < min_real)):
                break
            scan = _add(scan, col_dir)
        if depth - level != 2:
            first = _add(first, row_dir)
        else:
            first = _add(first, (1, 1)) if layer_seven_mod % 2 == 0 else _add(first, row_dir)
            layer_seven_mod += 1
        done = False
        while first not in gpm:
            first = _add(first, col_dir)
            if not _in_row_boundary(first, row_dir, min_real, max_real, min_imag, max_imag):
                done = True
                break
        if done:
            break
        last_seen = first
        while in_box(first):
            first = _add(first, rev_col_dir)
            if first in gpm:
                last_seen = first
        first = last_seen
    return plane


def sort_lattice(centers, width, height):
    """:657-705 — [level] -> list of node positions in emission order (+ the maps)."""
    gpm, heap = global_position_map(centers)
    keys = list(gpm[BASE_FRAC_DEPTH - 1].keys())
    min_real, max_real = min(k[0] for k in keys), max(k[0] for k in keys)
    min_imag, max_imag = min(k[1] for k in keys), max(k[1] for k in keys)
    center = (width // 2, height // 2)
    planes = []
    for level in range(BASE_FRAC_DEPTH):
        plane = scan_level(level, BASE_FRAC_DEPTH, center, gpm[level], min_real, max_real, min_imag, max_imag)
        if len(plane) != len(centers) * (1 << level):  # the reference's assert_eq! at :701
            raise AssertionError(f"level {level}: scan visited {len(plane)} nodes, expected {len(centers) * (1 << level)}")
        if len(set(plane)) != len(plane):
            raise AssertionError(f"level {level}: a node was visited twice")
        planes.append(plane)
    return planes, gpm, heap


def emission_order(centers, width, height):
    """Flat source index list in the order entropy_coding.rs:283-329 consumes coefficients of one
    channel: all DCs (coefficient 0) in level-0 order, all roots (coefficient 1) in level-0 order,
    then levels 1..8.  Entries are (tile_row, heap_index) with tile_row indexing `centers`; the
    `None` filter of the reference (`if let Some(value)`) is left to the caller."""
    planes, gpm, heap = sort_lattice(centers, width, height)
    row = {(int(c[0]), int(c[1])): i for i, c in enumerate(centers)}
    out = [(row[p], 0) for p in planes[0]]  # level-0 positions are the centres (:284-286)
    out += [(row[p], 1) for p in planes[0]]
    for level in range(1, BASE_FRAC_DEPTH):
        for p in planes[level]:
            c = gpm[level][p]
            out.append((row[c], heap[level][(c, p)]))
    return np.array(out, dtype=np.int64)
