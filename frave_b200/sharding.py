"""Sharding across the GPUs of one node.

Fractals, channels and frames are independent (wavelet_transform.rs:412-414, :191), so a batch
is split by frame with no exchange step: rank r of `world` takes one contiguous block.  There is
no collective on the data path; torch.distributed is used by the callers only for the timing
barrier and the max-over-ranks reduction.

A single huge image is split by contiguous ranges of tile groups instead
(`shard_image`): rank r transforms the tiles of its groups from the band of pixel rows they touch.
Tiles straddle the cut, so the row bands of neighbouring ranks overlap by a halo; on decode every rank
writes only the pixels its own tiles own, and the overlap rows are the one exchange step of the path —
bands that start from zero merge by addition (`overlaps` lists who shares which rows with whom).
"""
from __future__ import annotations


def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Contiguous block of frame indices owned by `rank` (sizes differ by at most one)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def frame_owner(frame: int, n_frames: int, world: int) -> int:
    """Inverse of shard_frames."""
    base, extra = divmod(n_frames, world)
    cut = extra * (base + 1)
    if frame < cut:
        return frame // (base + 1)
    return extra + (frame - cut) // base


def shard_image(plan, rank: int, world: int) -> dict:
    """Part `rank` of `world` of one image (capi.Plan.part): group / tile / pixel-row ranges, half open."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return plan.part(rank, world)


def overlaps(plan, rank: int, world: int) -> list[tuple[int, int, int]]:
    """(peer, row_lo, row_hi) for every other rank whose pixel-row band intersects this rank's: the rows both
    write into on decode (each its own pixels) and both read on encode."""
    mine = shard_image(plan, rank, world)
    out = []
    for peer in range(world):
        if peer == rank:
            continue
        other = shard_image(plan, peer, world)
        lo, hi = max(mine["row_begin"], other["row_begin"]), min(mine["row_end"], other["row_end"])
        if lo < hi:
            out.append((peer, lo, hi))
    return out


def halo_pushes(plan, rank: int, world: int) -> list[dict]:
    """The exchange step of the image split done by the transform kernel's own stores: for every peer sharing pixel
    rows with `rank`, the sub-range of `rank`'s groups that touch the shared rows (`first`, `last`) and the rows that
    sub-range touches in all (`span_begin`, `span_end`) — decoding those groups a second time into the PEER's band
    (fri_decode_tq_device_groups on a peer-mapped pointer) completes the peer's rows without a merge pass.  The peer's
    band needs a margin so that `span` fits: the groups along the cut also own pixels outside the shared rows."""
    mine = shard_image(plan, rank, world)
    out = []
    for peer, lo, hi in overlaps(plan, rank, world):
        g = plan.groups_in_rows(mine["group_begin"], mine["group_end"], lo, hi)
        if g["first"] < g["last"]:
            out.append({"peer": peer, "row_lo": lo, "row_hi": hi, **g})
    return out
