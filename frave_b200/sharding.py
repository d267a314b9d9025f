"""Frame sharding across the GPUs of one node.

Fractals, channels and frames are independent (wavelet_transform.rs:412-414, :191), so a batch
is split by frame with no exchange step: rank r of `world` takes one contiguous block.  There is
no collective on the data path; torch.distributed is used by the callers only for the timing
barrier and the max-over-ranks reduction.
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
