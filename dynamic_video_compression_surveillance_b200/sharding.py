"""Multi-GPU partitioning of the hot loop.

The loop shards by independent units: every camera stream owns its state (prev_gray, accumulated_mask /
mask window; frame_differencing.py:75-81), and the reference itself just loops over files
(windows.py:144).  One process per GPU, no collective on the per-frame path; a single all-reduce of the
statistics counters at the end (NCCL on the GPU box, gloo in the CPU tests).

A single long stream can instead be cut into contiguous frame chunks.  The window-vote loop needs the last
K raw masks and one gray frame, so a warm-up halo of K frames re-creates the state exactly; the EMA loop
(cv2.addWeighted) re-converges exactly after 14 frames for release_factor 0.5 (6 for 0.3) and otherwise hands
the state plane over (FramePipeline.get_state / set_state).
"""
from __future__ import annotations

from dataclasses import dataclass

COUNTER_FIELDS = ("frames", "pixels", "motion_pixels", "blocks", "static_blocks")


def shard_streams(n_streams: int, world_size: int, rank: int) -> list[int]:
    """Stream ids owned by ``rank``: contiguous blocks, sizes differing by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank outside world")
    base, extra = divmod(n_streams, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def shard_streams_weighted(n_streams: int, weights: list) -> list[list[int]]:
    """Contiguous stream ids per rank, counts proportional to ``weights`` (largest-remainder apportionment; ties go to the
    lower rank).  Used for the end-to-end path, where a rank's sustainable rate is its host link's share of the box's IO
    fabric, not its GPU: on the 8-GPU boxes of this pool four GPUs sit behind an uplink with 1.5x the bandwidth of the
    other four, so equal shards leave the faster links idle a third of the time."""
    w = [max(0.0, float(x)) for x in weights]
    if not w or sum(w) <= 0:
        raise ValueError("weights must contain a positive value")
    total = sum(w)
    quota = [n_streams * x / total for x in w]
    counts = [int(q) for q in quota]
    order = sorted(range(len(w)), key=lambda r: (-(quota[r] - counts[r]), r))
    for r in order[:n_streams - sum(counts)]:
        counts[r] += 1
    out, start = [], 0
    for c in counts:
        out.append(list(range(start, start + c)))
        start += c
    return out


@dataclass(frozen=True)
class FrameChunk:
    rank: int
    warm_start: int    # first frame fed to the loop (its outputs are discarded up to ``start``)
    start: int         # first frame whose outputs this rank keeps
    stop: int          # one past the last frame this rank keeps


def ema_exact_halo(release_factor: float) -> int | None:
    """Frames after which the uint8 EMA state no longer depends on its starting value (None: never exact)."""
    if release_factor == 0.5:
        return 14
    if release_factor <= 0.3:
        return 6
    return None


def frame_chunks(n_frames: int, world_size: int, halo: int) -> list[FrameChunk]:
    """Cut frames [1, n_frames) (frame 0 only seeds prev_gray) into ``world_size`` contiguous chunks; every chunk
    but the first re-processes ``halo`` earlier frames to rebuild its temporal state.  The frame before
    ``warm_start`` seeds prev_gray of that chunk."""
    n = max(0, n_frames - 1)
    base, extra = divmod(n, world_size)
    out, start = [], 1
    for r in range(world_size):
        size = base + (1 if r < extra else 0)
        warm = max(1, start - halo)
        out.append(FrameChunk(r, warm, start, start + size))
        start += size
    return out


def reduce_counters(counters: dict, device=None, group=None) -> dict:
    """Sum the statistics over all ranks: the only collective of the path."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(counters)
    t = torch.tensor([int(counters[k]) for k in COUNTER_FIELDS], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(COUNTER_FIELDS, (int(v) for v in t.tolist())))


def motion_percentage(counters: dict) -> float:
    return 100.0 * counters["motion_pixels"] / counters["pixels"] if counters["pixels"] else 0.0


def static_block_percentage(counters: dict) -> float:
    return 100.0 * counters["static_blocks"] / counters["blocks"] if counters["blocks"] else 0.0
