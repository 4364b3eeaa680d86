"""extractorb_b200/sharding.py -- multi-GPU host logic: frames are independent, so they are sharded across
ranks (one process per GPU) and NO data-path collective exists.  torch.distributed (NCCL on GPUs, gloo in the
CPU tests) is used only to reduce timing and statistics, as the reference's only concurrency is two host
threads for stereo left/right (reference src/Frame.cc:109-112) and nothing crosses frames.

Partitioning: frame i -> rank i mod G.  Results must be byte-identical regardless of G.
"""
from typing import Callable, Dict, Iterable, List, Tuple

import numpy as np


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin shard: the global frame indices owned by `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def owner_of(frame: int, world: int) -> int:
    return frame % world


def stereo_pairs_for_rank(n_pairs: int, rank: int, world: int) -> List[int]:
    """A stereo pair stays on one GPU (left/right run on two streams/handles of that GPU)."""
    return frames_for_rank(n_pairs, rank, world)


def run_shard(extract_fn: Callable[[np.ndarray], Tuple[int, np.ndarray, np.ndarray]], get_frame: Callable[[int], np.ndarray],
              n_frames: int, rank: int, world: int) -> Dict[int, Tuple[int, np.ndarray, np.ndarray]]:
    """Apply `extract_fn` (e.g. an ORBextractor bound to this rank's GPU) to this rank's frames."""
    return {i: extract_fn(get_frame(i)) for i in frames_for_rank(n_frames, rank, world)}


def reduce_stats(frames: int, keypoints: int, elapsed_ms: float, stage_ms: Iterable[float] = (), device=None, group=None):
    """Whole-job statistics: SUM of frames/keypoints/stage times, MAX of the elapsed time over ranks.
    Falls back to the local values when torch.distributed is not initialised (single process)."""
    import torch
    import torch.distributed as dist
    stage = list(stage_ms)
    sums = torch.tensor([float(frames), float(keypoints)] + [float(s) for s in stage], dtype=torch.float64, device=device)
    mx = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    world = 1
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        world = dist.get_world_size(group)
    s = sums.tolist()
    return {"frames": int(round(s[0])), "keypoints": int(round(s[1])), "stage_ms_sum": s[2:], "elapsed_ms_max": float(mx.item()),
            "world": world}


def gather_results(local: Dict[int, tuple], dst: int = 0, group=None):
    """Collect every rank's {frame: result} on `dst` (tests / small jobs only: results travel as objects)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    world = dist.get_world_size(group)
    out = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(local, out, dst=dst, group=group)
    if out is None:
        return None
    merged = {}
    for part in out:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError("frame %d extracted twice" % k)
            merged[k] = v
    return merged
