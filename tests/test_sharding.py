"""N>1 host logic on CPU: world_size-2 gloo process group, frames sharded round-robin, statistics reduced,
results independent of the number of ranks.  The per-frame work is done by the ORACLE here (this is a test
of the sharding logic, not of the CUDA path; the GPU multi-rank run is bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest

from common import synth_frame, kp_bytes_equal

N_FRAMES = 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist
    from extractorb_b200 import sharding
    from oracle import pyoracle
    from common import synth_frame as sf
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    o = pyoracle.OracleExtractor(300, 1.2, 4, 20, 7)
    local = sharding.run_shard(lambda img: o.extract(img, (0, 0)), lambda i: sf(900 + i, 320, 240), N_FRAMES, rank, world)
    stats = sharding.reduce_stats(len(local), sum(len(v[1]) for v in local.values()), 10.0 * (rank + 1), [1.0, 2.0])
    merged = sharding.gather_results(local, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), np.array([stats["frames"], stats["keypoints"], stats["elapsed_ms_max"],
                                                              stats["world"]] + stats["stage_ms_sum"]))
        for k, (ret, kps, desc) in merged.items():
            np.savez(os.path.join(out_dir, "f%d.npz" % k), ret=ret, kps=kps, desc=desc)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_covers_every_frame_once():
    from extractorb_b200 import sharding
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 4096):
            seen = sorted(i for r in range(world) for i in sharding.frames_for_rank(n, r, world))
            assert seen == list(range(n))
            assert all(sharding.owner_of(i, world) == r for r in range(world) for i in sharding.frames_for_rank(n, r, world))
    with pytest.raises(ValueError):
        sharding.frames_for_rank(4, 2, 2)


def test_two_ranks_gloo_equal_single_process(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    stats = np.load(tmp_path / "stats.npy")
    o = oracle.OracleExtractor(300, 1.2, 4, 20, 7)
    total_kp = 0
    for i in range(N_FRAMES):
        ret, kps, desc = o.extract(synth_frame(900 + i, 320, 240), (0, 0))
        with np.load(tmp_path / ("f%d.npz" % i)) as z:
            assert int(z["ret"]) == ret and kp_bytes_equal(z["kps"], kps) and np.array_equal(z["desc"], desc)
        total_kp += len(kps)
    assert stats[0] == N_FRAMES and stats[1] == total_kp
    assert stats[2] == 20.0 and stats[3] == 2          # MAX over ranks of the elapsed time; world size
    assert stats[4] == 2.0 and stats[5] == 4.0         # stage sums over ranks


def test_reduce_stats_without_process_group():
    from extractorb_b200 import sharding
    s = sharding.reduce_stats(3, 3000, 12.5, [1.0])
    assert s == {"frames": 3, "keypoints": 3000, "stage_ms_sum": [1.0], "elapsed_ms_max": 12.5, "world": 1}
