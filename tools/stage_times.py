#!/usr/bin/env python
"""tools/stage_times.py -- per-stage CUDA-event times of one launch group for any configuration.

    python tools/stage_times.py --width 3840 --height 2160 --nfeatures 8000 --nlevels 12 --frames 32
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
import extractorb_b200 as ex  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=640)
ap.add_argument("--height", type=int, default=480)
ap.add_argument("--nfeatures", type=int, default=1000)
ap.add_argument("--nlevels", type=int, default=8)
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--bases", type=int, default=4)
a = ap.parse_args()
F = a.frames
frames = bench.make_frames(F, seed=0, w=a.width, h=a.height, n_base=a.bases).cuda()
ext = ex.ORBextractor(a.nfeatures, 1.2, a.nlevels, 20, 7, max_batch=F, profile=True)
cap = ext.max_keypoints(a.width, a.height)
kps = torch.empty((F, cap, 7), dtype=torch.float32, device="cuda")
desc = torch.empty((F, cap, 32), dtype=torch.uint8, device="cuda")
counts = torch.zeros((F, 2), dtype=torch.int32, device="cuda")
s = torch.cuda.Stream()
for r in range(a.reps):
    ext.extract_batch_raw(frames.data_ptr(), ex.MEM_DEVICE, F, a.width, a.height, a.width, a.width * a.height, (0, 0), kps.data_ptr(),
                          desc.data_ptr(), cap, counts.data_ptr(), ex.MEM_DEVICE, s.cuda_stream)
    ms, n = ext.stage_times()
    tot = sum(ms.values())
    print("rep %d: %d launches, total %.3f ms (%.2f us/frame) " % (r, n, tot, 1e3 * tot / F) + " ".join("%s=%.3f" % kv for kv in ms.items()))
print("mean keypoints", float(counts[:, 0].float().mean()), "candidates frame 0:", [len(ext.level_candidates(l)[0]) for l in range(a.nlevels)])
