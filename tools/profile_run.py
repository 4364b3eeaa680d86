#!/usr/bin/env python
"""tools/profile_run.py -- a short, fixed workload for ncu: G frames (one launch group) of the bench
workload pushed through the device path R times.  Prints per-stage CUDA-event times.

    python tools/profile_run.py [--frames 256] [--reps 3]
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
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--width", type=int, default=640)
ap.add_argument("--height", type=int, default=480)
args = ap.parse_args()

F = args.frames
frames = bench.make_frames(F, seed=0).cuda()
ext = ex.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=F, profile=True)
cap = ext.max_keypoints(bench.W, bench.H)
kps = torch.empty((F, cap, 7), dtype=torch.float32, device="cuda")
desc = torch.empty((F, cap, 32), dtype=torch.uint8, device="cuda")
counts = torch.zeros((F, 2), dtype=torch.int32, device="cuda")
s = torch.cuda.Stream()
for r in range(args.reps):
    ext.extract_batch_raw(frames.data_ptr(), ex.MEM_DEVICE, F, bench.W, bench.H, bench.W, bench.W * bench.H, (0, 0),
                          kps.data_ptr(), desc.data_ptr(), cap, counts.data_ptr(), ex.MEM_DEVICE, s.cuda_stream)
    ms, n = ext.stage_times()
    tot = sum(ms.values())
    print("rep %d: %d launches, total %.3f ms (%.2f us/frame) " % (r, n, tot, 1e3 * tot / F) +
          " ".join("%s=%.3f" % kv for kv in ms.items()))
print("mean keypoints", float(counts[:, 0].float().mean()))
