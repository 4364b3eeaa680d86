#!/usr/bin/env python
"""tools/bench_clahe.py -- measurement for the pre-processing row (cv::createCLAHE(3.0, Size(8, 8))->apply, reference
src/orb_extractor/main_orb_extractor.cpp:19-22) on the bench workload's frames: 640x480, device-resident batch (CUDA
events on the launching stream, 3 warm-up + 10 timed passes), single host frame latency, and the CPU oracle / python-cv2
time for the same frame.  Algorithmic bytes per frame = read W*H + write W*H; HBM roofline against MEASURED_PEAKS.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import extractorb_b200 as ex  # noqa: E402
from oracle import pyoracle  # noqa: E402

W, H, F = 640, 480, 2048
peak = 6550.4
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
host = bench.make_frames(F, seed=0)
dev = host.cuda()
out = torch.empty_like(dev)
e = ex.ORBextractor(1000, 1.2, 8, 20, 7)
s = torch.cuda.Stream()
def step():
    ex.clahe_raw(e, dev.data_ptr(), ex.MEM_DEVICE, F, W, H, W, W * H, 3.0, (8, 8), out.data_ptr(), ex.MEM_DEVICE, W, W * H, s.cuda_stream)
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(10):
    step()
e1.record(s)
torch.cuda.synchronize()
fps = F * 10 / (e0.elapsed_time(e1) * 1e-3)
f0 = host[0].numpy()
assert np.array_equal(out[0].cpu().numpy(), pyoracle.clahe(f0, 3.0, (8, 8)))
ts = []
for i in range(250):
    t0 = time.perf_counter()
    r = ex.clahe(e, f0, 3.0, (8, 8))
    ts.append((time.perf_counter() - t0) * 1e6)
c = []
for _ in range(20):
    t0 = time.perf_counter()
    pyoracle.clahe(f0, 3.0, (8, 8))
    c.append((time.perf_counter() - t0) * 1e6)
res = {"workload": "640x480, clipLimit 3.0, 8x8 tiles, %d device-resident frames per pass" % F, "frames_per_s": fps, "us_per_frame": 1e6 / fps,
       "algorithmic_bytes_per_frame": 2 * W * H, "roofline": {"bound": "hbm", "achieved": 2 * W * H * fps / 1e9, "peak": peak, "unit": "GB/s",
                                                                "frac": 2 * W * H * fps / 1e9 / peak},
       "single_host_frame_p50_us": float(np.median(ts[50:])), "cpu_oracle_p50_us": float(np.median(c)), "cpu": "oracle C restatement, 1 thread"}
try:
    import cv2
    cl = cv2.createCLAHE(3.0, (8, 8))
    c2 = []
    for _ in range(50):
        t0 = time.perf_counter()
        cl.apply(f0)
        c2.append((time.perf_counter() - t0) * 1e6)
    res["python_cv2_p50_us"] = float(np.median(c2))
except Exception:
    pass
print(json.dumps(res))
