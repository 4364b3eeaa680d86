#!/usr/bin/env python
"""tools/bench_configs.py -- the other BASELINE.json configurations (they are parity-test cases; bench.py's line is
configs[1]): C1 luna single image (latency), C3 EuRoC 752x480 stereo pairs / 1200 features on two handles, C4 KITTI
1241x376 / 2000 features, C5 3840x2160 / 8000 features / 12 levels with a batch sweep.  Device-resident synthetic
frames, CUDA events on the launching stream, 3 warm-up + 5 timed passes.  One JSON object on stdout."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import synth_frame  # noqa: E402
import extractorb_b200 as ex  # noqa: E402

PEAK = 6550.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def frames_for(w, h, n, nbase=8, seed0=900):
    base = torch.from_numpy(np.stack([synth_frame(seed0 + i, w, h) for i in range(nbase)]))
    g = torch.Generator().manual_seed(seed0)
    out = torch.empty((n, h, w), dtype=torch.uint8)
    for i in range(n):
        out[i] = torch.roll(base[i % nbase], (int(torch.randint(0, h, (1,), generator=g)), int(torch.randint(0, w, (1,), generator=g))), (0, 1))
    return out.cuda()


def level_bytes(w, h, nl, nf):
    tot, cw, ch = w * h, float(w), float(h)
    sf = 1.0
    s = 0
    for l in range(nl):
        lw, lh = int(round(w / sf)), int(round(h / sf))
        s += (lw + 38) * (lh + 38)
        sf *= 1.2
    return tot + s + 60 * nf


def throughput(w, h, nf, nl, F, group, passes=5):
    fr = frames_for(w, h, F)
    e = ex.ORBextractor(nf, 1.2, nl, 20, 7, max_batch=group)
    cap = e.max_keypoints(w, h)
    k = torch.empty((F, cap, 7), dtype=torch.float32, device="cuda")
    d = torch.empty((F, cap, 32), dtype=torch.uint8, device="cuda")
    c = torch.zeros((F, 2), dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream()
    def step():
        e.extract_batch_raw(fr.data_ptr(), ex.MEM_DEVICE, F, w, h, w, w * h, (0, 0), k.data_ptr(), d.data_ptr(), cap, c.data_ptr(), ex.MEM_DEVICE, s.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(passes):
        step()
    e1.record(s)
    torch.cuda.synchronize()
    fps = F * passes / (e0.elapsed_time(e1) * 1e-3)
    kp = float(c[:, 0].float().mean())
    e.close()
    b = level_bytes(w, h, nl, nf)
    return {"frames_per_s": fps, "us_per_frame": 1e6 / fps, "frames": F, "group": group, "mean_keypoints": kp,
            "algorithmic_bytes_per_frame": b, "path_roofline_frac": b * fps / 1e9 / PEAK}


out = {"hbm_peak_gbs": PEAK}
# C1: the reference's own demo image, one operator() call at a time, host to host
with np.load(os.path.join(ROOT, "tests", "golden", "images.npz")) as z:
    luna = z["luna"]
e = ex.ORBextractor(1000, 1.2, 8, 20, 7)
ts = []
for i in range(350):
    t0 = time.perf_counter()
    ret, kps, desc = e(luna, None, (0, 1000))
    ts.append((time.perf_counter() - t0) * 1e6)
out["C1_luna_512x512_1000"] = {"p50_us": float(np.percentile(ts[50:], 50)), "p99_us": float(np.percentile(ts[50:], 99)), "keypoints": int(len(kps)),
                               "what": "ORBextractor.__call__ (orbx_extract), host image in, host keypoints/descriptors out, wall clock"}
e.close()
out["C2_tum_640x480_1000"] = throughput(640, 480, 1000, 8, 2048, 256)
# C3: stereo pairs = two images per frame; the two handles of bench_stereo.py are the latency view, this is the batch view
r = throughput(752, 480, 1200, 8, 2048, 256)
r["stereo_pairs_per_s"] = r["frames_per_s"] / 2
out["C3_euroc_752x480_1200"] = r
out["C4_kitti_1241x376_2000"] = throughput(1241, 376, 2000, 8, 1024, 128)
sweep = {}
for F in (1, 2, 4, 8, 16, 32, 64):
    sweep[str(F)] = throughput(3840, 2160, 8000, 12, F, min(F, 16), passes=3 if F >= 16 else 5)
out["C5_4k_8000_12levels_batch_sweep"] = sweep
print(json.dumps(out))
