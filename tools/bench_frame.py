#!/usr/bin/env python
"""tools/bench_frame.py -- measurement for the rows after the extractor in monocular initialisation (SURVEY.md 8(f)
ranks 3 and 2): Frame::UndistortKeyPoints + AssignFeaturesToGrid and ORBmatcher::SearchForInitialization on the
752x480 / 5000-feature configuration (the initialisation extractor uses 5 x nFeatures, reference
src/Tracking.cc:772-774), host buffers in and out.  Prints one JSON line with p50 latencies next to the CPU oracle's
time for the same inputs (parity asserted)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import FRAME_CAMERAS, second_view, synth_frame  # noqa: E402
import extractorb_b200 as ex  # noqa: E402
from oracle import pyoracle  # noqa: E402

out = {}
for name, cam, nf in (("752x480, 5000 features", "euroc752", 5000), ("640x480, 1000 features", "tum640", 1000)):
    w, h, K, dist = FRAME_CAMERAS[cam]
    im1 = synth_frame(11, w, h)
    im2 = second_view(im1, -12.0, -15, 9, 111)
    e = ex.ORBextractor(nf, 1.2, 8, 20, 7)
    _, k1, d1 = e(im1, None, (0, 1000))
    k1, d1 = k1.copy(), d1.copy()
    _, k2, d2 = e(im2, None, (0, 1000))
    cal = ex.image_bounds(e, *K, dist, w, h)
    t_grid, t_search = [], []
    for i in range(250):
        t0 = time.perf_counter()
        u2, s2, i2 = ex.undistort_grid(e, cal, k2)
        t1 = time.perf_counter()
        if i == 0:
            u1, _, _ = ex.undistort_grid(e, cal, k1)
            t1 = time.perf_counter()
        n, m12, prev = ex.search_for_initialization(e, cal, u1, d1, u2, d2, s2, i2, None, 100, 0.9, True)
        t2 = time.perf_counter()
        if i >= 50:
            t_grid.append((t1 - t0) * 1e6)
            t_search.append((t2 - t1) * 1e6)
    t_sep, t_fused = [], []
    for i in range(250):
        t0 = time.perf_counter()
        _, ka, da = e(im2, None, (0, 1000))
        ua, sa, ia = ex.undistort_grid(e, cal, ka)
        t1 = time.perf_counter()
        rb, kb, db, ub, sb, ib = ex.extract_frame(e, im2, cal, (0, 1000))
        t2 = time.perf_counter()
        if i >= 50:
            t_sep.append((t1 - t0) * 1e6)
            t_fused.append((t2 - t1) * 1e6)
    assert kb.tobytes() == ka.tobytes() and ub.tobytes() == ua.tobytes() and np.array_equal(sb, sa) and np.array_equal(ib, ia)
    ocal = pyoracle.make_calib(*K, dist, w, h)
    c_grid, c_search = [], []
    for _ in range(20):
        t0 = time.perf_counter()
        ou2 = pyoracle.undistort_keypoints(ocal, k2)
        os2, oi2 = pyoracle.assign_grid(ocal, ou2)
        t1 = time.perf_counter()
        on, om12, oprev = pyoracle.search_for_initialization(ocal, u1, d1, ou2, d2, os2, oi2, None, 100, 0.9, True)
        t2 = time.perf_counter()
        c_grid.append((t1 - t0) * 1e6)
        c_search.append((t2 - t1) * 1e6)
    assert ou2.tobytes() == u2.tobytes() and np.array_equal(os2, s2) and np.array_equal(oi2, i2)
    assert on == n and np.array_equal(om12, m12) and prev.tobytes() == oprev.tobytes()
    out[name] = {"keypoints": [int(len(k1)), int(len(k2))], "level0_keypoints_frame1": int((k1["octave"] == 0).sum()), "matches": int(n), "shortlist_fallbacks": int(e._L.orbx_last_init_fallbacks(e._h)),
                 "undistort_grid_p50_us": float(np.median(t_grid)), "extract_then_grid_p50_us": float(np.median(t_sep)),
                 "extract_frame_fused_p50_us": float(np.median(t_fused)), "search_for_initialization_p50_us": float(np.median(t_search)),
                 "cpu_oracle_undistort_grid_p50_us": float(np.median(c_grid)), "cpu_oracle_search_p50_us": float(np.median(c_search))}
    e.close()
out["cpu"] = "oracle C restatement, 1 thread"
out["parity"] = "mvKeysUn, mGrid, vnMatches12, vbPrevMatched bit-identical to the oracle"
print(json.dumps(out))
