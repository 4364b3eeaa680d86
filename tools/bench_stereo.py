#!/usr/bin/env python
"""tools/bench_stereo.py -- measurement for the next row of the path (Frame::ComputeStereoMatches, reference
src/Frame.cc:813-990) on the 752x480 / 1200-feature stereo configuration (BASELINE configs[2]): one stereo frame =
two extractions (two handles, as the reference's two host threads) + orbx_stereo_match, host buffers in and out.
Prints one JSON line with p50 latencies and the CPU reference / oracle time for the same pair."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import synth_frame, make_stereo_pair, STEREO_MB, STEREO_MBF  # noqa: E402
import extractorb_b200 as ex  # noqa: E402
from oracle import pyoracle  # noqa: E402

left = synth_frame(9, 752, 480)
right = make_stereo_pair(left, 1)
eL, eR = ex.ORBextractor(1200, 1.2, 8, 20, 7), ex.ORBextractor(1200, 1.2, 8, 20, 7)
t_ext, t_st = [], []
for i in range(250):
    t0 = time.perf_counter()
    _, kl, dl = eL(left, None, (0, 0))
    _, kr, dr = eR(right, None, (0, 0))
    t1 = time.perf_counter()
    u, d, n = ex.stereo_match(eL, eR, kl, dl, kr, dr, STEREO_MB, STEREO_MBF)
    t2 = time.perf_counter()
    if i >= 50:
        t_ext.append((t1 - t0) * 1e6)
        t_st.append((t2 - t1) * 1e6)
oL, oR = pyoracle.OracleExtractor(1200, 1.2, 8, 20, 7), pyoracle.OracleExtractor(1200, 1.2, 8, 20, 7)
_, okl, odl = oL.extract(left, (0, 0))
_, okr, odr = oR.extract(right, (0, 0))
pl = [oL.level_plane(l)[19:-19, 19:-19] for l in range(8)]
pr = [oR.level_plane(l)[19:-19, 19:-19] for l in range(8)]
c = []
for _ in range(20):
    t0 = time.perf_counter()
    ou, od = pyoracle.stereo_match(okl, odl, okr, odr, oL.mvScaleFactor, oL.mvInvScaleFactor, pl, pr, STEREO_MB, STEREO_MBF)
    c.append((time.perf_counter() - t0) * 1e6)
assert np.array_equal(u, ou) and np.array_equal(d, od)
print(json.dumps({"workload": "752x480 stereo pair, 1200 features per image", "keypoints_left": int(len(kl)), "keypoints_right": int(len(kr)),
                  "matches_kept": int(n), "extract_pair_p50_us": float(np.median(t_ext)), "stereo_match_p50_us": float(np.median(t_st)),
                  "stereo_frame_p50_us": float(np.median(np.array(t_ext) + np.array(t_st))),
                  "cpu_oracle_stereo_match_p50_us": float(np.median(c)), "cpu": "oracle C restatement, 1 thread (matching only)",
                  "parity": "mvuRight/mvDepth bit-identical to the oracle"}))
