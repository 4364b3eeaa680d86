#!/usr/bin/env python
"""tools/sanitize_case.py -- a small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck):
single frame, a 3-frame batch through the host pipeline (groups of 2), the stage-wise methods and the
stand-alone DistributeOctTree, all checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import synth_frame  # noqa: E402
import extractorb_b200 as ex  # noqa: E402
from oracle import pyoracle  # noqa: E402

frames = np.stack([synth_frame(300 + i, 320, 240) for i in range(3)])
o = pyoracle.OracleExtractor(400, 1.2, 4, 20, 7)
e = ex.ORBextractor(400, 1.2, 4, 20, 7, max_batch=2)
counts, kps, desc = e.extract_batch_host(frames, (50, 200))
for f in range(3):
    ret, okps, odesc = o.extract(frames[f], (50, 200))
    n = counts[f, 0]
    assert (counts[f, 1], n) == (ret, len(okps))
    assert kps[f, :n].tobytes() == okps.tobytes() and np.array_equal(desc[f, :n], odesc)
ret, k1, d1 = e(frames[0], None, (0, 0))
e.ComputePyramid(frames[1])
lv = e.ComputeKeyPointsOctTree()
assert sum(len(x) for x in lv) > 100
keys = np.zeros(500, ex.KP_DTYPE)
rng = np.random.default_rng(0)
pts = np.unique(rng.integers(0, [280, 200], (500, 2)), axis=0)
keys = keys[:len(pts)]
keys["x"], keys["y"], keys["response"] = pts[:, 0], pts[:, 1], rng.integers(7, 60, len(pts))
out = e.DistributeOctTree(keys, 16, 296, 16, 216, 100)
idx = pyoracle.distribute(pts[:, 0], pts[:, 1], keys["response"].astype(np.int32), 16, 296, 16, 216, 100)
assert out.tobytes() == keys[idx].tobytes()
e.close()
print("sanitize case ok")
