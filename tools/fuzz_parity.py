#!/usr/bin/env python
"""tools/fuzz_parity.py -- randomised parity sweep of the CUDA path against the oracle (not part of the test-suite: it
runs for as long as asked).  Random image sizes / contents / constructor arguments / lapping windows; single-frame calls
(the CUDA-graph latency path), batches of 2, 3 and 9 frames (graph replay, 512-thread quadtrees, the two-stream pipeline),
every stage compared (pyramid planes with borders, FAST candidates, blurred levels, per-level keypoints) on a random
subset.  Prints one line per failure and a summary; exit code 1 if anything differed.

    python tools/fuzz_parity.py [--seconds 60] [--seed 0] [--max-w 1400 --max-h 1000 --max-levels 10]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import synth_frame  # noqa: E402
import extractorb_b200 as ex  # noqa: E402
from oracle import pyoracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--max-w", type=int, default=1400)
ap.add_argument("--max-h", type=int, default=1000)
ap.add_argument("--max-levels", type=int, default=10)
args = ap.parse_args()
rng = np.random.default_rng(args.seed)
with np.load(os.path.join(ROOT, "tests", "golden", "images.npz")) as z:
    fixtures = [z[k] for k in z.files]


def random_image(w, h):
    kind = rng.integers(0, 5)
    if kind == 0:
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == 1:
        f = fixtures[rng.integers(0, len(fixtures))]
        reps = (h // f.shape[0] + 1, w // f.shape[1] + 1)
        return np.ascontiguousarray(np.tile(f, reps)[:h, :w])
    if kind == 2:
        return (synth_frame(int(rng.integers(0, 10000)), w, h) // int(rng.integers(2, 16)) + int(rng.integers(0, 100))).astype(np.uint8)
    if kind == 3:
        img = synth_frame(int(rng.integers(0, 10000)), w, h)
        img[rng.integers(0, h):, :] = int(rng.integers(0, 256))        # a flat band: empty cells, minThFAST retries
        return img
    return synth_frame(int(rng.integers(0, 10000)), w, h)


def compare(g, o, img, lap, nl, deep):
    oret, okps, odesc = o.extract(img, lap)
    gret, gkps, gdesc = g(img, None, lap)
    bad = None
    if gret != oret or len(gkps) != len(okps):
        bad = "final outputs: ret %d vs %d, n %d vs %d" % (gret, oret, len(gkps), len(okps))
    elif gkps.tobytes() != okps.tobytes():
        f = [k for k in gkps.dtype.names if not np.array_equal(gkps[k], okps[k])]
        bad = "final outputs: keypoint fields %s differ (%d rows)" % (f, int(np.sum(gkps[f[0]] != okps[f[0]])))
    elif not np.array_equal(gdesc, odesc):
        bad = "final outputs: %d descriptor rows differ" % int(np.sum(np.any(gdesc != odesc, axis=1)))
    if deep or bad:
        for l in range(nl):
            if not np.array_equal(g.pyramid_level(l, with_border=True), o.level_plane(l)):
                return "pyramid level %d" % l
            gx, gy, gs = g.level_candidates(l)
            ox, oy, os_ = o.level_candidates(l)
            if not (np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gs, os_)):
                return "candidates level %d" % l
            ob = o.level_blur(l)
            if ob is not None and not np.array_equal(g.blurred_level(l), ob):
                return "blur level %d" % l
            if g.level_keypoints(l).tobytes() != o.level_keypoints(l).tobytes():
                gl, ol = g.level_keypoints(l), o.level_keypoints(l)
                if len(gl) != len(ol):
                    return "keypoints level %d: count %d vs %d" % (l, len(gl), len(ol))
                f = [k for k in gl.dtype.names if not np.array_equal(gl[k], ol[k])]
                return "keypoints level %d: fields %s differ" % (l, f)
    return bad


t0 = time.time()
n_cases = n_fail = n_reject = 0
while time.time() - t0 < args.seconds:
    scale = float(rng.choice([1.1, 1.2, 1.2, 1.2, 1.3, 1.5, 2.0]))
    nl = int(rng.integers(1, args.max_levels + 1))
    w, h = int(rng.integers(70, args.max_w)), int(rng.integers(70, args.max_h))
    nf = int(rng.integers(50, 4000 if args.max_w <= 1400 else 12000))
    ini, mn = int(rng.integers(8, 40)), int(rng.integers(3, 20))
    lap = [(0, 0), (0, 1000), (int(rng.integers(0, w)), int(rng.integers(0, 2 * w)))][rng.integers(0, 3)]
    img = random_image(w, h)
    try:
        o = pyoracle.OracleExtractor(nf, scale, nl, ini, mn)
        oret, _, _ = o.extract(img, lap)
    except Exception:
        oret = -2
    if oret == -2:                      # a level too small for the cell grid / aspect < 0.5: UB in the reference
        g = ex.ORBextractor(nf, scale, nl, ini, mn)
        try:
            g(img, None, lap)
            print("FAIL: GPU accepted a geometry the oracle rejects", w, h, scale, nl)
            n_fail += 1
        except ex.OrbxError:
            n_reject += 1
        g.close()
        continue
    g = ex.ORBextractor(nf, scale, nl, ini, mn, max_batch=int(rng.choice([1, 2, 4, 16])))
    try:
        g.max_keypoints(w, h)
    except ex.OrbxError as e:           # documented limit: ~3 190 features on one level
        if "too large for the quadtree" in str(e):
            n_reject += 1
            g.close()
            continue
        raise
    n_cases += 1
    why = compare(g, o, img, lap, nl, deep=rng.random() < 0.5)
    if why is None and rng.random() < 0.4:
        F = int(rng.choice([2, 3, 9]))
        frames = np.stack([img] + [random_image(w, h) for _ in range(F - 1)])
        counts, kps, desc = g.extract_batch_host(frames, lap)
        for f in range(F):
            oret, okps, odesc = o.extract(frames[f], lap)
            n = counts[f, 0]
            if (counts[f, 1], n) != (oret, len(okps)) or kps[f, :n].tobytes() != okps.tobytes() or not np.array_equal(desc[f, :n], odesc):
                why = "batch of %d, frame %d" % (F, f)
                break
    if why is not None:
        n_fail += 1
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "fuzz_fail_%d.npz" % n_fail), img=img, args=np.array([w, h, nf, nl, ini, mn, lap[0], lap[1]]),
                            scale=np.float64(scale))
        print("FAIL (%s): w=%d h=%d nf=%d scale=%.1f nl=%d ini=%d min=%d lap=%s" % (why, w, h, nf, scale, nl, ini, mn, lap))
    g.close()
print("fuzz: %d cases compared, %d rejected geometries, %d failures, %.0f s" % (n_cases, n_reject, n_fail, time.time() - t0))
sys.exit(1 if n_fail else 0)
