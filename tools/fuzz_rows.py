#!/usr/bin/env python
"""tools/fuzz_rows.py -- randomised parity sweep of the rows around the extractor against the oracle: ComputeStereoMatches,
UndistortKeyPoints + AssignFeaturesToGrid (+ the fused orbx_extract_frame), SearchForInitialization and CLAHE, on random
images, sizes, cameras, feature counts, windows and ratios.  Not part of the test-suite (runs for as long as asked).

    python tools/fuzz_rows.py [--seconds 60] [--seed 0]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import FRAME_CAMERAS, make_stereo_pair, random_frame_pair, second_view, synth_frame  # noqa: E402
import extractorb_b200 as ex  # noqa: E402
from oracle import pyoracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
ap.add_argument("--seed", type=int, default=0)
args = ap.parse_args()
rng = np.random.default_rng(args.seed)
fails = {"stereo": 0, "frame": 0, "init": 0, "clahe": 0}
runs = dict.fromkeys(fails, 0)


def report(row, what, **kw):
    fails[row] += 1
    print("FAIL %s (%s): %s" % (row, what, kw), flush=True)


def fuzz_stereo():
    w, h = [(640, 480), (752, 480), (int(rng.integers(500, 1300)), int(rng.integers(300, 480)))][rng.integers(0, 3)]
    nf = int(rng.integers(200, 2500))
    left = synth_frame(int(rng.integers(0, 10000)), w, h)
    if rng.random() < 0.3:
        left = (left // int(rng.integers(2, 6)) + 60).astype(np.uint8)
    right = make_stereo_pair(left, int(rng.integers(0, 1000)))
    mbf = float(rng.choice([20.0, 40.0, 80.0])); mb = mbf / float(rng.choice([300.0, 435.0, 700.0]))
    eL, eR = ex.ORBextractor(nf, 1.2, 8, 20, 7), ex.ORBextractor(nf, 1.2, 8, 20, 7)
    _, kl, dl = eL(left, None, (0, 0))
    _, kr, dr = eR(right, None, (0, 0))
    u, d, n = ex.stereo_match(eL, eR, kl, dl, kr, dr, mb, mbf)
    oL, oR = pyoracle.OracleExtractor(nf, 1.2, 8, 20, 7), pyoracle.OracleExtractor(nf, 1.2, 8, 20, 7)
    _, okl, odl = oL.extract(left, (0, 0))
    _, okr, odr = oR.extract(right, (0, 0))
    pl = [oL.level_plane(l)[19:-19, 19:-19] for l in range(8)]
    pr = [oR.level_plane(l)[19:-19, 19:-19] for l in range(8)]
    ou, od = pyoracle.stereo_match(okl, odl, okr, odr, oL.mvScaleFactor, oL.mvInvScaleFactor, pl, pr, mb, mbf)
    if not (np.array_equal(u, ou) and np.array_equal(d, od)):
        report("stereo", "%d of %d entries differ" % (int(np.sum((u != ou) | (d != od))), len(u)), w=w, h=h, nf=nf, mb=mb, mbf=mbf)
    eL.close(); eR.close()


def fuzz_frame_init():
    cam = list(FRAME_CAMERAS)[rng.integers(0, len(FRAME_CAMERAS))]
    w, h, K, dist = FRAME_CAMERAS[cam]
    e = ex.ORBextractor(int(rng.integers(300, 5000)), 1.2, 8, 20, 7)
    cal = ex.image_bounds(e, *K, dist, w, h)
    ocal = pyoracle.make_calib(*K, dist, w, h)
    if rng.random() < 0.5:
        im1 = synth_frame(int(rng.integers(0, 10000)), w, h)
        im2 = second_view(im1, float(rng.uniform(-25, 25)), int(rng.integers(-30, 30)), int(rng.integers(-30, 30)), int(rng.integers(0, 1000)),
                          int(rng.integers(0, 8)))
        lap = [(0, 1000), (0, 0)][rng.integers(0, 2)]
        r1, k1, d1, u1, s1, i1 = ex.extract_frame(e, im1, cal, lap)
        r2, k2, d2, u2, s2, i2 = ex.extract_frame(e, im2, cal, lap)
        o = pyoracle.OracleExtractor(e.nfeatures, 1.2, 8, 20, 7)
        for (r, k, d, u, s, i, im) in ((r1, k1, d1, u1, s1, i1, im1), (r2, k2, d2, u2, s2, i2, im2)):
            runs["frame"] += 1
            oret, okps, odesc = o.extract(im, lap)
            ou = pyoracle.undistort_keypoints(ocal, okps)
            os_, oi = pyoracle.assign_grid(ocal, ou)
            if not (r == oret and k.tobytes() == okps.tobytes() and np.array_equal(d, odesc) and u.tobytes() == ou.tobytes()
                    and np.array_equal(s, os_) and np.array_equal(i, oi)):
                report("frame", "extract_frame", cam=cam, nf=e.nfeatures, lap=lap)
    else:
        k1, d1, k2, d2 = random_frame_pair(int(rng.integers(0, 100000)), int(rng.integers(1, 4000)), int(rng.integers(1, 6000)), w, h)
        u1, _, _ = ex.undistort_grid(e, cal, k1)
        u2, s2, i2 = ex.undistort_grid(e, cal, k2)
        runs["frame"] += 1
        ou2 = pyoracle.undistort_keypoints(ocal, k2)
        os2, oi2 = pyoracle.assign_grid(ocal, ou2)
        if not (u2.tobytes() == ou2.tobytes() and np.array_equal(s2, os2) and np.array_equal(i2, oi2)):
            report("frame", "undistort_grid", cam=cam, n=len(k2))
    win, ratio, chk = int(rng.choice([5, 10, 30, 100, 300])), float(rng.choice([0.6, 0.8, 0.9, 1.0])), bool(rng.integers(0, 2))
    prev = None
    for _ in range(2):
        runs["init"] += 1
        n, m12, pv = ex.search_for_initialization(e, cal, u1, d1, u2, d2, s2, i2, prev, win, ratio, chk)
        on, om12, opv = pyoracle.search_for_initialization(ocal, u1, d1, u2, d2, s2, i2, prev, win, ratio, chk)
        if not (n == on and np.array_equal(m12, om12) and pv.tobytes() == opv.tobytes()):
            report("init", "%d vs %d matches, %d entries differ" % (n, on, int(np.sum(m12 != om12))), cam=cam, n1=len(u1), n2=len(u2), win=win,
                   ratio=ratio, chk=chk)
        prev = opv
    e.close()


def fuzz_clahe(e):
    h, w = int(rng.integers(8, 900)), int(rng.integers(8, 1300))
    kind = rng.integers(0, 3)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8) if kind == 0 else synth_frame(int(rng.integers(0, 1000)), max(w, 32), max(h, 32))[:h, :w].copy()
    if kind == 2:
        img = (img // int(rng.integers(2, 30)) + int(rng.integers(0, 200))).astype(np.uint8)
    clip, tx, ty = float(rng.choice([0.0, 0.3, 1.0, 3.0, 40.0, 1000.0])), int(rng.integers(1, 33)), int(rng.integers(1, 33))
    if not np.array_equal(ex.clahe(e, img, clip, (tx, ty)), pyoracle.clahe(img, clip, (tx, ty))):
        report("clahe", "pixels differ", w=w, h=h, clip=clip, tx=tx, ty=ty)


t0 = time.time()
ec = ex.ORBextractor(500, 1.2, 4, 20, 7)
while time.time() - t0 < args.seconds:
    which = rng.integers(0, 4)
    if which == 0:
        runs["stereo"] += 1
        fuzz_stereo()
    elif which == 1:
        fuzz_frame_init()
    else:
        runs["clahe"] += 1
        fuzz_clahe(ec)
ec.close()
print("fuzz rows: runs %s, failures %s, %.0f s" % (runs, fails, time.time() - t0))
sys.exit(1 if any(fails.values()) else 0)
