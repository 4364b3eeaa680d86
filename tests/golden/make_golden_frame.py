#!/usr/bin/env python
"""tests/golden/make_golden_frame.py -- regenerates tests/golden/frame_*.npz and frame_kat.npz.

Run in the build container only (needs /root/reference for oracle/_ref/ref_extract_bump + ref_frame, and python
cv2 4.13.0 for the undistortPoints KAT).  The goldens are the outputs of the UNMODIFIED reference functions
Frame::UndistortKeyPoints / ComputeImageBounds / AssignFeaturesToGrid / GetFeaturesInArea (src/Frame.cc) and
ORBmatcher::SearchForInitialization (src/ORBmatcher.cc) on keypoints produced by the unmodified extractor.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import FRAME_CAMERAS, FRAME_CASES, random_frame_pair, second_view, synth_frame  # noqa: E402
from oracle import refio  # noqa: E402

COMBOS = [(100, True), (30, False), (10, True)]   # (windowSize, mbCheckOrientation); nnratio 0.9 as in Tracking.cc


def main():
    for case, cam, seed, nf, ang, dx, dy in FRAME_CASES:
        w, h, K, dist = FRAME_CAMERAS[cam]
        im1 = synth_frame(seed, w, h)
        im2 = second_view(im1, ang, dx, dy, seed + 100)
        r1 = refio.run_reference(im1, nfeatures=nf, lap=(0, 1000), bump=True)[0]
        r2 = refio.run_reference(im2, nfeatures=nf, lap=(0, 1000), bump=True)[0]
        out = dict(n1=np.int64(len(r1["kps"])), n2=np.int64(len(r2["kps"])))
        for win, chk in COMBOS:
            r = refio.run_reference_frame(w, h, *K, dist, r1["kps"], r1["desc"], r2["kps"], r2["desc"], win, 0.9, chk)
            tag = "w%d_o%d" % (win, int(chk))
            out["nmatches_" + tag] = np.int64(r["nmatches"])
            out["matches12_" + tag] = r["matches12"]
            out["prev_" + tag] = r["prev_matched"]
            print(case, tag, len(r1["kps"]), len(r2["kps"]), r["nmatches"])
        out["bounds"] = r["bounds"]
        for k in (1, 2):
            out["un_xy%d" % k] = np.stack([r["keys_un%d" % k]["x"], r["keys_un%d" % k]["y"]], 1)
            out["cell_start%d" % k] = r["cell_start%d" % k]
            out["cell_items%d" % k] = r["cell_items%d" % k]
        np.savez_compressed(os.path.join(HERE, "frame_%s.npz" % case), **out)

    # image-free case: ambiguous descriptors, take-overs, rotation filter
    k1, d1, k2, d2 = random_frame_pair(7)
    w, h, K, dist = FRAME_CAMERAS["tum640"]
    out = {}
    for win, chk in COMBOS + [(100, False)]:
        r = refio.run_reference_frame(w, h, *K, dist, k1, d1, k2, d2, win, 0.9, chk)
        tag = "w%d_o%d" % (win, int(chk))
        out["nmatches_" + tag] = np.int64(r["nmatches"]); out["matches12_" + tag] = r["matches12"]; out["prev_" + tag] = r["prev_matched"]
        print("random", tag, r["nmatches"])
    out["bounds"] = r["bounds"]
    out["un_xy1"] = np.stack([r["keys_un1"]["x"], r["keys_un1"]["y"]], 1)
    out["cell_start2"], out["cell_items2"] = r["cell_start2"], r["cell_items2"]
    np.savez_compressed(os.path.join(HERE, "frame_random.npz"), **out)

    # cv::undistortPoints KAT from cv2 4.13.0
    import cv2
    rng = np.random.default_rng(413)
    kat = dict(cv2_version=np.array(cv2.__version__))
    for cam, (w, h, K, dist) in FRAME_CAMERAS.items():
        if dist[0] == 0:
            continue
        pts = (rng.random((3000, 2)) * [w, h]).astype(np.float32)
        pts = np.vstack([pts, np.float32([[0, 0], [w, 0], [0, h], [w, h], [-40, -30], [w + 50, h + 50]])])
        Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
        D = np.array(dist, np.float32).reshape(-1, 1)
        kat["pts_" + cam] = pts
        kat["und_" + cam] = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, D, None, Km).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "frame_kat.npz"), **kat)
    print("frame_kat written")


if __name__ == "__main__":
    main()
