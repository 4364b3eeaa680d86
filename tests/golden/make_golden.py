#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerate the committed golden fixtures.

Runs in the build container only (needs /root/reference, python cv2 4.13.0 and the binaries built by
oracle/Makefile).  Produces:

  images.npz      decoded 8-bit gray fixture images of the reference (pic/luna.jpg, two pic/robot
                  frames, one pic/TUM frame) -- decoded once here so that tests do not depend on a JPEG/PNG
                  decoder version.
  ref_<case>.npz  outputs of the UNMODIFIED reference (oracle/_ref/ref_extract_bump: ORBextractor.cc
                  compiled verbatim + monotonic allocator) for each (image, config) case: return value,
                  keypoints, descriptors, per-level keypoints, crc32 of every bordered pyramid plane.
  prims_kat.npz   known-answer vectors of the five OpenCV primitives produced by cv2 4.13.0.

usage: python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refio  # noqa: E402

REF = "/root/reference/pic"
IMAGES = {
    "luna": "luna.jpg",
    "robot866": "robot/866_im.jpg",
    "robot2196": "robot/2196_im.jpg",
    "tum_room4": "TUM/dataset-room4_512_16/mav0/cam0/data/1520531124150444163.png",
}
# (case, image, nfeatures, scale, nlevels, ini, min, lap)
CASES = [
    ("luna_1000_mono", "luna", 1000, 1.2, 8, 20, 7, (0, 1000)),      # BASELINE configs[0], mono call path (Frame.cc:307)
    ("luna_1000_stereo", "luna", 1000, 1.2, 8, 20, 7, (0, 0)),       # rectified stereo call path (Frame.cc:109-110)
    ("robot866_1000_mono", "robot866", 1000, 1.2, 8, 20, 7, (0, 1000)),
    ("robot866_1000_lap", "robot866", 1000, 1.2, 8, 20, 7, (200, 420)),  # fisheye-style lapping window
    ("robot2196_1200_stereo", "robot2196", 1200, 1.2, 8, 20, 7, (0, 0)),
    ("tum_room4_1500", "tum_room4", 1500, 1.2, 8, 20, 7, (0, 1000)),  # the README screenshot: 1420 keypoints
    ("robot866_7500_demo", "robot866", 7500, 1.2, 8, 20, 7, (0, 1000)),  # demos use 5*1500 (main_orb_extractor.cpp:43)
    ("luna_500_5lv_s15", "luna", 500, 1.5, 5, 25, 10, (0, 0)),
    ("robot866_800_3lv_s19", "robot866", 800, 1.9, 3, 20, 7, (0, 0)),   # source columns 1 or 2 apart per destination column
    ("luna_600_3lv_s25", "luna", 600, 2.5, 3, 20, 7, (0, 0)),           # ... 2 or 3 apart: the resize kernel's one-column pass
    ("kitti1241_2000", "robot866_kitti1241", 2000, 1.2, 8, 20, 7, (0, 0)),     # BASELINE C4 geometry (tests/common.derived_images)
    ("euroc752_1200", "robot2196_euroc752", 1200, 1.2, 8, 20, 7, (0, 0)),      # BASELINE C3 geometry
]


def main():
    imgs = {k: cv2.imread(os.path.join(REF, v), cv2.IMREAD_GRAYSCALE) for k, v in IMAGES.items()}
    for k, v in imgs.items():
        assert v is not None and v.dtype == np.uint8, k
    np.savez_compressed(os.path.join(HERE, "images.npz"), **imgs)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import derived_images
    imgs = derived_images(imgs)
    for case, img, nf, sc, nl, ini, mn, lap in CASES:
        r = refio.run_reference(imgs[img], nfeatures=nf, scale=sc, nlevels=nl, ini=ini, mn=mn, lap=lap, dump_pyr=True, bump=True)[0]
        out = dict(image=np.array(img), cfg=np.array([nf, nl, ini, mn, lap[0], lap[1]], np.int64), scale=np.float32(sc),
                   ret=np.int64(r["ret"]), kps=r["kps"], desc=r["desc"], counts=r["counts"],
                   pyr_crc=np.array([zlib.crc32(p.tobytes()) for p in r["pyr"]], np.uint64),
                   pyr_shape=np.array([p.shape for p in r["pyr"]], np.int64))
        for l, k in enumerate(r["level_kps"]):
            out["level_kps_%d" % l] = k
        np.savez_compressed(os.path.join(HERE, "ref_%s.npz" % case), **out)
        print(case, r["ret"], len(r["kps"]), r["counts"].tolist())

    # Frame::ComputeStereoMatches goldens: the unmodified extractor (ref_extract_bump) on a synthetic rectified
    # pair, then the unmodified stereo function (oracle/_ref/ref_stereo) on its outputs.
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import make_stereo_pair, synth_frame, STEREO_MB, STEREO_MBF
    for case, left in (("robot866", imgs["robot866"]), ("synth752", synth_frame(9, 752, 480))):
        right = make_stereo_pair(left, 1)
        rl = refio.run_reference(left, nfeatures=1200, lap=(0, 0), dump_pyr=True)[0]
        rr = refio.run_reference(right, nfeatures=1200, lap=(0, 0), dump_pyr=True)[0]
        pl = [p[19:-19, 19:-19] for p in rl["pyr"]]
        pr = [p[19:-19, 19:-19] for p in rr["pyr"]]
        sc = np.ones(8, np.float32)
        for i in range(1, 8):
            sc[i] = np.float32(np.float64(sc[i - 1]) * np.float64(np.float32(1.2)))
        isc = (np.float32(1.0) / sc).astype(np.float32)
        u, d = refio.run_reference_stereo(rl["kps"], rl["desc"], rr["kps"], rr["desc"], sc, isc, pl, pr, STEREO_MB, STEREO_MBF)
        np.savez_compressed(os.path.join(HERE, "stereo_%s.npz" % case), u_right=u, depth=d, n_left=np.int64(len(rl["kps"])),
                            n_right=np.int64(len(rr["kps"])), mb=np.float32(STEREO_MB), mbf=np.float32(STEREO_MBF))
        print("stereo", case, len(rl["kps"]), len(rr["kps"]), int((d > 0).sum()))

    # primitive KATs from cv2 4.13.0
    rng = np.random.default_rng(4130)
    src = rng.integers(0, 256, (97, 133), dtype=np.uint8)
    nat = imgs["robot866"][100:260, 200:420].copy()
    kat = dict(cv2_version=np.array(cv2.__version__), src=src, nat=nat)
    kat["resize_src_111x81"] = cv2.resize(src, (111, 81), interpolation=cv2.INTER_LINEAR)
    kat["resize_nat_183x133"] = cv2.resize(nat, (183, 133), interpolation=cv2.INTER_LINEAR)
    kat["resize_nat_110x80"] = cv2.resize(nat, (110, 80), interpolation=cv2.INTER_LINEAR)   # exact 2x -> INTER_AREA path
    kat["resize_nat_116x84"] = cv2.resize(nat, (116, 84), interpolation=cv2.INTER_LINEAR)   # scale 1.90: source steps of 1 and 2
    kat["resize_nat_88x64"] = cv2.resize(nat, (88, 64), interpolation=cv2.INTER_LINEAR)     # scale 2.5: source steps of 2 and 3
    kat["resize_nat_73x53"] = cv2.resize(nat, (73, 53), interpolation=cv2.INTER_LINEAR)     # scale 3.01
    kat["resize_src_70x51"] = cv2.resize(src, (70, 51), interpolation=cv2.INTER_LINEAR)     # scale 1.90 on noise
    kat["border_src"] = cv2.copyMakeBorder(src, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    kat["blur_src"] = cv2.GaussianBlur(src, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    kat["blur_nat"] = cv2.GaussianBlur(nat, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    for name, im in (("src", src), ("nat", nat)):
        for th in (7, 20):
            det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            kp = det.detect(im)
            kat["fast_%s_%d" % (name, th)] = np.array([(k.pt[0], k.pt[1], k.response) for k in kp], np.int32).reshape(-1, 3)
    yx = rng.integers(-200000, 200000, (4000, 2)).astype(np.float32)
    yx = np.concatenate([yx, np.array([[0, 0], [0, 1], [1, 0], [-1, 0], [0, -1], [5, 5], [-5, 5], [5, -5], [-5, -5], [-1e-30, 1]], np.float32)])
    kat["atan2_yx"] = yx
    kat["atan2_out"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "prims_kat.npz"), **kat)
    print("prims_kat written")


if __name__ == "__main__":
    main()
