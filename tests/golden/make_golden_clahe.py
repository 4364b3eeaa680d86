#!/usr/bin/env python
"""tests/golden/make_golden_clahe.py -- regenerates tests/golden/clahe_kat.npz from python cv2 4.13.0 (build container only):
cv::createCLAHE(clip, Size(tx, ty))->apply on fixture images and seeded arrays.  Full outputs are kept for the small
inputs, CRC-32 plus one 64x64 crop for the large ones."""
import os
import sys
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
from common import clahe_cases  # noqa: E402


def main():
    with np.load(os.path.join(HERE, "images.npz")) as z:
        images = {k: z[k] for k in z.files}
    kat = dict(cv2_version=np.array(cv2.__version__))
    for name, img, clip, tx, ty in clahe_cases(images):
        ref = cv2.createCLAHE(clip, (tx, ty)).apply(img)
        kat["crc_" + name] = np.uint64(zlib.crc32(ref.tobytes()))
        if ref.size <= 20000:
            kat["out_" + name] = ref
        else:
            kat["crop_" + name] = ref[40:104, 72:136].copy()
        print(name, img.shape, clip, tx, ty, hex(zlib.crc32(ref.tobytes())))
    np.savez_compressed(os.path.join(HERE, "clahe_kat.npz"), **kat)


if __name__ == "__main__":
    main()
