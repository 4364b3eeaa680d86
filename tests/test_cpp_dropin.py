"""The C++ drop-in class (include/ORBextractor.h, ORB_SLAM3::ORBextractor + the global-namespace alias of
include/ORBExtractor.h) driven by tests/cpp/dropin_main.cpp, the twin of oracle/ref_main.cpp: both programs
read the same frame file and write the same result file, so the comparison is program against program."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT, load_golden
from common import kp_bytes_equal, desc_bit_agreement

DEMO = os.path.join(ROOT, "tests", "cpp", "dropin_main")


@pytest.fixture(scope="module")
def demo():
    from extractorb_b200 import build
    build.build_host()
    assert os.access(DEMO, os.X_OK)
    return DEMO


def run_demo(demo, frames, p, mode="six", dump=True):
    from oracle import refio
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.orbf"), os.path.join(td, "out.orbr")
        refio.write_frames(fin, frames)
        cmd = [demo, "run", fin, fout, str(p["nfeatures"]), repr(float(p["scale"])), str(p["nlevels"]), str(p["ini"]),
               str(p["mn"]), str(p["lap"][0]), str(p["lap"][1]), "1" if dump else "0", mode]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stderr)
        return refio.read_results(fout), r.stderr


def test_cpp_class_fails_loudly_without_gpu(demo, images):
    """No CPU fallback: without a CUDA device operator() returns -1 and reports why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    p = dict(nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7, lap=(0, 0))
    res, err = run_demo(demo, images["luna"], p, dump=False)
    assert res[0]["ret"] == -1 and len(res[0]["kps"]) == 0
    assert "no CUDA device" in err or "CUDA" in err


@pytest.mark.gpu
@pytest.mark.parametrize("case,mode", [("luna_1000_mono", "six"), ("robot866_1000_lap", "six"), ("tum_room4_1500", "six"),
                                       ("robot866_1000_mono", "five"), ("robot2196_1200_stereo", "threads"),
                                       ("luna_500_5lv_s15", "six")])
def test_cpp_class_matches_reference(demo, images, case, mode):
    g = load_golden(case)
    p = g["params"]
    res, _ = run_demo(demo, images[g["image_name"]], p, mode=mode)
    r = res[0]
    assert r["ret"] == int(g["ret"])
    ref = g["kps"]
    assert len(r["kps"]) == len(ref)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(r["kps"][f], ref[f]), f
    assert np.max(np.abs(r["kps"]["angle"] - ref["angle"]), initial=0.0) <= 1e-3
    assert desc_bit_agreement(r["desc"], g["desc"]) >= 0.999
    import zlib
    for l in range(p["nlevels"]):
        assert zlib.crc32(r["pyr"][l].tobytes()) == int(g["pyr_crc"][l])          # mvImagePyramid incl. border
        assert len(r["level_kps"][l]) == int(g["counts"][l])                       # allLevelsKeypoints / ComputeKeyPointsOctTree
        for f in ("x", "y", "size", "response", "octave"):
            assert np.array_equal(r["level_kps"][l][f], g["level_kps"][l][f])


@pytest.mark.gpu
def test_cpp_class_multiple_frames_and_empty(demo, images):
    frames = np.stack([images["robot866"], np.full((480, 640), 77, np.uint8), images["robot2196"]])
    p = dict(nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7, lap=(0, 1000))
    res, _ = run_demo(demo, frames, p, dump=False)
    g = load_golden("robot866_1000_mono")
    assert kp_bytes_equal(res[0]["kps"][["x", "y", "size", "response", "octave", "class_id"]],
                          g["kps"][["x", "y", "size", "response", "octave", "class_id"]])
    assert res[1]["ret"] == 0 and len(res[1]["kps"]) == 0                            # flat frame: descriptors released
    assert len(res[2]["kps"]) > 500


@pytest.mark.gpu
def test_cpp_stereo_matches_reference(demo, images):
    """include/ORBstereo.h: two extractors on two host threads + ComputeStereoMatches, against the golden of the
    unmodified reference chain (tests/golden/stereo_robot866.npz)."""
    import struct
    from oracle import refio
    from common import make_stereo_pair, STEREO_MBF
    left = images["robot866"]
    frames = np.stack([left, make_stereo_pair(left, 1)])
    with np.load(os.path.join(ROOT, "tests", "golden", "stereo_robot866.npz")) as z:
        gu, gd = z["u_right"], z["depth"]
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.orbf"), os.path.join(td, "out.bin")
        refio.write_frames(fin, frames)
        # argv: run in out nfeatures scale nlevels ini min <bf> <fx> dump mode
        r = subprocess.run([demo, "run", fin, fout, "1200", "1.2", "8", "20", "7", repr(STEREO_MBF), "435.0", "0", "stereo"],
                           capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stderr)
        buf = open(fout, "rb").read()
    n = struct.unpack_from("<i", buf, 0)[0]
    u = np.frombuffer(buf, "<f4", n, 4)
    d = np.frombuffer(buf, "<f4", n, 4 + 4 * n)
    kept = struct.unpack_from("<i", buf, 4 + 8 * n)[0]
    assert n == len(gu) and np.array_equal(u, gu) and np.array_equal(d, gd) and kept == int((gd > 0).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tum640", "nodist640"])
def test_cpp_frame_matches_reference(demo, case):
    """include/ORBframe.h: extraction of two views, ComputeImageBounds, UndistortAndAssignToGrid (mvKeysUn + mGrid) and
    SearchForInitialization, against the golden of the unmodified reference chain (tests/golden/frame_*.npz)."""
    from oracle import refio
    from common import FRAME_CAMERAS, FRAME_CASES, second_view, synth_frame
    _, cam, seed, nf, ang, dx, dy = next(c for c in FRAME_CASES if c[0] == case)
    w, h, K, dist = FRAME_CAMERAS[cam]
    im1 = synth_frame(seed, w, h)
    frames = np.stack([im1, second_view(im1, ang, dx, dy, seed + 100)])
    with np.load(os.path.join(ROOT, "tests", "golden", "frame_%s.npz" % case)) as z:
        g = {k: z[k] for k in z.files}
    d = list(dist) + [0.0] * (5 - len(dist))
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.orbf"), os.path.join(td, "out.bin")
        refio.write_frames(fin, frames)
        cmd = [demo, "run", fin, fout, str(nf), "1.2", "8", "20", "7", "0", "1000", "0", "frame"] + [repr(float(v)) for v in K] + \
              [str(len(dist))] + [repr(float(v)) for v in d] + ["100", "1"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stderr)
        out = refio.parse_frame_output(open(fout, "rb").read())
    assert len(out["keys_un1"]) == int(g["n1"]) and len(out["keys_un2"]) == int(g["n2"])
    assert out["bounds"].tobytes() == g["bounds"].tobytes()
    for k in ("1", "2"):
        xy = np.stack([out["keys_un" + k]["x"], out["keys_un" + k]["y"]], 1)
        assert xy.tobytes() == g["un_xy" + k].tobytes()
        assert np.array_equal(out["cell_start" + k], g["cell_start" + k]) and np.array_equal(out["cell_items" + k], g["cell_items" + k])
    assert out["nmatches"] == int(g["nmatches_w100_o1"]) and np.array_equal(out["matches12"], g["matches12_w100_o1"])
    assert out["prev_matched"].tobytes() == g["prev_w100_o1"].tobytes()


@pytest.mark.gpu
def test_cpp_clahe_matches_cv2_golden(demo, images):
    """include/ORBclahe.h: ORB_SLAM3::ApplyCLAHE(ext, image, out, 3.0, Size(8, 8)) against the cv2 golden (the demos' call,
    reference src/orb_extractor/main_orb_extractor.cpp:19-22)."""
    import zlib
    from oracle import refio
    with np.load(os.path.join(ROOT, "tests", "golden", "clahe_kat.npz")) as z:
        crcs = {"robot866": int(z["crc_robot866_3_8x8"])}
    frames = np.stack([images["robot866"], images["robot2196"]])
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.orbf"), os.path.join(td, "out.bin")
        refio.write_frames(fin, frames)
        # argv: run in out nfeatures <clip> nlevels ini min <tilesX> <tilesY> dump mode
        r = subprocess.run([demo, "run", fin, fout, "1000", "3.0", "8", "20", "7", "8", "8", "0", "clahe"], capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stderr)
        out = np.frombuffer(open(fout, "rb").read(), np.uint8).reshape(frames.shape)
    assert zlib.crc32(out[0].tobytes()) == crcs["robot866"]
    from oracle import pyoracle
    assert np.array_equal(out[1], pyoracle.clahe(frames[1], 3.0, (8, 8)))
