"""The oracle's end-to-end restatement (oracle/orb_oracle.c) pinned against the UNMODIFIED reference:
committed golden outputs (tests/golden/ref_*.npz, produced by oracle/_ref/ref_extract_bump) and, when the
prebuilt binaries are present, a live re-run of the reference.  CPU only."""
import zlib

import numpy as np
import pytest

from conftest import golden_cases, load_golden
from common import synth_frame, kp_bytes_equal


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_equals_reference_golden(oracle, images, case):
    g = load_golden(case)
    p = g["params"]
    o = oracle.OracleExtractor(p["nfeatures"], p["scale"], p["nlevels"], p["ini"], p["mn"])
    ret, kps, desc = o.extract(images[g["image_name"]], p["lap"])
    assert ret == int(g["ret"])
    assert kp_bytes_equal(kps, g["kps"])                      # every field of every cv::KeyPoint, byte for byte
    assert np.array_equal(desc, g["desc"])
    for l in range(p["nlevels"]):
        plane = o.level_plane(l)
        assert tuple(plane.shape) == tuple(g["pyr_shape"][l])
        assert zlib.crc32(plane.tobytes()) == int(g["pyr_crc"][l])
        assert kp_bytes_equal(o.level_keypoints(l), g["level_kps"][l])


def test_known_answer_counts():
    """Counts the survey reproduced with cv2 primitives; 1420 is the number in the
    reference's README screenshot (img_folder/Screenshot.png: 'ORB_SLAM3 has total 1420 keypoints')."""
    assert len(load_golden("luna_1000_mono")["kps"]) == 1009
    assert load_golden("luna_1000_mono")["counts"].tolist() == [219, 181, 153, 127, 105, 88, 75, 61]
    assert len(load_golden("robot866_1000_mono")["kps"]) == 840
    assert len(load_golden("tum_room4_1500")["kps"]) == 1420


def test_constructor_tables(oracle):
    o = oracle.OracleExtractor(1000, 1.2, 8, 20, 7)
    assert o.mnFeaturesPerLevel.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert o.umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert o.mvScaleFactor[1] == np.float32(1.2000000476837158)
    sizes = []
    o.extract(np.zeros((480, 640), np.uint8))
    for l in range(8):
        sizes.append(o.level_size(l))
    assert sizes == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
    o2 = oracle.OracleExtractor(8000, 1.2, 12, 20, 7)
    assert o2.mnFeaturesPerLevel.tolist() == [1502, 1251, 1043, 869, 724, 604, 503, 419, 349, 291, 243, 202]


def test_two_ended_output_order(oracle, images):
    """vLappingArea semantics (ORBextractor.cc:1147-1156): lapping keypoints are filled from the back."""
    img = images["robot866"]
    o = oracle.OracleExtractor(1000, 1.2, 8, 20, 7)
    r_all, k_all, d_all = o.extract(img, (0, 1000))           # mono call path: every keypoint is "lapping"
    r_none, k_none, d_none = o.extract(img, (0, 0))           # stereo call path: none is (x >= 19 always)
    assert r_all == 0 and r_none == len(k_none) == len(k_all)
    assert kp_bytes_equal(k_all[::-1], k_none) and np.array_equal(d_all[::-1], d_none)
    r_mid, k_mid, _ = o.extract(img, (200, 420))
    inside = (k_mid["x"] >= 200) & (k_mid["x"] <= 420)
    assert not inside[:r_mid].any() and inside[r_mid:].all()


def test_distribute_tie_rule_matters(oracle):
    """Equal-size nodes: the later-created node is split first (monotonic-allocator rule); the restatement
    must be deterministic and keep exactly one key per final node."""
    rng = np.random.default_rng(5)
    pts = np.unique(rng.integers(0, [600, 440], (2500, 2)), axis=0)
    sc = rng.integers(7, 30, len(pts))
    a = oracle.distribute(pts[:, 0], pts[:, 1], sc, 16, 616, 16, 456, 200)
    b = oracle.distribute(pts[:, 0], pts[:, 1], sc, 16, 616, 16, 456, 200)
    assert np.array_equal(a, b) and 200 <= len(a) <= 202 and len(set(a.tolist())) == len(a)
    few = oracle.distribute(pts[:50, 0], pts[:50, 1], sc[:50], 16, 616, 16, 456, 200)
    assert sorted(few.tolist()) == list(range(50))            # fewer candidates than quota: all are kept


def test_small_level_is_an_error_not_ub(oracle):
    o = oracle.OracleExtractor(100, 1.2, 8, 20, 7)
    with pytest.raises(ValueError):
        o.extract(synth_frame(0, 80, 70))                     # level 7 is 22x20: no 30-px cell (UB in the reference)


def test_live_reference_binary_if_present(oracle):
    from oracle import refio
    if not refio.have_ref(bump=True):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    img = synth_frame(42, 400, 300)
    r = refio.run_reference(img, nfeatures=600, nlevels=6, lap=(100, 250), dump_pyr=True)[0]
    o = oracle.OracleExtractor(600, 1.2, 6, 20, 7)
    ret, kps, desc = o.extract(img, (100, 250))
    assert ret == r["ret"] and kp_bytes_equal(kps, r["kps"]) and np.array_equal(desc, r["desc"])
    for l in range(6):
        assert np.array_equal(o.level_plane(l), r["pyr"][l])
