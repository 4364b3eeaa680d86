"""Next rows of the path: the per-frame post-processing of every Frame
constructor (Frame::UndistortKeyPoints / ComputeImageBounds / AssignFeaturesToGrid / GetFeaturesInArea, reference
src/Frame.cc:383-417, :655-812) and ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:705-814).
CPU: the oracle's restatement against goldens produced by the unmodified reference functions (oracle/_ref/ref_frame)
and against cv2 4.13.0 for cv::undistortPoints.  GPU: orbx_frame_* / orbx_search_for_initialization against the
oracle and the same goldens.  Everything is compared bit for bit (keypoint coordinates as raw float bits)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from common import FRAME_CAMERAS, FRAME_CASES, random_frame_pair, second_view, synth_frame

COMBOS = [(100, True), (30, False), (10, True)]


def load(case):
    with np.load(os.path.join(GOLDEN, "frame_%s.npz" % case)) as z:
        return {k: z[k] for k in z.files}


_extract_cache = {}


def frame_inputs(oracle, case):
    """Keypoints + descriptors of the two views of a case, from the oracle extractor (itself pinned to the reference)."""
    if case not in _extract_cache:
        _, cam, seed, nf, ang, dx, dy = next(c for c in FRAME_CASES if c[0] == case)
        w, h, K, dist = FRAME_CAMERAS[cam]
        im1 = synth_frame(seed, w, h)
        im2 = second_view(im1, ang, dx, dy, seed + 100)
        o = oracle.OracleExtractor(nf, 1.2, 8, 20, 7)
        _, k1, d1 = o.extract(im1, (0, 1000))
        k1, d1 = k1.copy(), d1.copy()
        _, k2, d2 = o.extract(im2, (0, 1000))
        _extract_cache[case] = (w, h, K, dist, k1, d1, k2.copy(), d2.copy())
    return _extract_cache[case]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ------------------------------------------------------------------------------------------------------ CPU
def test_undistort_points_kat(oracle):
    """cv::undistortPoints restated (oracle/cv_prims.c) == cv2 4.13.0 on committed vectors."""
    with np.load(os.path.join(GOLDEN, "frame_kat.npz")) as z:
        for cam, (w, h, K, dist) in FRAME_CAMERAS.items():
            if dist[0] == 0:
                continue
            got = oracle.undistort_points(z["pts_" + cam], *K, dist)
            assert np.array_equal(bits(got), bits(z["und_" + cam])), cam


def test_undistort_points_live_against_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(99)
    for cam, (w, h, K, dist) in FRAME_CAMERAS.items():
        pts = (rng.random((2000, 2)) * [w, h]).astype(np.float32)
        Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, np.array(dist, np.float32).reshape(-1, 1), None, Km).reshape(-1, 2)
        assert np.array_equal(bits(oracle.undistort_points(pts, *K, dist)), bits(ref)), cam


@pytest.mark.parametrize("case", [c[0] for c in FRAME_CASES])
def test_oracle_frame_equals_reference_golden(oracle, case):
    g = load(case)
    w, h, K, dist, k1, d1, k2, d2 = frame_inputs(oracle, case)
    assert len(k1) == int(g["n1"]) and len(k2) == int(g["n2"])
    cal = oracle.make_calib(*K, dist, w, h)
    assert np.array_equal(bits([cal["min_x"][0], cal["max_x"][0], cal["min_y"][0], cal["max_y"][0]]), bits(g["bounds"]))
    u1, u2 = oracle.undistort_keypoints(cal, k1), oracle.undistort_keypoints(cal, k2)
    for u, k, tag in ((u1, k1, "1"), (u2, k2, "2")):
        assert np.array_equal(bits(np.stack([u["x"], u["y"]], 1)), bits(g["un_xy" + tag]))
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(u[f], k[f])
        s, it = oracle.assign_grid(cal, u)
        assert np.array_equal(s, g["cell_start" + tag]) and np.array_equal(it, g["cell_items" + tag])
    s2, i2 = oracle.assign_grid(cal, u2)
    for win, chk in COMBOS:
        tag = "w%d_o%d" % (win, int(chk))
        n, m12, prev = oracle.search_for_initialization(cal, u1, d1, u2, d2, s2, i2, None, win, 0.9, chk)
        assert n == int(g["nmatches_" + tag]) and np.array_equal(m12, g["matches12_" + tag])
        assert np.array_equal(bits(prev), bits(g["prev_" + tag]))
        assert n == int((m12 >= 0).sum())
    assert int(g["nmatches_w100_o1"]) > 100


def test_oracle_random_pair_equals_reference_golden(oracle):
    g = load("random")
    k1, d1, k2, d2 = random_frame_pair(7)
    w, h, K, dist = FRAME_CAMERAS["tum640"]
    cal = oracle.make_calib(*K, dist, w, h)
    u1, u2 = oracle.undistort_keypoints(cal, k1), oracle.undistort_keypoints(cal, k2)
    s2, i2 = oracle.assign_grid(cal, u2)
    assert np.array_equal(s2, g["cell_start2"]) and np.array_equal(i2, g["cell_items2"])
    for win, chk in COMBOS + [(100, False)]:
        tag = "w%d_o%d" % (win, int(chk))
        n, m12, prev = oracle.search_for_initialization(cal, u1, d1, u2, d2, s2, i2, None, win, 0.9, chk)
        assert n == int(g["nmatches_" + tag]) and np.array_equal(m12, g["matches12_" + tag]) and np.array_equal(bits(prev), bits(g["prev_" + tag]))
    # the rotation filter removes matches on this input (otherwise the case would not exercise it)
    assert int(g["nmatches_w100_o1"]) < int(g["nmatches_w100_o0"])


def test_features_in_area_against_brute_force(oracle):
    """Frame::GetFeaturesInArea restated: the returned set equals a brute-force window test over the keypoints inside the
    grid, and the order is (cell column, cell row, index)."""
    k1, d1, k2, d2 = random_frame_pair(3, 800, 900)
    w, h, K, dist = FRAME_CAMERAS["euroc752"]
    cal = oracle.make_calib(*K, dist, 752, 480)
    u = oracle.undistort_keypoints(cal, k2)
    s, it = oracle.assign_grid(cal, u)
    cell_of = np.full(len(u), -1)
    for c in range(64 * 48):
        cell_of[it[s[c]:s[c + 1]]] = c
    rng = np.random.default_rng(5)
    for _ in range(60):
        x, y, r = float(rng.random() * 752), float(rng.random() * 480), float(rng.choice([5, 10, 30, 100]))
        lo, hi = (0, 0) if rng.random() < 0.5 else (-1, -1)
        got = oracle.features_in_area(cal, u, s, it, x, y, r, lo, hi)
        ok = (np.abs(u["x"] - np.float32(x)) < r) & (np.abs(u["y"] - np.float32(y)) < r) & (cell_of >= 0)
        if lo == 0:
            ok &= u["octave"] == 0
        assert sorted(got.tolist()) == np.nonzero(ok)[0].tolist()
        keys = [(cell_of[i], i) for i in got]
        assert keys == sorted(keys)


def test_live_reference_frame_if_present(oracle):
    from oracle import refio
    if not refio.have_ref_frame():
        pytest.skip("oracle/_ref/ref_frame not built (needs /root/reference)")
    for seed, n1, n2, cam, win, ratio, chk in ((11, 700, 650, "euroc752", 60, 0.9, True), (12, 2500, 1800, "tum640", 100, 0.9, True),
                                               (13, 300, 4000, "nodist640", 10, 0.6, False), (14, 1200, 1, "euroc752", 300, 1.0, True)):
        k1, d1, k2, d2 = random_frame_pair(seed, n1, n2)
        w, h, K, dist = FRAME_CAMERAS[cam]
        r = refio.run_reference_frame(w, h, *K, dist, k1, d1, k2, d2, win, ratio, chk)
        cal = oracle.make_calib(*K, dist, w, h)
        u1, u2 = oracle.undistort_keypoints(cal, k1), oracle.undistort_keypoints(cal, k2)
        assert u1.tobytes() == r["keys_un1"].tobytes() and u2.tobytes() == r["keys_un2"].tobytes()
        s2, i2 = oracle.assign_grid(cal, u2)
        assert np.array_equal(s2, r["cell_start2"]) and np.array_equal(i2, r["cell_items2"])
        n, m12, prev = oracle.search_for_initialization(cal, u1, d1, u2, d2, s2, i2, None, win, ratio, chk)
        assert n == r["nmatches"] and np.array_equal(m12, r["matches12"]) and np.array_equal(bits(prev), bits(r["prev_matched"])), seed


# ------------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gpu_ext():
    import extractorb_b200 as ex
    e = ex.ORBextractor(1000, 1.2, 8, 20, 7)
    yield ex, e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c[0] for c in FRAME_CASES])
def test_gpu_frame_matches_reference_and_oracle(oracle, gpu_ext, case):
    ex, ext = gpu_ext
    g = load(case)
    w, h, K, dist, k1, d1, k2, d2 = frame_inputs(oracle, case)
    cal = ex.image_bounds(ext, *K, dist, w, h)
    assert np.array_equal(bits([cal.min_x, cal.max_x, cal.min_y, cal.max_y]), bits(g["bounds"]))
    ocal = oracle.make_calib(*K, dist, w, h)
    u1, s1, i1 = ex.undistort_grid(ext, cal, k1)
    u2, s2, i2 = ex.undistort_grid(ext, cal, k2)
    assert u1.tobytes() == oracle.undistort_keypoints(ocal, k1).tobytes() and u2.tobytes() == oracle.undistort_keypoints(ocal, k2).tobytes()
    for u, s, it, tag in ((u1, s1, i1, "1"), (u2, s2, i2, "2")):
        assert np.array_equal(bits(np.stack([u["x"], u["y"]], 1)), bits(g["un_xy" + tag]))
        assert np.array_equal(s, g["cell_start" + tag]) and np.array_equal(it, g["cell_items" + tag])
    for win, chk in COMBOS:
        tag = "w%d_o%d" % (win, int(chk))
        n, m12, prev = ex.search_for_initialization(ext, cal, u1, d1, u2, d2, s2, i2, None, win, 0.9, chk)
        assert n == int(g["nmatches_" + tag])
        assert np.array_equal(m12, g["matches12_" + tag])
        assert np.array_equal(bits(prev), bits(g["prev_" + tag]))


@pytest.mark.gpu
def test_gpu_random_pairs_match_oracle(oracle, gpu_ext):
    """Ambiguous descriptors: take-overs of earlier matches, shortlist exhaustion (the filtered re-enumeration path)
    and the rotation filter, on several seeds, windows and ratios; second rounds reuse the updated vbPrevMatched."""
    ex, ext = gpu_ext
    g = load("random")
    w, h, K, dist = FRAME_CAMERAS["tum640"]
    cal = ex.image_bounds(ext, *K, dist, w, h)
    ocal = oracle.make_calib(*K, dist, w, h)
    for seed, n1, n2 in ((7, 1500, 1600), (8, 3000, 2500), (9, 400, 5000), (10, 2000, 60)):
        k1, d1, k2, d2 = random_frame_pair(seed, n1, n2)
        u1, _, _ = ex.undistort_grid(ext, cal, k1)
        u2, s2, i2 = ex.undistort_grid(ext, cal, k2)
        os2, oi2 = oracle.assign_grid(ocal, oracle.undistort_keypoints(ocal, k2))
        assert np.array_equal(s2, os2) and np.array_equal(i2, oi2)
        for win, ratio, chk in ((100, 0.9, True), (30, 0.9, False), (10, 0.9, True), (100, 0.9, False), (200, 0.6, True), (1000, 1.0, True)):
            n, m12, prev = ex.search_for_initialization(ext, cal, u1, d1, u2, d2, s2, i2, None, win, ratio, chk)
            on, om12, oprev = oracle.search_for_initialization(ocal, u1, d1, u2, d2, s2, i2, None, win, ratio, chk)
            assert n == on and np.array_equal(m12, om12) and np.array_equal(bits(prev), bits(oprev)), (seed, win, ratio, chk)
            if seed == 7 and ratio == 0.9 and "nmatches_w%d_o%d" % (win, int(chk)) in g:
                assert np.array_equal(m12, g["matches12_w%d_o%d" % (win, int(chk))])
            # second round from the updated vbPrevMatched, as Tracking does frame after frame
            n_b, m12_b, prev_b = ex.search_for_initialization(ext, cal, u1, d1, u2, d2, s2, i2, prev, win, ratio, chk)
            on_b, om12_b, oprev_b = oracle.search_for_initialization(ocal, u1, d1, u2, d2, s2, i2, oprev, win, ratio, chk)
            assert n_b == on_b and np.array_equal(m12_b, om12_b) and np.array_equal(bits(prev_b), bits(oprev_b))


@pytest.mark.gpu
def test_gpu_frame_edge_cases(oracle, gpu_ext):
    ex, ext = gpu_ext
    w, h, K, dist = FRAME_CAMERAS["euroc752"]
    cal = ex.image_bounds(ext, *K, dist, w, h)
    ocal = oracle.make_calib(*K, dist, w, h)
    k1, d1, k2, d2 = random_frame_pair(4, 300, 280, w, h)
    # keypoints outside the undistorted image bounds are dropped from the grid (PosInGrid false)
    k2["x"][:40] = np.linspace(-300, 1200, 40).astype(np.float32)
    k2["y"][40:80] = np.linspace(-300, 900, 40).astype(np.float32)
    u2, s2, i2 = ex.undistort_grid(ext, cal, k2)
    ou2 = oracle.undistort_keypoints(ocal, k2)
    os2, oi2 = oracle.assign_grid(ocal, ou2)
    assert u2.tobytes() == ou2.tobytes() and np.array_equal(s2, os2) and np.array_equal(i2, oi2) and len(i2) < len(k2)
    u1, _, _ = ex.undistort_grid(ext, cal, k1)
    n, m12, prev = ex.search_for_initialization(ext, cal, u1, d1, u2, d2, s2, i2, None, 100, 0.9, True)
    on, om12, oprev = oracle.search_for_initialization(ocal, u1, d1, u2, d2, os2, oi2, None, 100, 0.9, True)
    assert n == on and np.array_equal(m12, om12) and np.array_equal(bits(prev), bits(oprev))
    # empty second frame, empty first frame, no level-0 keypoint
    e_u, e_s, e_i = ex.undistort_grid(ext, cal, k2[:0])
    assert len(e_u) == 0 and e_s[-1] == 0 and len(e_i) == 0
    n, m12, _ = ex.search_for_initialization(ext, cal, u1, d1, e_u, d2[:0], e_s, e_i, None, 100, 0.9, True)
    assert n == 0 and (m12 == -1).all()
    n, m12, _ = ex.search_for_initialization(ext, cal, u1[:0], d1[:0], u2, d2, s2, i2, None, 100, 0.9, True)
    assert n == 0 and len(m12) == 0
    hi = u1.copy(); hi["octave"] = 2
    n, m12, _ = ex.search_for_initialization(ext, cal, hi, d1, u2, d2, s2, i2, None, 100, 0.9, True)
    assert n == 0 and (m12 == -1).all()
    # dist[0] == 0: mvKeysUn = mvKeys, bounds = image rectangle (src/Frame.cc:750, :805)
    w0, h0, K0, dist0 = FRAME_CAMERAS["nodist640"]
    c0 = ex.image_bounds(ext, *K0, dist0, w0, h0)
    assert (c0.min_x, c0.max_x, c0.min_y, c0.max_y) == (0.0, 640.0, 0.0, 480.0)
    u0, _, _ = ex.undistort_grid(ext, c0, k1)
    assert u0.tobytes() == k1.tobytes()
    # bad arguments fail loudly
    bad = ex.FrameCalib()
    with pytest.raises(ex.OrbxError):
        ex.undistort_grid(ext, bad, k1)


@pytest.mark.gpu
def test_gpu_extract_frame_equals_separate_calls(oracle, gpu_ext):
    """orbx_extract_frame (extraction + undistort + grid with the keypoints resident on the GPU) == orbx_extract followed by
    orbx_frame_undistort_grid == the oracle chain, for a distorted and an undistorted camera and both lapping conventions."""
    ex, _ = gpu_ext
    for cam, seed, nf, lap in (("tum640", 3, 2000, (0, 1000)), ("euroc752", 11, 1200, (0, 0)), ("nodist640", 21, 1000, (100, 300))):
        w, h, K, dist = FRAME_CAMERAS[cam]
        img = synth_frame(seed, w, h)
        ext = ex.ORBextractor(nf, 1.2, 8, 20, 7)
        cal = ex.image_bounds(ext, *K, dist, w, h)
        ret, kps, desc, un, start, items = ex.extract_frame(ext, img, cal, lap)
        ret2, kps2, desc2 = ext(img, None, lap)
        un2, start2, items2 = ex.undistort_grid(ext, cal, kps2)
        assert ret == ret2 and kps.tobytes() == kps2.tobytes() and np.array_equal(desc, desc2)
        assert un.tobytes() == un2.tobytes() and np.array_equal(start, start2) and np.array_equal(items, items2)
        o = oracle.OracleExtractor(nf, 1.2, 8, 20, 7)
        oret, okps, odesc = o.extract(img, lap)
        ocal = oracle.make_calib(*K, dist, w, h)
        oun = oracle.undistort_keypoints(ocal, okps)
        ostart, oitems = oracle.assign_grid(ocal, oun)
        assert ret == oret and kps.tobytes() == okps.tobytes() and np.array_equal(desc, odesc)
        assert un.tobytes() == oun.tobytes() and np.array_equal(start, ostart) and np.array_equal(items, oitems)
        ext.close()


@pytest.mark.gpu
def test_gpu_search_for_initialization_device_memory(oracle, gpu_ext):
    """orbx_search_for_initialization_mem with every array resident on the device == the host-memory call == the oracle."""
    import torch
    ex, ext = gpu_ext
    w, h, K, dist = FRAME_CAMERAS["tum640"]
    cal = ex.image_bounds(ext, *K, dist, w, h)
    k1, d1, k2, d2 = random_frame_pair(9, 1200, 1300, w, h)
    u1, _, _ = ex.undistort_grid(ext, cal, k1)
    u2, s2, i2 = ex.undistort_grid(ext, cal, k2)
    n_h, m12_h, prev_h = ex.search_for_initialization(ext, cal, u1, d1, u2, d2, s2, i2, None, 100, 0.9, True)
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    tu1, td1, tu2, td2, ts2, ti2 = t(u1), t(d1), t(u2), t(d2), t(s2), t(i2)
    prev0 = np.stack([u1["x"], u1["y"]], 1).astype(np.float32)
    tprev = t(prev0)
    tm12 = torch.full((len(u1),), -5, dtype=torch.int32, device=dev)
    n_d = ex.search_for_initialization_raw(ext, cal, tu1.data_ptr(), td1.data_ptr(), len(u1), tu2.data_ptr(), td2.data_ptr(), len(u2), ts2.data_ptr(),
                                           ti2.data_ptr(), tprev.data_ptr(), 100, 0.9, True, tm12.data_ptr(), ex.MEM_DEVICE)
    assert n_d == n_h and n_h > 50
    assert np.array_equal(tm12.cpu().numpy(), m12_h)
    assert np.array_equal(tprev.cpu().numpy().view(np.float32).reshape(-1, 2).view(np.uint32), prev_h.view(np.uint32))
    ocal = oracle.make_calib(*K, dist, w, h)
    on, om12, _ = oracle.search_for_initialization(ocal, u1, d1, u2, d2, s2, i2, None, 100, 0.9, True)
    assert n_d == on and np.array_equal(m12_h, om12)
