"""Next row of the path: Frame::ComputeStereoMatches (reference src/Frame.cc:813-990).
CPU: the oracle's restatement against goldens produced by the unmodified reference function (oracle/_ref/ref_stereo).
GPU: orbx_stereo_match, with both pyramids resident on the device, against the oracle and the same goldens.
mvuRight / mvDepth are compared bit for bit (every float step is restated without FMA)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from common import make_stereo_pair, synth_frame, STEREO_MB, STEREO_MBF


def left_image(case, images):
    return images["robot866"] if case == "robot866" else synth_frame(9, 752, 480)


def load(case):
    with np.load(os.path.join(GOLDEN, "stereo_%s.npz" % case)) as z:
        return {k: z[k] for k in z.files}


def oracle_stereo(oracle, left, right, nf=1200):
    oL, oR = oracle.OracleExtractor(nf, 1.2, 8, 20, 7), oracle.OracleExtractor(nf, 1.2, 8, 20, 7)
    _, kl, dl = oL.extract(left, (0, 0))
    _, kr, dr = oR.extract(right, (0, 0))
    pl = [oL.level_plane(l)[19:-19, 19:-19] for l in range(8)]
    pr = [oR.level_plane(l)[19:-19, 19:-19] for l in range(8)]
    u, d = oracle.stereo_match(kl, dl, kr, dr, oL.mvScaleFactor, oL.mvInvScaleFactor, pl, pr, STEREO_MB, STEREO_MBF)
    return kl, dl, kr, dr, u, d


@pytest.mark.parametrize("case", ["robot866", "synth752"])
def test_oracle_stereo_equals_reference_golden(oracle, images, case):
    g = load(case)
    left = left_image(case, images)
    kl, dl, kr, dr, u, d = oracle_stereo(oracle, left, make_stereo_pair(left, 1))
    assert len(kl) == int(g["n_left"]) and len(kr) == int(g["n_right"])
    assert np.array_equal(u, g["u_right"]) and np.array_equal(d, g["depth"])
    assert (d > 0).sum() > 500
    # the synthetic pair has a known disparity profile: 6 + 34 * y / h pixels
    ok = d > 0
    disp = kl["x"][ok] - u[ok]
    expect = 6 + 34 * kl["y"][ok] / left.shape[0]
    assert np.median(np.abs(disp - expect)) < 1.0


def test_descriptor_distance_bit_trick(oracle):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    for i in range(200):
        ref = int(np.unpackbits(np.bitwise_xor(a[i], b[i])).sum())
        assert oracle.lib().orb_oracle_descriptor_distance(a[i].ctypes.data, b[i].ctypes.data) == ref


def test_live_reference_stereo_if_present(oracle):
    from oracle import refio
    if not refio.have_ref_stereo():
        pytest.skip("oracle/_ref/ref_stereo not built (needs /root/reference)")
    left = synth_frame(31, 640, 480)
    right = make_stereo_pair(left, 5)
    oL, oR = oracle.OracleExtractor(800, 1.2, 8, 20, 7), oracle.OracleExtractor(800, 1.2, 8, 20, 7)
    _, kl, dl = oL.extract(left, (0, 0))
    _, kr, dr = oR.extract(right, (0, 0))
    pl = [oL.level_plane(l)[19:-19, 19:-19] for l in range(8)]
    pr = [oR.level_plane(l)[19:-19, 19:-19] for l in range(8)]
    u1, d1 = oracle.stereo_match(kl, dl, kr, dr, oL.mvScaleFactor, oL.mvInvScaleFactor, pl, pr, STEREO_MB, STEREO_MBF)
    u2, d2 = refio.run_reference_stereo(kl, dl, kr, dr, oL.mvScaleFactor, oL.mvInvScaleFactor, pl, pr, STEREO_MB, STEREO_MBF)
    assert np.array_equal(u1, u2) and np.array_equal(d1, d2)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["robot866", "synth752"])
def test_gpu_stereo_matches_reference_and_oracle(oracle, images, case):
    import extractorb_b200 as ex
    g = load(case)
    left = left_image(case, images)
    right = make_stereo_pair(left, 1)
    eL, eR = ex.ORBextractor(1200, 1.2, 8, 20, 7), ex.ORBextractor(1200, 1.2, 8, 20, 7)
    _, kl, dl = eL(left, None, (0, 0))
    _, kr, dr = eR(right, None, (0, 0))
    u, d, n = ex.stereo_match(eL, eR, kl, dl, kr, dr, STEREO_MB, STEREO_MBF)
    assert np.array_equal(u, g["u_right"]) and np.array_equal(d, g["depth"])       # unmodified reference chain
    assert n == int((d > 0).sum())
    _, _, _, _, ou, od = oracle_stereo(oracle, left, right)
    assert np.array_equal(u, ou) and np.array_equal(d, od)
    eL.close(); eR.close()


@pytest.mark.gpu
def test_gpu_stereo_edge_cases(oracle):
    import extractorb_b200 as ex
    left = synth_frame(44, 640, 480)
    eL, eR = ex.ORBextractor(600, 1.2, 8, 20, 7), ex.ORBextractor(600, 1.2, 8, 20, 7)
    _, kl, dl = eL(left, None, (0, 0))
    flat = np.full_like(left, 90)
    _, kr, dr = eR(flat, None, (0, 0))                       # right image without keypoints: nothing matches
    u, d, n = ex.stereo_match(eL, eR, kl, dl, kr, dr, STEREO_MB, STEREO_MBF)
    assert n == 0 and np.all(u == -1) and np.all(d == -1)
    # identical images: every window distance is 0, so the median threshold (1.5 * 1.4 * 0) rejects all matches --
    # in the reference too; what matters is that both sides agree
    _, kr, dr = eR(left, None, (0, 0))
    u, d, n = ex.stereo_match(eL, eR, kl, dl, kr, dr, STEREO_MB, STEREO_MBF)
    _, _, _, _, ou, od = oracle_stereo(oracle, left, left, nf=600)
    assert np.array_equal(u, ou) and np.array_equal(d, od) and n == int((od > 0).sum())
    # a one-pixel shift with noise: sub-pixel disparities around 1 px, many of them clamped or rejected
    shifted = np.roll(left, -1, axis=1)
    shifted = np.clip(shifted.astype(np.int32) + np.random.default_rng(3).integers(-2, 3, left.shape), 0, 255).astype(np.uint8)
    _, kr, dr = eR(shifted, None, (0, 0))
    u, d, n = ex.stereo_match(eL, eR, kl, dl, kr, dr, STEREO_MB, STEREO_MBF)
    _, _, _, _, ou, od = oracle_stereo(oracle, left, shifted, nf=600)
    assert np.array_equal(u, ou) and np.array_equal(d, od) and n == int((od > 0).sum()) and n > 50
    other = ex.ORBextractor(600, 1.2, 8, 20, 7)
    other(synth_frame(45, 320, 240), None, (0, 0))
    with pytest.raises(ex.OrbxError):                        # mismatching image sizes
        ex.stereo_match(eL, other, kl, dl, kr, dr, STEREO_MB, STEREO_MBF)
    eL.close(); eR.close(); other.close()
