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


@pytest.mark.gpu
def test_gpu_stereo_batch_device_resident(oracle):
    """orbx_stereo_match_batch: three stereo pairs extracted as one launch group per camera with device outputs, matched on the
    device with nothing travelling to the host in between.  mvuRight / mvDepth per pair == the single-pair host call (itself
    equal to the reference goldens above) and == the oracle for one of the pairs."""
    import torch
    import extractorb_b200 as ex
    lefts = np.stack([synth_frame(9, 752, 480), synth_frame(12, 752, 480), synth_frame(9, 752, 480)[::-1].copy()])
    rights = np.stack([make_stereo_pair(l, 1 + i) for i, l in enumerate(lefts)])
    eL, eR = ex.ORBextractor(1200, 1.2, 8, 20, 7, max_batch=3), ex.ORBextractor(1200, 1.2, 8, 20, 7, max_batch=3)
    cap = eL.max_keypoints(752, 480)
    dev = torch.device("cuda")
    out = {}
    for name, e, imgs in (("l", eL, lefts), ("r", eR, rights)):
        d_img = torch.from_numpy(imgs).to(dev)
        k = torch.zeros((3, cap, 7), dtype=torch.float32, device=dev)
        d = torch.zeros((3, cap, 32), dtype=torch.uint8, device=dev)
        c = torch.zeros((3, 2), dtype=torch.int32, device=dev)
        e.extract_batch_raw(d_img.data_ptr(), ex.MEM_DEVICE, 3, 752, 480, 752, 752 * 480, (0, 0), k.data_ptr(), d.data_ptr(), cap, c.data_ptr(),
                            ex.MEM_DEVICE, None)
        out[name] = (k, d, c)
    u = torch.full((3, cap), -7.0, dtype=torch.float32, device=dev)
    z = torch.full((3, cap), -7.0, dtype=torch.float32, device=dev)
    nm = torch.zeros(3, dtype=torch.int32, device=dev)
    ex.stereo_match_batch_raw(eL, eR, 3, out["l"][0].data_ptr(), out["l"][1].data_ptr(), out["l"][2].data_ptr(), out["r"][0].data_ptr(),
                              out["r"][1].data_ptr(), out["r"][2].data_ptr(), cap, STEREO_MB, STEREO_MBF, u.data_ptr(), z.data_ptr(), nm.data_ptr())
    eL.synchronize()
    u, z, nm = u.cpu().numpy(), z.cpu().numpy(), nm.cpu().numpy()
    cl, cr = out["l"][2].cpu().numpy(), out["r"][2].cpu().numpy()
    sL, sR = ex.ORBextractor(1200, 1.2, 8, 20, 7), ex.ORBextractor(1200, 1.2, 8, 20, 7)
    for p in range(3):
        _, kl, dl = sL(lefts[p], None, (0, 0))
        _, kr, dr = sR(rights[p], None, (0, 0))
        assert cl[p, 0] == len(kl) and cr[p, 0] == len(kr)
        u1, z1, n1 = ex.stereo_match(sL, sR, kl, dl, kr, dr, STEREO_MB, STEREO_MBF)
        assert np.array_equal(u[p, :len(kl)], u1) and np.array_equal(z[p, :len(kl)], z1) and nm[p] == n1, p
        assert np.all(u[p, len(kl):] == -7.0)
        assert n1 > 300
    kl, dl, kr, dr, uo, do = oracle_stereo(oracle, lefts[1], rights[1])
    assert np.array_equal(u[1, :len(kl)], uo) and np.array_equal(z[1, :len(kl)], do)
    for e in (eL, eR, sL, sR):
        e.close()
