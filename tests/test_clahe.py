"""Pre-processing row of the demos: cv::createCLAHE(3.0, Size(8, 8))->apply
(reference src/orb_extractor/main_orb_extractor.cpp:19-22, src/clahe/main_clahe.cpp:7-11).  The reference tree holds no
CLAHE code (it is OpenCV's), so the oracle restates OpenCV's algorithm (oracle/cv_prims.c) and is pinned against python
cv2 4.13.0: committed goldens (tests/golden/clahe_kat.npz) and live when cv2 is importable.  GPU: orbx_clahe, byte-exact."""
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN
from common import clahe_cases, synth_frame


@pytest.fixture(scope="module")
def kat():
    with np.load(os.path.join(GOLDEN, "clahe_kat.npz")) as z:
        return {k: z[k] for k in z.files}


def check_golden(kat, name, out):
    assert zlib.crc32(np.ascontiguousarray(out).tobytes()) == int(kat["crc_" + name]), name
    if "out_" + name in kat:
        assert np.array_equal(out, kat["out_" + name]), name
    else:
        assert np.array_equal(out[40:104, 72:136], kat["crop_" + name]), name


def test_oracle_clahe_equals_cv2_golden(oracle, images, kat):
    for name, img, clip, tx, ty in clahe_cases(images):
        check_golden(kat, name, oracle.clahe(img, clip, (tx, ty)))


def test_oracle_clahe_live_against_cv2(oracle, images):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    extra = [("rand_%d" % i, rng.integers(0, 256, (int(rng.integers(20, 200)), int(rng.integers(20, 300))), dtype=np.uint8),
              float(rng.choice([0.5, 2.0, 3.0, 40.0])), int(rng.integers(1, 12)), int(rng.integers(1, 12))) for i in range(12)]
    for name, img, clip, tx, ty in clahe_cases(images) + extra:
        ref = cv2.createCLAHE(clip, (tx, ty)).apply(img)
        assert np.array_equal(oracle.clahe(img, clip, (tx, ty)), ref), (name, img.shape, clip, tx, ty)


def test_clahe_properties(oracle):
    """Size-independent properties: a constant image maps to a constant; the output depends on the pixel only through the
    four surrounding tile LUTs, so equal pixels at the same position of two images that share those tiles map equally."""
    flat = np.full((96, 128), 90, np.uint8)
    out = oracle.clahe(flat, 3.0, (8, 8))
    assert len(np.unique(out)) == 1
    img = synth_frame(4, 128, 96)
    a = oracle.clahe(img, 3.0, (8, 8))
    assert a.shape == img.shape and a.dtype == np.uint8
    # monotone: inside one image, LUTs are non-decreasing, so at a fixed position a brighter input is never darker
    brighter = np.clip(img.astype(np.int32) + 0, 0, 255).astype(np.uint8)
    assert np.array_equal(oracle.clahe(brighter, 3.0, (8, 8)), a)


# ------------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gpu_ext():
    import extractorb_b200 as ex
    e = ex.ORBextractor(1000, 1.2, 8, 20, 7)
    yield ex, e
    e.close()


@pytest.mark.gpu
def test_gpu_clahe_matches_cv2_golden_and_oracle(oracle, images, kat, gpu_ext):
    ex, ext = gpu_ext
    for name, img, clip, tx, ty in clahe_cases(images):
        out = ex.clahe(ext, img, clip, (tx, ty))
        check_golden(kat, name, out)
        assert np.array_equal(out, oracle.clahe(img, clip, (tx, ty))), name


@pytest.mark.gpu
def test_gpu_clahe_random_geometries(oracle, gpu_ext):
    ex, ext = gpu_ext
    rng = np.random.default_rng(8)
    for i in range(40):
        h, w = int(rng.integers(16, 260)), int(rng.integers(16, 400))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8) if i % 3 else (synth_frame(i, max(w, 32), max(h, 32))[:h, :w]).copy()
        clip, tx, ty = float(rng.choice([0.0, 0.7, 3.0, 40.0])), int(rng.integers(1, 14)), int(rng.integers(1, 14))
        assert np.array_equal(ex.clahe(ext, img, clip, (tx, ty)), oracle.clahe(img, clip, (tx, ty))), (h, w, clip, tx, ty)


@pytest.mark.gpu
def test_gpu_clahe_batch_device_and_feeds_extractor(oracle, gpu_ext):
    """A device-resident batch with padded strides equals frame-by-frame results, several launch groups included; and the
    demo chain CLAHE -> ORBextractor (src/clahe/main_show_clahe_keypoint.cpp:22, :83) equals the oracle chain."""
    import torch
    ex, ext = gpu_ext
    F, h, w = 200, 480, 640                     # 61 MB: two launch groups of the 40 MB L2 budget
    frames = np.stack([synth_frame(300 + (i % 5), w, h) for i in range(F)])
    frames[:, ::7, ::5] = (frames[:, ::7, ::5].astype(np.int32) + np.arange(F)[:, None, None]).clip(0, 255).astype(np.uint8)
    pitch = 704
    src = torch.zeros((F, h, pitch), dtype=torch.uint8, device="cuda")
    src[:, :, :w] = torch.from_numpy(frames).cuda()
    dst = torch.zeros((F, h + 3, pitch), dtype=torch.uint8, device="cuda")
    ex.clahe_raw(ext, src.data_ptr(), ex.MEM_DEVICE, F, w, h, pitch, pitch * h, 3.0, (8, 8), dst.data_ptr(), ex.MEM_DEVICE, pitch, pitch * (h + 3))
    torch.cuda.synchronize()
    got = dst[:, :h, :w].cpu().numpy()
    for f in (0, 1, 99, 130, 131, 199):
        assert np.array_equal(got[f], oracle.clahe(frames[f], 3.0, (8, 8))), f
    assert int(dst[:, :, w:].sum()) == 0 and int(dst[:, h:, :].sum()) == 0      # padding untouched
    host = ex.clahe(ext, frames[:3], 3.0, (8, 8))
    assert np.array_equal(host, got[:3])
    # CLAHE -> extractor
    ret, kps, desc = ext(got[0], None, (0, 1000))
    o = oracle.OracleExtractor(1000, 1.2, 8, 20, 7)
    oret, okps, odesc = o.extract(oracle.clahe(frames[0], 3.0, (8, 8)), (0, 1000))
    assert ret == oret and kps.tobytes() == okps.tobytes() and np.array_equal(desc, odesc)
    with pytest.raises(ex.OrbxError):
        ex.clahe(ext, frames[0], 3.0, (65, 8))
