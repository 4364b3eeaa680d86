"""Shared helpers for the test-suite: synthetic frame generators and comparison utilities."""
import numpy as np


def synth_frame(seed, w=640, h=480):
    """Corner-rich synthetic frame ('synthetic set'): three octaves of box-filtered
    uniform noise plus 40 random filled rectangles/discs, quantised to uint8.  Pure numpy, seeded."""
    rng = np.random.default_rng(20240 + seed)
    img = np.zeros((h, w), np.float32)
    for k, amp in ((4, 40.0), (8, 60.0), (16, 80.0)):
        gh, gw = h // k + 2, w // k + 2
        n = rng.random((gh, gw), dtype=np.float32)
        up = np.kron(n, np.ones((k, k), np.float32))[:h, :w]
        # cheap 3-tap box smoothing in both directions
        up = (up + np.roll(up, 1, 0) + np.roll(up, -1, 0)) / 3.0
        up = (up + np.roll(up, 1, 1) + np.roll(up, -1, 1)) / 3.0
        img += amp * up
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(40):
        cx, cy = rng.integers(0, w), rng.integers(0, h)
        sx, sy = rng.integers(8, max(9, w // 6)), rng.integers(8, max(9, h // 6))
        val = float(rng.integers(0, 256))
        if rng.random() < 0.5:
            m = (np.abs(xx - cx) < sx) & (np.abs(yy - cy) < sy)
        else:
            m = (xx - cx) ** 2 + (yy - cy) ** 2 < min(sx, sy) ** 2
        img[m] = 0.5 * img[m] + 0.5 * val
    img -= img.min()
    img *= 255.0 / max(float(img.max()), 1.0)
    return img.astype(np.uint8)


def derived_images(images):
    """Fixture images of the other BASELINE frame sizes, cut from the reference's own 640x480 robot frames (tiled
    side by side, so no decoder or resampler is involved): KITTI 1241x376 (a width that is not a multiple of 4)
    and EuRoC 752x480."""
    out = dict(images)
    out["robot866_kitti1241"] = np.ascontiguousarray(np.tile(images["robot866"], (1, 2))[52:428, :1241])
    out["robot2196_euroc752"] = np.ascontiguousarray(np.tile(images["robot2196"], (1, 2))[:, :752])
    return out


def synth_batch(n, w=640, h=480, seed0=0):
    return np.stack([synth_frame(seed0 + i, w, h) for i in range(n)])


def kp_bytes_equal(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


def desc_bit_agreement(a, b):
    """Fraction of identical descriptor bits between two (n,32) uint8 arrays."""
    if a.size == 0:
        return 1.0
    x = np.bitwise_xor(a, b)
    return 1.0 - float(np.unpackbits(x).sum()) / (a.size * 8)


def make_stereo_pair(left, seed=1):
    """Synthetic rectified right image: right(x, y) = left(x + d(y), y) with the disparity d growing from 6 px at the
    top to 40 px at the bottom (a receding ground plane), plus +-3 grey levels of seeded noise."""
    rng = np.random.default_rng(seed)
    h, w = left.shape
    right = np.empty_like(left)
    for y in range(h):
        d = 6 + (34 * y) // h
        right[y, :w - d] = left[y, d:]
        right[y, w - d:] = left[y, w - 1]
    noise = rng.integers(-3, 4, left.shape)
    return np.clip(right.astype(np.int32) + noise, 0, 255).astype(np.uint8)


STEREO_MBF = 40.0            # bf (EuRoC-like: baseline 0.11 m x fx 435 px ~ 47.9; any positive value works)
STEREO_MB = 40.0 / 435.0     # mb = mbf / fx (reference src/Frame.cc:163)


def second_view(img, angle_deg=0.0, dx=0, dy=0, seed=5, noise=4):
    """A second camera view of `img` for the monocular-initialisation tests: rotation about the image centre by
    `angle_deg` and an integer shift, sampled with nearest-neighbour look-ups (integer arithmetic after one float64
    coordinate transform), plus +-`noise` grey levels of seeded noise."""
    h, w = img.shape
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    a = np.deg2rad(angle_deg)
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    xs = np.cos(a) * (xx - cx) + np.sin(a) * (yy - cy) + cx - dx
    ys = -np.sin(a) * (xx - cx) + np.cos(a) * (yy - cy) + cy - dy
    xi = np.clip(np.floor(xs + 0.5).astype(np.int64), 0, w - 1)
    yi = np.clip(np.floor(ys + 0.5).astype(np.int64), 0, h - 1)
    out = img[yi, xi].astype(np.int32)
    rng = np.random.default_rng(seed)
    out += rng.integers(-noise, noise + 1, img.shape)
    return np.clip(out, 0, 255).astype(np.uint8)


# Camera models of the datasets BASELINE.json names (ORB-SLAM3 example settings): (fx, fy, cx, cy), distortion
FRAME_CAMERAS = {
    "tum640": (640, 480, (517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
    "euroc752": (752, 480, (458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
    "nodist640": (640, 480, (500.0, 500.0, 320.0, 240.0), (0.0, 0.0, 0.0, 0.0)),
}
# (case, camera, seed, nfeatures, rotation of the second view, dx, dy)
FRAME_CASES = [
    ("tum640", "tum640", 3, 2000, 4.0, 7, -4),
    ("euroc752", "euroc752", 11, 5000, -12.0, -15, 9),
    ("nodist640", "nodist640", 21, 1000, 0.0, 3, 2),
]


def random_frame_pair(seed, n1=1500, n2=1600, w=640, h=480):
    """Two synthetic keypoint/descriptor sets with many ambiguous matches (descriptors are noisy copies of a few
    prototypes, positions cluster): exercises re-matching (vnMatches21 take-overs), ties and the rotation filter of
    ORBmatcher::SearchForInitialization without any image."""
    rng = np.random.default_rng(seed)
    kp_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
    proto = rng.integers(0, 256, (60, 32), dtype=np.uint8)
    centres = rng.random((25, 2)) * [w - 80, h - 80] + 40

    def make(n):
        k = np.zeros(n, kp_dtype)
        c = centres[rng.integers(0, len(centres), n)]
        k["x"] = np.clip(np.round(c[:, 0] + rng.normal(0, 25, n)), 19, w - 20)
        k["y"] = np.clip(np.round(c[:, 1] + rng.normal(0, 25, n)), 19, h - 20)
        k["size"] = 31
        k["angle"] = (rng.random(n) * 360).astype(np.float32)
        k["angle"][rng.random(n) < 0.5] = 90.0
        k["response"] = rng.integers(7, 200, n)
        k["octave"] = (rng.random(n) < 0.25) * rng.integers(1, 8, n)
        k["class_id"] = -1
        d = proto[rng.integers(0, len(proto), n)].copy()
        flips = rng.integers(0, 256, (n, 12))
        for j in range(12):
            on = rng.random(n) < 0.6
            d[np.arange(n)[on], flips[on, j] // 8] ^= (1 << (flips[on, j] % 8)).astype(np.uint8)
        return k, d

    k1, d1 = make(n1)
    k2, d2 = make(n2)
    return k1, d1, k2, d2


def clahe_cases(images):
    """(name, image, clipLimit, tilesX, tilesY) for the CLAHE tests and their cv2 goldens (tests/golden/clahe_kat.npz)."""
    rng = np.random.default_rng(77)
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    low = (rng.integers(0, 40, (96, 120)) + 100).astype(np.uint8)
    tiny = rng.integers(0, 256, (33, 47), dtype=np.uint8)
    flat = np.full((64, 80), 117, np.uint8)
    return [
        ("robot866_3_8x8", images["robot866"], 3.0, 8, 8),           # the demos' call (main_orb_extractor.cpp:19-22)
        ("luna_3_8x8", images["luna"], 3.0, 8, 8),
        ("robot2196_40_8x8", images["robot2196"], 40.0, 8, 8),       # cv::createCLAHE() defaults
        ("robot866_crop_2_4x6", images["robot866"][:477, :635].copy(), 2.0, 4, 6),   # sides not multiples of the grid
        ("noise_3_8x8", noise, 3.0, 8, 8),
        ("noise_0_8x8", noise, 0.0, 8, 8),                            # no clipping
        ("low_3_8x8", low, 3.0, 8, 8),
        ("low_3_16x3", low, 3.0, 16, 3),
        ("tiny_3_8x8", tiny, 3.0, 8, 8),                              # 33x47: extended to 40x48, 6x5-pixel tiles
        ("tiny_2_4x6", tiny, 2.0, 4, 6),
        ("flat_3_8x8", flat, 3.0, 8, 8),                              # one histogram bin holds every pixel
    ]


# (nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) for the constructor-table parity test: BASELINE.json's configs,
# the demos' 5 * 1500, other scales / level counts, and degenerate quotas
CTOR_TABLE_CASES = [
    (1000, 1.2, 8, 20, 7), (1200, 1.2, 8, 20, 7), (2000, 1.2, 8, 20, 7), (8000, 1.2, 12, 20, 7), (7500, 1.2, 8, 20, 7),
    (1500, 1.2, 8, 20, 7), (500, 1.5, 5, 25, 10), (1200, 2.0, 3, 20, 7), (1000, 1.1, 10, 15, 5), (20000, 1.2, 8, 20, 7),
    (1, 1.2, 8, 20, 7), (0, 1.2, 1, 20, 7), (300, 1.3, 16, 20, 7), (1000, 1.01, 4, 20, 7),
]


def parse_tables(text):
    """Output of `ref_extract tables` / `dropin_main tables`: name followed by hex float bit patterns or ints."""
    out = {}
    for line in text.strip().splitlines():
        name, *vals = line.split()
        if name in ("mnFeaturesPerLevel", "umax", "levels"):
            out[name] = [int(v) for v in vals]
        else:
            out[name] = [int(v, 16) for v in vals]      # float bit patterns: compared exactly
    return out
