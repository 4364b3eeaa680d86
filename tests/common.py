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
