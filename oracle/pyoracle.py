"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of oracle/liborb_oracle.so (the plain-C restatement, oracle/orb_oracle.c + cv_prims.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .refio import KP_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liborb_oracle.so")

_u8p = C.POINTER(C.c_uint8)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)
_lib = None


def build(force=False):
    """Compile the C restatement (and, when /root/reference is present, the verbatim reference)."""
    if force:
        subprocess.run(["make", "-C", HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orb_oracle_create.restype = C.c_void_p
        L.orb_oracle_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orb_oracle_destroy.argtypes = [C.c_void_p]
        L.orb_oracle_scale_table.argtypes = [C.c_void_p, C.c_int, _fp]
        L.orb_oracle_quota.argtypes = [C.c_void_p, _ip]
        L.orb_oracle_umax.argtypes = [C.c_void_p, _ip]
        L.orb_oracle_extract.restype = C.c_int
        L.orb_oracle_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int, _ip]
        L.orb_oracle_level_size.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
        L.orb_oracle_level_plane.restype = C.c_void_p
        L.orb_oracle_level_plane.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
        L.orb_oracle_level_blur.restype = C.c_void_p
        L.orb_oracle_level_blur.argtypes = [C.c_void_p, C.c_int]
        L.orb_oracle_level_candidates.restype = C.c_int
        L.orb_oracle_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orb_oracle_level_keypoints.restype = C.c_int
        L.orb_oracle_level_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_distribute.restype = C.c_int
        L.orb_oracle_distribute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_ic_angle.restype = C.c_float
        L.orb_oracle_ic_angle.argtypes = [C.c_void_p, C.c_size_t, _ip]
        L.orb_oracle_descriptor.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]
        L.orb_oracle_stereo.restype = C.c_int
        L.orb_oracle_stereo.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orb_oracle_descriptor_distance.restype = C.c_int
        L.orb_oracle_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.orb_oracle_image_bounds.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orb_oracle_undistort_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orb_oracle_assign_grid.restype = C.c_int
        L.orb_oracle_assign_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orb_oracle_features_in_area.restype = C.c_int
        L.orb_oracle_features_in_area.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                                  C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_search_for_initialization.restype = C.c_int
        L.orb_oracle_search_for_initialization.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.ocv_clahe_u8.restype = C.c_int
        L.ocv_clahe_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.ocv_undistort_points_f32.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        L.ocv_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.ocv_copy_make_border_reflect101_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t,
                                                         C.c_int, C.c_int, C.c_int, C.c_int]
        L.ocv_fast9_16_nms.restype = C.c_int
        L.ocv_fast9_16_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ocv_fast9_16_best.restype = C.c_int
        L.ocv_fast9_16_best.argtypes = [C.c_void_p, C.c_size_t]
        L.ocv_gaussian_blur_7x7_s2_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ocv_fast_atan2.restype = C.c_float
        L.ocv_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib = L
    return _lib


# ----- primitives ---------------------------------------------------------------------------------
def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().ocv_resize_linear_u8(src.ctypes.data, src.shape[1], src.shape[0], src.strides[0], dst.ctypes.data, dw, dh, dw)
    return dst


def border101(src, b=19):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h + 2 * b, w + 2 * b), np.uint8)
    lib().ocv_copy_make_border_reflect101_u8(src.ctypes.data, w, h, src.strides[0], dst.ctypes.data, w + 2 * b, b, b, b, b)
    return dst


def fast_nms(img, th):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = ((w + 1) // 2) * ((h + 1) // 2) + 16
    xs, ys, sc = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().ocv_fast9_16_nms(img.ctypes.data, w, h, img.strides[0], th, xs.ctypes.data, ys.ctypes.data, sc.ctypes.data, cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def fast_best_map(img):
    """best(p) for every pixel >= 3 px inside the image, 0 elsewhere (int16)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h, w), np.int16)
    L = lib()
    base = img.ctypes.data
    for y in range(3, h - 3):
        for x in range(3, w - 3):
            out[y, x] = L.ocv_fast9_16_best(base + y * img.strides[0] + x, img.strides[0])
    return out


def gaussian_blur(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    dst = np.empty_like(img)
    lib().ocv_gaussian_blur_7x7_s2_u8(img.ctypes.data, w, h, img.strides[0], dst.ctypes.data, w)
    return dst


def fast_atan2(y, x):
    return float(lib().ocv_fast_atan2(float(y), float(x)))


def distribute(xs, ys, scores, min_x, max_x, min_y, max_y, n_quota):
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    scores = np.ascontiguousarray(scores, np.int32)
    cap = len(xs) + 8
    kept = np.empty(cap, np.int32)
    n = lib().orb_oracle_distribute(xs.ctypes.data, ys.ctypes.data, scores.ctypes.data, len(xs), min_x, max_x, min_y,
                                    max_y, n_quota, kept.ctypes.data, cap)
    if n < 0:
        raise ValueError("distribute failed: %d" % n)
    return kept[:n].copy()


def stereo_match(kps_l, desc_l, kps_r, desc_r, scale, inv_scale, pyr_l, pyr_r, mb, mbf):
    """Frame::ComputeStereoMatches restated (oracle/stereo_oracle.c).  pyr_l / pyr_r: lists of 2-D uint8 level images
    (without border).  -> (mvuRight, mvDepth) float32 arrays."""
    kps_l = np.ascontiguousarray(kps_l, KP_DTYPE); kps_r = np.ascontiguousarray(kps_r, KP_DTYPE)
    desc_l = np.ascontiguousarray(desc_l, np.uint8); desc_r = np.ascontiguousarray(desc_r, np.uint8)
    nl = len(pyr_l)
    pl = [np.ascontiguousarray(p, np.uint8) for p in pyr_l]
    pr = [np.ascontiguousarray(p, np.uint8) for p in pyr_r]
    lw = np.array([p.shape[1] for p in pl], np.int32); lh = np.array([p.shape[0] for p in pl], np.int32)
    ptr_l = (C.c_void_p * nl)(*[p.ctypes.data for p in pl]); ptr_r = (C.c_void_p * nl)(*[p.ctypes.data for p in pr])
    pit_l = (C.c_size_t * nl)(*[p.strides[0] for p in pl]); pit_r = (C.c_size_t * nl)(*[p.strides[0] for p in pr])
    sc = np.ascontiguousarray(scale, np.float32); isc = np.ascontiguousarray(inv_scale, np.float32)
    u = np.empty(len(kps_l), np.float32); d = np.empty(len(kps_l), np.float32)
    rc = lib().orb_oracle_stereo(len(kps_l), kps_l.ctypes.data, desc_l.ctypes.data, len(kps_r), kps_r.ctypes.data, desc_r.ctypes.data,
                                 nl, sc.ctypes.data, isc.ctypes.data, lw.ctypes.data, lh.ctypes.data, ptr_l, pit_l, ptr_r, pit_r,
                                 float(mb), float(mbf), u.ctypes.data, d.ctypes.data)
    if rc < 0:
        raise ValueError("stereo oracle: a right keypoint's row band leaves the image (UB in the reference)")
    return u, d


# ----- Frame post-processing + SearchForInitialization (oracle/frame_oracle.c) -----------------------
GRID_COLS, GRID_ROWS = 64, 48
CALIB_DTYPE = np.dtype([("fx", "<f4"), ("fy", "<f4"), ("cx", "<f4"), ("cy", "<f4"), ("dist", "<f4", (5,)), ("n_dist", "<i4"),
                        ("min_x", "<f4"), ("max_x", "<f4"), ("min_y", "<f4"), ("max_y", "<f4")])


def make_calib(fx, fy, cx, cy, dist, width, height):
    """OrbOracleCalib with the image bounds of Frame::ComputeImageBounds filled in."""
    c = np.zeros(1, CALIB_DTYPE)
    c["fx"], c["fy"], c["cx"], c["cy"] = fx, fy, cx, cy
    c["dist"][0, :len(dist)] = dist
    c["n_dist"] = len(dist)
    lib().orb_oracle_image_bounds(c.ctypes.data, width, height)
    return c


def undistort_points(xy, fx, fy, cx, cy, dist):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    d = np.ascontiguousarray(dist, np.float32)
    out = np.empty_like(xy)
    lib().ocv_undistort_points_f32(xy.ctypes.data, len(xy), fx, fy, cx, cy, d.ctypes.data, len(d), out.ctypes.data)
    return out


def undistort_keypoints(calib, kps):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.empty_like(kps)
    lib().orb_oracle_undistort_keypoints(calib.ctypes.data, kps.ctypes.data, len(kps), out.ctypes.data)
    return out


def assign_grid(calib, kps_un):
    """-> (cell_start[64*48+1], cell_items): mGrid[ix][iy] = cell_items[cell_start[ix*48+iy] : cell_start[ix*48+iy+1]]."""
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE)
    start = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
    items = np.zeros(max(len(kps_un), 1), np.int32)
    n = lib().orb_oracle_assign_grid(calib.ctypes.data, kps_un.ctypes.data, len(kps_un), start.ctypes.data, items.ctypes.data)
    return start, items[:n].copy()


def features_in_area(calib, kps_un, start, items, x, y, r, min_level=-1, max_level=-1):
    kps_un = np.ascontiguousarray(kps_un, KP_DTYPE)
    items = np.ascontiguousarray(items, np.int32)
    out = np.zeros(max(len(kps_un), 1), np.int32)
    n = lib().orb_oracle_features_in_area(calib.ctypes.data, kps_un.ctypes.data, start.ctypes.data, items.ctypes.data, x, y, r,
                                          min_level, max_level, out.ctypes.data, len(out))
    return out[:n].copy()


def search_for_initialization(calib, kps_un1, desc1, kps_un2, desc2, start2, items2, prev_matched=None, window=100, nn_ratio=0.9,
                              check_orientation=True):
    """ORBmatcher::SearchForInitialization restated.  -> (nmatches, vnMatches12, vbPrevMatched updated)."""
    k1 = np.ascontiguousarray(kps_un1, KP_DTYPE); k2 = np.ascontiguousarray(kps_un2, KP_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    items2 = np.ascontiguousarray(items2, np.int32)
    if prev_matched is None:
        prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    else:
        prev = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = lib().orb_oracle_search_for_initialization(calib.ctypes.data, k1.ctypes.data, d1.ctypes.data, len(k1), k2.ctypes.data, d2.ctypes.data,
                                                   len(k2), start2.ctypes.data, items2.ctypes.data, prev.ctypes.data, window, nn_ratio,
                                                   1 if check_orientation else 0, m12.ctypes.data)
    return n, m12[:len(k1)].copy(), prev


def clahe(img, clip_limit=3.0, tiles=(8, 8)):
    """cv::createCLAHE(clip_limit, Size(tiles[0], tiles[1]))->apply restated (oracle/cv_prims.c)."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    rc = lib().ocv_clahe_u8(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], float(clip_limit), int(tiles[0]), int(tiles[1]),
                            out.ctypes.data, out.strides[0])
    if rc != 0:
        raise ValueError("bad CLAHE arguments")
    return out


# ----- the extractor --------------------------------------------------------------------------------
class OracleExtractor:
    """Mirror of ORBextractor (reference inc/ORBextractor.h:44-111) on top of the C restatement."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, cell_w=30):
        self._L = lib()
        self._h = self._L.orb_oracle_create(nfeatures, scale_factor, nlevels, ini_th, min_th, cell_w)
        if not self._h:
            raise ValueError("bad oracle parameters")
        self.nfeatures, self.nlevels = nfeatures, nlevels
        tabs = []
        for which in range(4):
            t = np.empty(nlevels, np.float32)
            self._L.orb_oracle_scale_table(self._h, which, t.ctypes.data_as(_fp))
            tabs.append(t)
        self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2 = tabs
        q = np.empty(nlevels, np.int32)
        self._L.orb_oracle_quota(self._h, q.ctypes.data_as(_ip))
        self.mnFeaturesPerLevel = q
        u = np.empty(16, np.int32)
        self._L.orb_oracle_umax(self._h, u.ctypes.data_as(_ip))
        self.umax = u

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orb_oracle_destroy(self._h)
            self._h = None

    def extract(self, img, lap=(0, 0)):
        """-> (ret, kps[KP_DTYPE], desc[n,32])"""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.nfeatures + 8 * self.nlevels + 64
        while True:
            kps = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = C.c_int(0)
            ret = self._L.orb_oracle_extract(self._h, img.ctypes.data, w, h, img.strides[0], lap[0], lap[1],
                                             kps.ctypes.data, desc.ctypes.data, cap, C.byref(n))
            if ret < -1:
                raise ValueError("oracle: level too small for the cell grid (UB in the reference)")
            if n.value <= cap:
                return ret, kps[:n.value].copy(), desc[:n.value].copy()
            cap = n.value

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        self._L.orb_oracle_level_size(self._h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_plane(self, level):
        w, h = self.level_size(level)
        pitch = C.c_size_t()
        p = self._L.orb_oracle_level_plane(self._h, level, C.byref(pitch))
        buf = (C.c_uint8 * ((h + 38) * pitch.value)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(h + 38, pitch.value)[:, :w + 38].copy()

    def level_blur(self, level):
        w, h = self.level_size(level)
        p = self._L.orb_oracle_level_blur(self._h, level)
        if not p:
            return None
        buf = (C.c_uint8 * (h * w)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(h, w).copy()

    def level_candidates(self, level):
        n = self._L.orb_oracle_level_candidates(self._h, level, None, None, None, 0)
        xs, ys, sc = (np.empty(max(n, 1), np.int32) for _ in range(3))
        self._L.orb_oracle_level_candidates(self._h, level, xs.ctypes.data, ys.ctypes.data, sc.ctypes.data, n)
        return xs[:n], ys[:n], sc[:n]

    def level_keypoints(self, level):
        n = self._L.orb_oracle_level_keypoints(self._h, level, None, 0)
        kps = np.zeros(max(n, 1), KP_DTYPE)
        self._L.orb_oracle_level_keypoints(self._h, level, kps.ctypes.data, n)
        return kps[:n]
