"""extractorb_b200 -- B200-native ORB extraction path (sm_100a CUDA behind the C-ABI of include/orbx.h).

This package is a thin ctypes view of ``libextractorb_cuda.so`` for tests, benchmarks and Python callers.
The reference's own interface is C++ (``ORB_SLAM3::ORBextractor``, reference inc/ORBextractor.h:44-111);
its drop-in lives in include/ORBextractor.h + extractorb_b200/csrc/ORBextractor.cpp.  ``ORBextractor``
below mirrors that class (same constructor arguments, member names and return conventions).

There is no CPU fallback: importing works anywhere, but constructing an extractor without the built
library or without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

__all__ = ["ORBextractor", "OrbxError", "OrbxParams", "KP_DTYPE", "load_library", "library_path", "STAGE_NAMES", "stereo_match",
           "FrameCalib", "image_bounds", "undistort_grid", "search_for_initialization", "FRAME_GRID_COLS", "FRAME_GRID_ROWS", "clahe", "extract_frame", "extract_batch_multi",
           "extract_batch_multi_raw"]

HERE = os.path.dirname(os.path.abspath(__file__))
STAGE_NAMES = ("pyramid", "fast", "octree", "blur", "describe")
MEM_HOST, MEM_DEVICE = 0, 1
FLAG_PROFILE, FLAG_NO_GRAPH, FLAG_SINGLE_STREAM, FLAG_COPY_ONLY = 1, 2, 4, 8

# cv::KeyPoint layout (28 bytes)
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


class OrbxParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("cell_size", C.c_int32),
                ("max_batch", C.c_int32), ("cand_per_cell", C.c_int32), ("flags", C.c_int32)]


class FrameCalib(C.Structure):
    """OrbxFrameCalib (include/orbx.h): mK, mDistCoef and the image bounds of a Frame."""
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("dist", C.c_float * 5),
                ("n_dist", C.c_int32), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]


FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48


class OrbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("orbx error %d: %s" % (code, msg))
        self.code = code


_lib = None


def library_path():
    return os.path.join(HERE, "libextractorb_cuda.so")


def load_library():
    """Load libextractorb_cuda.so (built in-tree by extractorb_b200.build).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise OrbxError(-9, "%s not built (run `python -m extractorb_b200.build`); there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    L.orbx_status_string.restype = C.c_char_p
    L.orbx_status_string.argtypes = [i32]
    L.orbx_last_error.restype = C.c_char_p
    L.orbx_last_error.argtypes = [vp]
    L.orbx_create.argtypes = [C.POINTER(OrbxParams), i32, C.POINTER(vp)]
    L.orbx_destroy.argtypes = [vp]
    L.orbx_destroy.restype = None
    L.orbx_get_tables.argtypes = [vp] + [vp] * 6
    L.orbx_ctor_tables.argtypes = [C.POINTER(OrbxParams)] + [vp] * 6
    L.orbx_max_keypoints.argtypes = [vp, i32, i32]
    L.orbx_extract.argtypes = [vp, vp, i32, i32, sz, i32, i32, vp, vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.orbx_extract_batch.argtypes = [vp, vp, i32, i32, i32, i32, sz, sz, i32, i32, vp, vp, i32, vp, i32, vp]
    L.orbx_extract_batch_multi.argtypes = [C.POINTER(vp), i32, vp, i32, i32, i32, sz, sz, i32, i32, vp, vp, i32, vp, vp]
    L.orbx_compute_pyramid.argtypes = [vp, vp, i32, i32, sz]
    L.orbx_compute_keypoints_octtree.argtypes = [vp]
    L.orbx_distribute_octtree.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, C.POINTER(i32)]
    L.orbx_get_level_size.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.orbx_get_pyramid_level.argtypes = [vp, i32, i32, vp, sz, i32]
    L.orbx_get_level_keypoints.argtypes = [vp, i32, i32, vp, i32, C.POINTER(i32)]
    L.orbx_get_all_level_keypoints.argtypes = [vp, i32, vp, i32, vp, C.POINTER(i32)]
    L.orbx_get_level_candidates.argtypes = [vp, i32, i32, vp, vp, vp, vp, i32, C.POINTER(i32)]
    L.orbx_get_blurred_level.argtypes = [vp, i32, i32, vp, sz]
    L.orbx_stereo_match.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, C.c_float, C.c_float, vp, vp, C.POINTER(i32)]
    L.orbx_stereo_match_batch.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, C.c_float, C.c_float, vp, vp, vp, vp]
    L.orbx_search_for_initialization_mem.argtypes = [vp, C.POINTER(FrameCalib), vp, vp, i32, vp, vp, i32, vp, vp, vp, i32, C.c_float, i32,
                                                     vp, C.POINTER(i32), i32]
    L.orbx_get_pyramid_layout.argtypes = [vp, i32, i32, C.POINTER(sz), vp, vp, vp, vp]
    L.orbx_download_pyramid.argtypes = [vp, i32, vp, sz]
    L.orbx_set_pyramid_output.argtypes = [vp, vp, sz]
    L.orbx_host_alloc.restype = vp
    L.orbx_host_alloc.argtypes = [sz, i32]
    L.orbx_host_free.restype = None
    L.orbx_host_free.argtypes = [vp]
    L.orbx_frame_image_bounds.argtypes = [vp, C.POINTER(FrameCalib), i32, i32]
    L.orbx_frame_undistort_grid.argtypes = [vp, C.POINTER(FrameCalib), vp, i32, vp, vp, vp, C.POINTER(i32)]
    L.orbx_extract_frame.argtypes = [vp, vp, i32, i32, sz, i32, i32, C.POINTER(FrameCalib), vp, vp, i32, C.POINTER(i32), C.POINTER(i32), vp, vp, vp,
                                     C.POINTER(i32)]
    L.orbx_search_for_initialization.argtypes = [vp, C.POINTER(FrameCalib), vp, vp, i32, vp, vp, i32, vp, vp, vp, i32, C.c_float, i32,
                                                 vp, C.POINTER(i32)]
    L.orbx_clahe.argtypes = [vp, vp, i32, i32, i32, i32, sz, sz, C.c_double, i32, i32, vp, i32, sz, sz, vp]
    L.orbx_last_init_fallbacks.argtypes = [vp]
    L.orbx_stage_times.argtypes = [vp, vp, C.POINTER(C.c_int64)]
    L.orbx_launch_count.restype = C.c_int64
    L.orbx_launch_count.argtypes = [vp]
    L.orbx_uses_tma.argtypes = [vp]
    L.orbx_synchronize.argtypes = [vp]
    L.orbx_get_stream.restype = vp
    L.orbx_get_stream.argtypes = [vp]
    _lib = L
    return L


class ORBextractor:
    """Mirror of ORB_SLAM3::ORBextractor (reference inc/ORBextractor.h:44-111) over the C-ABI.

    ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) as in the reference
    (src/orb_extractor/ORBextractor.cc:408-411); keyword-only extras configure the GPU side.
    """

    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, *, device=0, max_batch=1,
                 cell_size=30, cand_per_cell=0, profile=False, flags=0):
        self._L = load_library()
        self._h = C.c_void_p()
        prm = OrbxParams(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, cell_size, max_batch, cand_per_cell,
                         (FLAG_PROFILE if profile else 0) | flags)
        rc = self._L.orbx_create(C.byref(prm), device, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise OrbxError(rc, self._L.orbx_status_string(rc).decode())
        self.nfeatures, self.scaleFactor, self.nlevels = nfeatures, float(np.float32(scaleFactor)), nlevels
        self.iniThFAST, self.minThFAST = iniThFAST, minThFAST
        self.device, self.max_batch = device, max_batch
        t = [np.empty(nlevels, np.float32) for _ in range(4)]
        q = np.empty(nlevels, np.int32)
        u = np.empty(16, np.int32)
        self._check(self._L.orbx_get_tables(self._h, *[a.ctypes.data for a in t], q.ctypes.data, u.ctypes.data))
        self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2 = t
        self.mnFeaturesPerLevel, self.umax = q, u

    # -- reference accessors (inc/ORBextractor.h:63-83) --
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self.scaleFactor
    def GetScaleFactors(self): return self.mvScaleFactor.copy()
    def GetInverseScaleFactors(self): return self.mvInvScaleFactor.copy()
    def GetScaleSigmaSquares(self): return self.mvLevelSigma2.copy()
    def GetInverseScaleSigmaSquares(self): return self.mvInvLevelSigma2.copy()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.orbx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise OrbxError(rc, "%s: %s" % (self._L.orbx_status_string(rc).decode(), self._L.orbx_last_error(self._h).decode()))

    def max_keypoints(self, width, height):
        n = self._L.orbx_max_keypoints(self._h, width, height)
        if n < 0:
            self._check(n)
        return n

    # -- operator() (ORBextractor.cc:1078-1162) --
    def __call__(self, image, mask=None, vLappingArea=(0, 0)):
        """-> (ret, keypoints[KP_DTYPE], descriptors[n,32]).  `mask` is ignored like in the reference.
        An empty image returns (-1, empty, empty) as the reference does (:1083)."""
        if image is None or image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        image = np.asarray(image)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise ValueError("image must be CV_8UC1 (2-D uint8)")  # reference: assert(type == CV_8UC1), :1087
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        cap = self.max_keypoints(w, h)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        self._check(self._L.orbx_extract(self._h, image.ctypes.data, w, h, image.strides[0], int(vLappingArea[0]),
                                         int(vLappingArea[1]), kps.ctypes.data, desc.ctypes.data, cap, C.byref(n), C.byref(mono)))
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch_host(self, frames, vLappingArea=(0, 0)):
        """frames: (F,H,W) uint8 host array -> (counts[F,2]={n,mono}, kps[F,cap], desc[F,cap,32])."""
        frames = np.ascontiguousarray(frames, np.uint8)
        F, h, w = frames.shape
        cap = self.max_keypoints(w, h)
        kps = np.zeros((F, cap), KP_DTYPE)
        desc = np.zeros((F, cap, 32), np.uint8)
        counts = np.zeros((F, 2), np.int32)
        self._check(self._L.orbx_extract_batch(self._h, frames.ctypes.data, MEM_HOST, F, w, h, w, w * h, int(vLappingArea[0]),
                                               int(vLappingArea[1]), kps.ctypes.data, desc.ctypes.data, cap, counts.ctypes.data,
                                               MEM_HOST, None))
        return counts, kps, desc

    def extract_batch_raw(self, images_ptr, in_mem, n_frames, width, height, row_stride, frame_stride, lap, kps_ptr, desc_ptr,
                          cap_per_frame, counts_ptr, out_mem, stream=None):
        """Direct call of orbx_extract_batch with raw pointers (device or host)."""
        self._check(self._L.orbx_extract_batch(self._h, images_ptr, in_mem, n_frames, width, height, row_stride, frame_stride,
                                               int(lap[0]), int(lap[1]), kps_ptr, desc_ptr, cap_per_frame, counts_ptr, out_mem,
                                               stream))

    # -- stage-wise public methods of the reference (inc/ORBextractor.h:89-92) --
    def ComputePyramid(self, image):
        image = np.ascontiguousarray(image, np.uint8)
        h, w = image.shape
        self._check(self._L.orbx_compute_pyramid(self._h, image.ctypes.data, w, h, image.strides[0]))

    def ComputeKeyPointsOctTree(self):
        """-> allKeypoints: list (per level) of KP_DTYPE arrays in level coordinates, angle set."""
        self._check(self._L.orbx_compute_keypoints_octtree(self._h))
        return [self.level_keypoints(l) for l in range(self.nlevels)]

    def DistributeOctTree(self, keys, minX, maxX, minY, maxY, N, level=0):
        keys = np.ascontiguousarray(keys, KP_DTYPE)
        cap = max(N + 2, 4 * max(1, int(round((maxX - minX) / max(1, maxY - minY))))) + 8
        out = np.zeros(cap, KP_DTYPE)
        n = C.c_int(0)
        self._check(self._L.orbx_distribute_octtree(self._h, keys.ctypes.data, len(keys), minX, maxX, minY, maxY, N,
                                                    out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    # -- state of the last extraction --
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        self._check(self._L.orbx_get_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def pyramid_level(self, level, frame=0, with_border=False):
        """mvImagePyramid[level] (inc/ORBextractor.h:85); with_border adds the 19-px reflect-101 frame."""
        w, h = self.level_size(level)
        b = 19 if with_border else 0
        out = np.empty((h + 2 * b, w + 2 * b), np.uint8)
        self._check(self._L.orbx_get_pyramid_level(self._h, frame, level, out.ctypes.data, out.strides[0], int(with_border)))
        return out

    @property
    def mvImagePyramid(self):
        return [self.pyramid_level(l) for l in range(self.nlevels)]

    def GetPyramid(self):
        return self.mvImagePyramid

    def blurred_level(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        self._check(self._L.orbx_get_blurred_level(self._h, frame, level, out.ctypes.data, out.strides[0]))
        return out

    def level_keypoints(self, level, frame=0):
        n = C.c_int(0)
        self._check(self._L.orbx_get_level_keypoints(self._h, frame, level, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), KP_DTYPE)
        self._check(self._L.orbx_get_level_keypoints(self._h, frame, level, out.ctypes.data, n.value, C.byref(n)))
        return out[:n.value]

    def level_candidates(self, level, frame=0):
        """FAST candidates sorted into the reference's emission order -> (x, y, score) int32 arrays."""
        n = C.c_int(0)
        self._check(self._L.orbx_get_level_candidates(self._h, frame, level, None, None, None, None, 0, C.byref(n)))
        m = max(n.value, 1)
        xs, ys, sc = (np.zeros(m, np.int32) for _ in range(3))
        od = np.zeros(m, np.uint32)
        self._check(self._L.orbx_get_level_candidates(self._h, frame, level, xs.ctypes.data, ys.ctypes.data, sc.ctypes.data,
                                                      od.ctypes.data, n.value, C.byref(n)))
        order = np.argsort(od[:n.value], kind="stable")
        return xs[:n.value][order], ys[:n.value][order], sc[:n.value][order]

    def pyramid_layout(self, width, height):
        """-> (frame_bytes, plane_offset[l], pitch[l], level_w[l], level_h[l]) of the bordered pyramid block (orbx_get_pyramid_layout)."""
        fb = C.c_size_t(0)
        off = (C.c_size_t * self.nlevels)()
        pitch, lw, lh = ((C.c_int32 * self.nlevels)() for _ in range(3))
        self._check(self._L.orbx_get_pyramid_layout(self._h, width, height, C.byref(fb), C.cast(off, C.c_void_p), C.cast(pitch, C.c_void_p),
                                                    C.cast(lw, C.c_void_p), C.cast(lh, C.c_void_p)))
        return fb.value, list(off), list(pitch), list(lw), list(lh)

    def set_pyramid_output(self, host_ptr, frame_stride):
        """Sink for every frame's bordered pyramid block in the following extract calls (None switches it off)."""
        self._check(self._L.orbx_set_pyramid_output(self._h, host_ptr, frame_stride))

    def download_pyramid(self, frame=0):
        """The whole bordered block of resident frame `frame` in one copy -> uint8 array of frame_bytes."""
        w, h = self.level_size(0)
        fb = self.pyramid_layout(w, h)[0]
        out = np.empty(fb, np.uint8)
        self._check(self._L.orbx_download_pyramid(self._h, frame, out.ctypes.data, fb))
        return out

    def stage_times(self):
        """-> (dict stage -> ms since last call, kernel launches since last call); needs profile=True."""
        ms = np.zeros(len(STAGE_NAMES), np.float32)
        n = C.c_int64(0)
        self._check(self._L.orbx_stage_times(self._h, ms.ctypes.data, C.byref(n)))
        return dict(zip(STAGE_NAMES, ms.tolist())), n.value

    def uses_tma(self):
        return bool(self._L.orbx_uses_tma(self._h))

    def launch_count(self):
        return int(self._L.orbx_launch_count(self._h))

    def synchronize(self):
        self._check(self._L.orbx_synchronize(self._h))

    @property
    def stream(self):
        return self._L.orbx_get_stream(self._h)


def extract_batch_multi_raw(extractors, images_ptr, n_frames, width, height, row_stride, frame_stride, lap, kps_ptr, desc_ptr,
                            cap_per_frame, counts_ptr):
    """orbx_extract_batch_multi: one process, one host thread + handle per device, launch groups pulled from a shared cursor.
    `extractors`: ORBextractor instances (one per device); all pointers are HOST memory.  -> frames processed per extractor."""
    n = len(extractors)
    hs = (C.c_void_p * n)(*[e._h for e in extractors])
    per = (C.c_int32 * n)()
    e0 = extractors[0]
    e0._check(e0._L.orbx_extract_batch_multi(hs, n, images_ptr, n_frames, width, height, row_stride, frame_stride, int(lap[0]), int(lap[1]),
                                             kps_ptr, desc_ptr, cap_per_frame, counts_ptr, C.cast(per, C.c_void_p)))
    return list(per)


def extract_batch_multi(extractors, frames, vLappingArea=(0, 0)):
    """frames: (F,H,W) uint8 host array sharded over `extractors` -> (counts[F,2], kps[F,cap], desc[F,cap,32], frames per extractor)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    F, h, w = frames.shape
    cap = extractors[0].max_keypoints(w, h)
    kps = np.zeros((F, cap), KP_DTYPE)
    desc = np.zeros((F, cap, 32), np.uint8)
    counts = np.zeros((F, 2), np.int32)
    per = extract_batch_multi_raw(extractors, frames.ctypes.data, F, w, h, w, w * h, vLappingArea, kps.ctypes.data, desc.ctypes.data, cap,
                                  counts.ctypes.data)
    return counts, kps, desc, per


def stereo_match(left, right, keys_l, desc_l, keys_r, desc_r, mb, mbf):
    """Frame::ComputeStereoMatches (reference src/Frame.cc:813-990) on the GPU: `left` / `right` are the ORBextractor
    instances that just extracted the two images (their pyramids are still resident), keys/desc what they returned.
    -> (mvuRight, mvDepth, n_matched); unmatched entries are -1 like in the reference."""
    keys_l = np.ascontiguousarray(keys_l, KP_DTYPE); keys_r = np.ascontiguousarray(keys_r, KP_DTYPE)
    desc_l = np.ascontiguousarray(desc_l, np.uint8); desc_r = np.ascontiguousarray(desc_r, np.uint8)
    u = np.full(len(keys_l), -1.0, np.float32); d = np.full(len(keys_l), -1.0, np.float32)
    n = C.c_int(0)
    left._check(left._L.orbx_stereo_match(left._h, right._h, keys_l.ctypes.data, desc_l.ctypes.data, len(keys_l), keys_r.ctypes.data,
                                          desc_r.ctypes.data, len(keys_r), float(mb), float(mbf), u.ctypes.data, d.ctypes.data, C.byref(n)))
    return u, d, n.value


def stereo_match_batch_raw(left, right, n_pairs, kps_l, desc_l, counts_l, kps_r, desc_r, counts_r, cap_per_frame, mb, mbf, u_right, depth,
                           n_matched, stream=None):
    """orbx_stereo_match_batch on raw DEVICE pointers: all stereo pairs of the launch group both extractors hold resident."""
    left._check(left._L.orbx_stereo_match_batch(left._h, right._h, n_pairs, kps_l, desc_l, counts_l, kps_r, desc_r, counts_r, cap_per_frame,
                                                float(mb), float(mbf), u_right, depth, n_matched, stream))


def search_for_initialization_raw(ext, calib, k1, d1, n1, k2, d2, n2, cell_start2, cell_items2, prev, window, nn_ratio, check_orientation,
                                  matches12, mem):
    """orbx_search_for_initialization_mem on raw pointers in `mem` memory -> nmatches."""
    n = C.c_int(0)
    ext._check(ext._L.orbx_search_for_initialization_mem(ext._h, C.byref(calib), k1, d1, n1, k2, d2, n2, cell_start2, cell_items2, prev, int(window),
                                                         float(nn_ratio), 1 if check_orientation else 0, matches12, C.byref(n), mem))
    return n.value


# ---- the rows after the extractor in a Frame constructor / monocular initialisation ----
def image_bounds(ext, fx, fy, cx, cy, dist, width, height):
    """Frame::ComputeImageBounds (reference src/Frame.cc:784-812) -> FrameCalib with mnMinX .. mnMaxY filled in."""
    c = FrameCalib()
    c.fx, c.fy, c.cx, c.cy = fx, fy, cx, cy
    for i, v in enumerate(dist):
        c.dist[i] = v
    c.n_dist = len(dist)
    ext._check(ext._L.orbx_frame_image_bounds(ext._h, C.byref(c), width, height))
    return c


def undistort_grid(ext, calib, keys):
    """Frame::UndistortKeyPoints + AssignFeaturesToGrid (src/Frame.cc:748-782, :383-417) on the GPU.
    -> (mvKeysUn, cell_start[64*48+1], cell_items): mGrid[ix][iy] = cell_items[cell_start[ix*48+iy]:cell_start[ix*48+iy+1]]."""
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    n = len(keys)
    un = np.zeros(n, KP_DTYPE)
    start = np.zeros(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, np.int32)
    items = np.zeros(max(n, 1), np.int32)
    placed = C.c_int(0)
    ext._check(ext._L.orbx_frame_undistort_grid(ext._h, C.byref(calib), keys.ctypes.data, n, un.ctypes.data, start.ctypes.data,
                                                items.ctypes.data, C.byref(placed)))
    return un, start, items[:placed.value].copy()


def search_for_initialization(ext, calib, keys_un1, desc1, keys_un2, desc2, cell_start2, cell_items2, prev_matched=None,
                              window=100, nn_ratio=0.9, check_orientation=True):
    """ORBmatcher(nn_ratio, check_orientation).SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, window)
    (src/ORBmatcher.cc:705-814).  prev_matched defaults to F1's undistorted keypoint positions, as in
    Tracking::MonocularInitialization.  -> (nmatches, vnMatches12, vbPrevMatched updated)."""
    k1 = np.ascontiguousarray(keys_un1, KP_DTYPE); k2 = np.ascontiguousarray(keys_un2, KP_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8); d2 = np.ascontiguousarray(desc2, np.uint8)
    cs = np.ascontiguousarray(cell_start2, np.int32); ci = np.ascontiguousarray(cell_items2, np.int32)
    if prev_matched is None:
        prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    else:
        prev = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = C.c_int(0)
    ext._check(ext._L.orbx_search_for_initialization(ext._h, C.byref(calib), k1.ctypes.data, d1.ctypes.data, len(k1), k2.ctypes.data,
                                                     d2.ctypes.data, len(k2), cs.ctypes.data, ci.ctypes.data, prev.ctypes.data, int(window),
                                                     float(nn_ratio), 1 if check_orientation else 0, m12.ctypes.data, C.byref(n)))
    return n.value, m12[:len(k1)].copy(), prev


def clahe(ext, images, clip_limit=3.0, tiles=(8, 8)):
    """cv::createCLAHE(clip_limit, Size(tiles[0], tiles[1]))->apply on one (h, w) or a stack (n, h, w) of uint8 host images
    (the pre-processing step of the reference's demos, src/orb_extractor/main_orb_extractor.cpp:19-22), on the GPU."""
    img = np.ascontiguousarray(images, np.uint8)
    single = img.ndim == 2
    a = img[None] if single else img
    n, h, w = a.shape
    out = np.empty_like(a)
    ext._check(ext._L.orbx_clahe(ext._h, a.ctypes.data, MEM_HOST, n, w, h, w, w * h, float(clip_limit), int(tiles[0]), int(tiles[1]),
                                 out.ctypes.data, MEM_HOST, w, w * h, None))
    return out[0] if single else out


def clahe_raw(ext, src_ptr, src_mem, n, w, h, row_stride, frame_stride, clip_limit, tiles, dst_ptr, dst_mem, dst_row_stride,
              dst_frame_stride, stream=None):
    """orbx_clahe on raw pointers (device- or host-resident frames)."""
    ext._check(ext._L.orbx_clahe(ext._h, src_ptr, src_mem, n, w, h, row_stride, frame_stride, float(clip_limit), int(tiles[0]),
                                 int(tiles[1]), dst_ptr, dst_mem, dst_row_stride, dst_frame_stride, stream))


def extract_frame(ext, image, calib, lapping=(0, 1000)):
    """The monocular Frame constructor's use of one image (reference src/Frame.cc:307-347): ExtractORB + UndistortKeyPoints +
    AssignFeaturesToGrid in one call, keypoints resident on the GPU in between.
    -> (ret, mvKeys, mDescriptors, mvKeysUn, cell_start, cell_items)."""
    img = np.ascontiguousarray(image, np.uint8)
    h, w = img.shape
    cap = ext.max_keypoints(w, h)
    kps = np.zeros(cap, KP_DTYPE); un = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    start = np.zeros(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, np.int32)
    items = np.zeros(cap, np.int32)
    n, mono, placed = C.c_int(0), C.c_int(0), C.c_int(0)
    ext._check(ext._L.orbx_extract_frame(ext._h, img.ctypes.data, w, h, img.strides[0], int(lapping[0]), int(lapping[1]), C.byref(calib),
                                         kps.ctypes.data, desc.ctypes.data, cap, C.byref(n), C.byref(mono), un.ctypes.data, start.ctypes.data,
                                         items.ctypes.data, C.byref(placed)))
    return mono.value, kps[:n.value].copy(), desc[:n.value].copy(), un[:n.value].copy(), start, items[:placed.value].copy()
