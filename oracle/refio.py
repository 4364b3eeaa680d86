"""oracle/refio.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

File formats and subprocess helpers for the verbatim-reference binaries built by oracle/Makefile
(`oracle/_ref/ref_extract`, `oracle/_ref/ref_extract_bump`; see oracle/ref_main.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import json
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "ref_extract")
REF_BIN_BUMP = os.path.join(HERE, "_ref", "ref_extract_bump")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def have_ref(bump=True):
    return os.access(REF_BIN_BUMP if bump else REF_BIN, os.X_OK)


def write_frames(path, frames):
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    if frames.ndim == 2:
        frames = frames[None]
    n, h, w = frames.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", 0x4642524F, n, w, h))
        f.write(frames.tobytes())


def read_results(path):
    """-> list of dicts per frame: ret, kps (KP_DTYPE), desc (n,32), counts, level_kps [list], pyr [list]|None."""
    with open(path, "rb") as f:
        buf = f.read()
    magic, nframes, nlevels, dump = struct.unpack_from("<4i", buf, 0)
    assert magic == 0x5242524F
    off = 16
    out = []
    for _ in range(nframes):
        ret, n = struct.unpack_from("<2i", buf, off)
        off += 8
        counts = np.frombuffer(buf, "<i4", nlevels, off).copy()
        off += 4 * nlevels
        kps = np.frombuffer(buf, KP_DTYPE, n, off).copy()
        off += 28 * n
        desc = np.frombuffer(buf, np.uint8, 32 * n, off).reshape(n, 32).copy()
        off += 32 * n
        lv = []
        for c in counts:
            lv.append(np.frombuffer(buf, KP_DTYPE, int(c), off).copy())
            off += 28 * int(c)
        pyr = None
        if dump:
            pyr = []
            for _l in range(nlevels):
                w, h = struct.unpack_from("<2i", buf, off)
                off += 8
                plane = np.frombuffer(buf, np.uint8, (w + 38) * (h + 38), off).reshape(h + 38, w + 38).copy()
                off += (w + 38) * (h + 38)
                pyr.append(plane)
        out.append(dict(ret=ret, kps=kps, desc=desc, counts=counts, level_kps=lv, pyr=pyr))
    assert off == len(buf)
    return out


def run_reference(frames, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7, lap=(0, 0), dump_pyr=False, bump=True):
    """Run the unmodified reference on a stack of gray frames (n,h,w) u8."""
    exe = REF_BIN_BUMP if bump else REF_BIN
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.orbf"), os.path.join(td, "out.orbr")
        write_frames(fin, frames)
        cmd = [exe, "run", fin, fout, str(nfeatures), repr(float(scale)), str(nlevels), str(ini), str(mn),
               str(lap[0]), str(lap[1]), "1" if dump_pyr else "0"]
        subprocess.run(cmd, check=True)
        return read_results(fout)


REF_STEREO = os.path.join(HERE, "_ref", "ref_stereo")


def have_ref_stereo():
    return os.access(REF_STEREO, os.X_OK)


def run_reference_stereo(kps_l, desc_l, kps_r, desc_r, scale, inv_scale, pyr_l, pyr_r, mb, mbf):
    """Run the unmodified Frame::ComputeStereoMatches (oracle/_ref/ref_stereo).  -> (mvuRight, mvDepth)."""
    kps_l = np.ascontiguousarray(kps_l, KP_DTYPE); kps_r = np.ascontiguousarray(kps_r, KP_DTYPE)
    nl = len(pyr_l)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.stin"), os.path.join(td, "out.stou")
        with open(fin, "wb") as f:
            f.write(struct.pack("<4i", 0x4E495453, nl, len(kps_l), len(kps_r)))
            f.write(struct.pack("<2f", mb, mbf))
            f.write(np.ascontiguousarray(scale, "<f4").tobytes()); f.write(np.ascontiguousarray(inv_scale, "<f4").tobytes())
            f.write(kps_l.tobytes()); f.write(np.ascontiguousarray(desc_l, np.uint8).tobytes())
            f.write(kps_r.tobytes()); f.write(np.ascontiguousarray(desc_r, np.uint8).tobytes())
            for a, b in zip(pyr_l, pyr_r):
                a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
                assert a.shape == b.shape
                f.write(struct.pack("<2i", a.shape[1], a.shape[0])); f.write(a.tobytes()); f.write(b.tobytes())
        subprocess.run([REF_STEREO, fin, fout], check=True)
        buf = open(fout, "rb").read()
    magic, n = struct.unpack_from("<2i", buf, 0)
    assert magic == 0x554F5453 and n == len(kps_l)
    u = np.frombuffer(buf, "<f4", n, 8).copy()
    d = np.frombuffer(buf, "<f4", n, 8 + 4 * n).copy()
    return u, d


REF_FRAME = os.path.join(HERE, "_ref", "ref_frame")


def have_ref_frame():
    return os.access(REF_FRAME, os.X_OK)


def run_reference_frame(width, height, fx, fy, cx, cy, dist, kps1, desc1, kps2, desc2, window=100, nn_ratio=0.9,
                        check_orientation=True, prev_matched=None):
    """Run the unmodified Frame::UndistortKeyPoints / ComputeImageBounds / AssignFeaturesToGrid and
    ORBmatcher::SearchForInitialization (oracle/_ref/ref_frame).  -> dict."""
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    n1, n2 = len(kps1), len(kps2)
    d = np.zeros(5, "<f4"); d[:len(dist)] = dist
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.frin"), os.path.join(td, "out.frou")
        with open(fin, "wb") as f:
            f.write(struct.pack("<9i", 0x4E495246, width, height, n1, n2, window, 1 if check_orientation else 0, len(dist),
                                0 if prev_matched is None else 1))
            f.write(struct.pack("<5f", nn_ratio, fx, fy, cx, cy)); f.write(d.tobytes())
            f.write(kps1.tobytes()); f.write(np.ascontiguousarray(desc1, np.uint8).tobytes())
            f.write(kps2.tobytes()); f.write(np.ascontiguousarray(desc2, np.uint8).tobytes())
            if prev_matched is not None:
                f.write(np.ascontiguousarray(prev_matched, "<f4").tobytes())
        subprocess.run([REF_FRAME, fin, fout], check=True)
        buf = open(fout, "rb").read()
    return parse_frame_output(buf, n1, n2)


def parse_frame_output(buf, n1=None, n2=None):
    """Result file of oracle/ref_frame_main.cpp (and of tests/cpp/dropin_main.cpp in `frame` mode)."""
    magic, a, b = struct.unpack_from("<3i", buf, 0)
    n1 = a if n1 is None else n1
    n2 = b if n2 is None else n2
    assert magic == 0x554F5246 and a == n1 and b == n2
    off = 12
    out = {"bounds": np.frombuffer(buf, "<f4", 4, off).copy()}; off += 16
    out["keys_un1"] = np.frombuffer(buf, KP_DTYPE, n1, off).copy(); off += 28 * n1
    out["keys_un2"] = np.frombuffer(buf, KP_DTYPE, n2, off).copy(); off += 28 * n2
    for k in (1, 2):
        start = np.frombuffer(buf, "<i4", 64 * 48 + 1, off).copy(); off += 4 * (64 * 48 + 1)
        items = np.frombuffer(buf, "<i4", int(start[-1]), off).copy(); off += 4 * int(start[-1])
        out["cell_start%d" % k], out["cell_items%d" % k] = start, items
    out["nmatches"] = struct.unpack_from("<i", buf, off)[0]; off += 4
    out["matches12"] = np.frombuffer(buf, "<i4", n1, off).copy(); off += 4 * n1
    out["prev_matched"] = np.frombuffer(buf, "<f4", 2 * n1, off).copy().reshape(-1, 2)
    return out


def bench_reference(frames, threads, seconds, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7, lap=(0, 0)):
    """Time the unmodified reference (normal allocator), one extractor + one frame per thread."""
    with tempfile.TemporaryDirectory() as td:
        fin = os.path.join(td, "in.orbf")
        write_frames(fin, frames)
        cmd = [REF_BIN, "bench", fin, str(threads), repr(float(seconds)), str(nfeatures), repr(float(scale)),
               str(nlevels), str(ini), str(mn), str(lap[0]), str(lap[1])]
        r = subprocess.run(cmd, check=True, capture_output=True, text=True)
        return json.loads(r.stdout.strip().splitlines()[-1])
