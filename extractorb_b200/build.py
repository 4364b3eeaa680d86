"""extractorb_b200/build.py -- in-tree build of the CUDA library (sm_100a only, no other targets).

    python -m extractorb_b200.build            # builds extractorb_b200/libextractorb_cuda.so

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the tree.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libextractorb_cuda.so")
HOST_LIB = os.path.join(HERE, "libORBextractor.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, "..", "include", "orbx.h")]
    if not force and not _stale(LIB, srcs):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", "-o", LIB, os.path.join(CSRC, "orbx_api.cu"), "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed")
    with open(os.path.join(HERE, "ptxas_info.txt"), "w") as f:
        f.write(r.stderr)
    return LIB


def build_host(force=False, verbose=False):
    """The C++ drop-in class (include/ORBextractor.h) against the compat cv types, plus the test driver."""
    build_cuda(force=False)
    inc = os.path.join(HERE, "..", "include")
    src = os.path.join(CSRC, "ORBextractor.cpp")
    deps = [src, os.path.join(inc, "ORBextractor.h"), os.path.join(inc, "orbx_cv_compat.hpp"), os.path.join(inc, "orbx.h"), LIB]
    cxx = shutil.which("g++") or "g++"
    if force or _stale(HOST_LIB, deps):
        cmd = [cxx, "-O2", "-std=c++11", "-fPIC", "-shared", "-DORBX_FORCE_CV_COMPAT", "-I", inc, "-o", HOST_LIB, src,
               "-L", HERE, "-lextractorb_cuda", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("g++ failed (libORBextractor.so)")
    demo_src = os.path.join(HERE, "..", "tests", "cpp", "dropin_main.cpp")
    demo = os.path.join(HERE, "..", "tests", "cpp", "dropin_main")
    if os.path.exists(demo_src) and (force or _stale(demo, [demo_src, HOST_LIB])):
        cmd = [cxx, "-O2", "-std=c++11", "-pthread", "-DORBX_FORCE_CV_COMPAT", "-I", inc, "-o", demo, demo_src, "-L", HERE,
               "-lORBextractor", "-lextractorb_cuda", "-Wl,-rpath,$ORIGIN/../../extractorb_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("g++ failed (dropin_main)")
    return HOST_LIB


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv, verbose=True))
