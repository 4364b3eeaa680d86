"""CPU-side checks of the C-ABI boundary: the shared library loads, exports every symbol that
include/orbx.h declares, and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from extractorb_b200 import build
    build.build_cuda()
    import extractorb_b200 as ex
    return ex.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "orbx.h")).read()
    return sorted(set(re.findall(r"ORBX_API[^;(]*?\b(orbx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("orbx_create", "orbx_destroy", "orbx_extract", "orbx_extract_batch", "orbx_get_pyramid_level",
                 "orbx_get_level_keypoints", "orbx_get_tables", "orbx_distribute_octtree", "orbx_compute_pyramid",
                 "orbx_compute_keypoints_octtree"):
        assert must in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert getattr(lib, name) is not None, name


def test_status_strings(lib):
    assert lib.orbx_status_string(0) == b"ok"
    assert lib.orbx_status_string(-1) == b"empty image"
    assert b"unknown" in lib.orbx_status_string(-1234)


def test_bad_arguments_rejected_before_touching_cuda(lib):
    import extractorb_b200 as ex
    h = C.c_void_p()
    assert lib.orbx_create(None, 0, C.byref(h)) == -2
    bad = ex.OrbxParams(1000, 1.0, 8, 20, 7, 0, 0, 0, 0)       # scaleFactor must be > 1
    assert lib.orbx_create(C.byref(bad), 0, C.byref(h)) == -2
    bad = ex.OrbxParams(1000, 1.2, 0, 20, 7, 0, 0, 0, 0)       # nlevels >= 1
    assert lib.orbx_create(C.byref(bad), 0, C.byref(h)) == -2


def test_no_cpu_fallback(lib):
    """Without a CUDA device construction must fail loudly; with one it must succeed."""
    import torch
    import extractorb_b200 as ex
    if torch.cuda.is_available():
        e = ex.ORBextractor(1000, 1.2, 8, 20, 7)
        e.close()
    else:
        with pytest.raises(ex.OrbxError) as ei:
            ex.ORBextractor(1000, 1.2, 8, 20, 7)
        assert ei.value.code in (-9, -5)


def test_python_mirror_matches_reference_class_surface():
    """Names of ORB_SLAM3::ORBextractor (reference inc/ORBextractor.h:44-111) exist on the Python mirror."""
    import extractorb_b200 as ex
    for name in ("GetLevels", "GetScaleFactor", "GetScaleFactors", "GetInverseScaleFactors", "GetScaleSigmaSquares",
                 "GetInverseScaleSigmaSquares", "ComputePyramid", "ComputeKeyPointsOctTree", "DistributeOctTree",
                 "mvImagePyramid", "HARRIS_SCORE", "FAST_SCORE", "__call__"):
        assert hasattr(ex.ORBextractor, name), name


def test_cpp_headers_compile_standalone_and_link(tmp_path):
    """Each C++ header of include/ compiles on its own against the cv compat layer (C++11, as the reference's CMake asks), and a
    translation unit that names every entry point of ORBextractor.h / ORBstereo.h / ORBframe.h / ORBclahe.h links against
    libORBextractor.so + libextractorb_cuda.so (no call is made: there is no GPU here)."""
    import shutil
    import subprocess
    from extractorb_b200 import build
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    build.build_host()
    inc = os.path.join(ROOT, "include")
    for hdr in ("ORBextractor.h", "ORBExtractor.h", "ORBstereo.h", "ORBframe.h", "ORBclahe.h", "orbx.h"):
        tu = tmp_path / ("tu_%s.cpp" % hdr.replace(".", "_"))
        tu.write_text('#include "%s"\nint main() { return 0; }\n' % hdr)
        r = subprocess.run([cxx, "-std=c++11", "-Wall", "-DORBX_FORCE_CV_COMPAT", "-I", inc, "-fsyntax-only", str(tu)], capture_output=True, text=True)
        assert r.returncode == 0, (hdr, r.stderr)
    tu = tmp_path / "link.cpp"
    tu.write_text('''
#include "ORBExtractor.h"
#include "ORBstereo.h"
#include "ORBframe.h"
#include "ORBclahe.h"
int main(int argc, char**) {
    if (argc > 1000) {   // never taken: only the symbols have to resolve
        ORB_SLAM3::ORBextractor e(1000, 1.2f, 8, 20, 7);
        std::vector<cv::KeyPoint> k, u; cv::Mat d, im; std::vector<int> lap(2, 0); std::vector<float> a, b;
        e(im, cv::Mat(), k, d, lap);
        ORB_SLAM3::ComputeStereoMatches(e, e, k, d, k, d, 0.1f, 40.f, a, b);
        OrbxFrameCalib c; static ORB_SLAM3::FrameGridCells g;
        ORB_SLAM3::ComputeImageBounds(e, im, im, 640, 480, c);
        ORB_SLAM3::UndistortAndAssignToGrid(e, c, k, u, g);
        ORB_SLAM3::ExtractFrame(e, c, im, 0, 1000, k, d, u, g);
        std::vector<cv::Point2f> p; std::vector<int> m;
        ORB_SLAM3::SearchForInitialization(e, c, u, d, u, d, g, p, m, 100, 0.9f, true);
        ORB_SLAM3::ApplyCLAHE(e, im, im, 3.0, cv::Size(8, 8));
    }
    return 0;
}
''')
    exe = str(tmp_path / "link_check")
    libdir = os.path.join(ROOT, "extractorb_b200")
    r = subprocess.run([cxx, "-std=c++11", "-DORBX_FORCE_CV_COMPAT", "-I", inc, "-o", exe, str(tu), "-L", libdir, "-lORBextractor", "-lextractorb_cuda",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([exe]).returncode == 0


def test_sources_parse_against_opencv_api(tmp_path):
    """The non-compat branch of include/orbx_cv_compat.hpp (`__has_include(<opencv2/core.hpp>)`): every header of include/, the
    host side of the drop-in class and its test driver are parsed against declarations spelled like the public OpenCV core API
    (tests/cpp/opencv_stub; this image has no OpenCV C++ headers).  Catches uses of cv:: types that only the compat layer offers."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    inc, stub = os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "opencv_stub")
    probe = tmp_path / "probe.cpp"
    probe.write_text('#include "orbx_cv_compat.hpp"\n#ifndef ORBX_HAVE_OPENCV\n#error "compat layer active: the stub was not picked up"\n#endif\n'
                     '#ifndef OPENCV_CORE_HPP_STUB\n#error "stub header not included"\n#endif\nint main() { return 0; }\n')
    srcs = [str(probe), os.path.join(ROOT, "extractorb_b200", "csrc", "ORBextractor.cpp"), os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp")]
    for hdr in ("ORBextractor.h", "ORBExtractor.h", "ORBstereo.h", "ORBframe.h", "ORBclahe.h"):
        tu = tmp_path / ("cv_%s.cpp" % hdr.replace(".", "_"))
        tu.write_text('#include "%s"\nint main() { return 0; }\n' % hdr)
        srcs.append(str(tu))
    for src in srcs:
        r = subprocess.run([cxx, "-std=c++11", "-Wall", "-I", stub, "-I", inc, "-fsyntax-only", src], capture_output=True, text=True)
        assert r.returncode == 0, (src, r.stderr)
