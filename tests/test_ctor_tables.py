"""the constructor tables -- mvScaleFactor, mvInvScaleFactor, mvLevelSigma2,
mvInvLevelSigma2, mnFeaturesPerLevel, umax -- that ORB-SLAM3 copies out of the extractor into every Frame through
the accessors (reference inc/ORBextractor.h:63-83, consumed at src/Frame.cc:97-103).  Compared bit for bit with the
UNMODIFIED reference class (tests/golden/ctor_tables.json, written by make_golden_tables.py from oracle/_ref/ref_extract;
checked live against that binary when it is present).  The tables are host arithmetic: no GPU is needed, and the C++
drop-in's getters must be right even on a machine without one."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, GOLDEN
from common import CTOR_TABLE_CASES, parse_tables

FLOAT_TABLES = ("mvScaleFactor", "mvInvScaleFactor", "mvLevelSigma2", "mvInvLevelSigma2")


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "ctor_tables.json")) as f:
        return json.load(f)


def key(c):
    return "/".join(str(x) for x in c)


def abi_tables(nfeatures, scale, nlevels):
    import extractorb_b200 as ex
    L = ex.load_library()
    prm = ex.OrbxParams(nfeatures, scale, nlevels, 20, 7, 0, 0, 0, 0)
    t = [np.zeros(nlevels, np.uint32) for _ in range(4)]
    q, u = np.zeros(nlevels, np.int32), np.zeros(16, np.int32)
    L.orbx_ctor_tables.argtypes = [C.POINTER(ex.OrbxParams)] + [C.c_void_p] * 6
    rc = L.orbx_ctor_tables(C.byref(prm), *[a.ctypes.data for a in t], q.ctypes.data, u.ctypes.data)
    assert rc == 0
    out = dict(zip(FLOAT_TABLES, [a.tolist() for a in t]))
    out["mnFeaturesPerLevel"], out["umax"] = q.tolist(), u.tolist()
    return out


@pytest.mark.parametrize("case", CTOR_TABLE_CASES, ids=key)
def test_c_abi_tables_equal_reference(golden, case):
    """orbx_ctor_tables (device-free) against the reference constructor, every float compared by bit pattern."""
    g = golden[key(case)]
    got = abi_tables(case[0], case[1], case[2])
    for name in FLOAT_TABLES + ("mnFeaturesPerLevel", "umax"):
        assert got[name] == g[name], name


@pytest.mark.parametrize("case", CTOR_TABLE_CASES, ids=key)
def test_cpp_getters_equal_reference(golden, case):
    """GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares / GetLevels /
    GetScaleFactor and the public mnFeaturesPerLevel / umax of the C++ drop-in (tests/cpp/dropin_main.cpp `tables`)."""
    from extractorb_b200 import build
    build.build_host()
    demo = os.path.join(ROOT, "tests", "cpp", "dropin_main")
    r = subprocess.run([demo, "tables"] + [repr(x) if isinstance(x, float) else str(x) for x in case], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert parse_tables(r.stdout) == golden[key(case)]


def test_golden_equals_live_reference(golden):
    """The committed table file is what the unmodified reference prints here (skipped where oracle/_ref is absent)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_extract")
    if not os.access(ref, os.X_OK):
        pytest.skip("oracle/_ref/ref_extract not built")
    for case in CTOR_TABLE_CASES:
        r = subprocess.run([ref, "tables"] + [repr(x) if isinstance(x, float) else str(x) for x in case], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert parse_tables(r.stdout) == golden[key(case)], case


def test_ctor_tables_rejects_bad_arguments():
    import extractorb_b200 as ex
    L = ex.load_library()
    L.orbx_ctor_tables.argtypes = [C.POINTER(ex.OrbxParams)] + [C.c_void_p] * 6
    for nf, sc, nl in ((1000, 1.0, 8), (1000, 1.2, 0), (-1, 1.2, 8), (1000, 1.2, 65)):
        prm = ex.OrbxParams(nf, sc, nl, 20, 7, 0, 0, 0, 0)
        assert L.orbx_ctor_tables(C.byref(prm), None, None, None, None, None, None) == -2
    assert L.orbx_ctor_tables(None, None, None, None, None, None, None) == -2


@pytest.mark.gpu
def test_handle_tables_equal_ctor_tables(golden):
    """orbx_get_tables of a live handle == the device-free tables == the reference."""
    import extractorb_b200 as ex
    for case in CTOR_TABLE_CASES:
        if case[0] > 14000:
            continue
        e = ex.ORBextractor(*case)
        g = golden[key(case)]
        for name, arr in zip(FLOAT_TABLES, (e.mvScaleFactor, e.mvInvScaleFactor, e.mvLevelSigma2, e.mvInvLevelSigma2)):
            assert arr.view(np.uint32).tolist() == g[name], name
        assert e.mnFeaturesPerLevel.tolist() == g["mnFeaturesPerLevel"] and e.umax.tolist() == g["umax"]
        e.close()
