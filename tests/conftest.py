"""pytest configuration: the `gpu` marker (tests that need a CUDA device) and shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def images():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import derived_images
    with np.load(os.path.join(GOLDEN, "images.npz")) as z:
        return derived_images({k: z[k] for k in z.files})


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()  # builds oracle/liborb_oracle.so on first use
    return pyoracle


def golden_cases():
    return sorted(f[4:-4] for f in os.listdir(GOLDEN) if f.startswith("ref_") and f.endswith(".npz"))


def load_golden(case):
    with np.load(os.path.join(GOLDEN, "ref_%s.npz" % case)) as z:
        g = {k: z[k] for k in z.files}
    nf, nl, ini, mn, lap0, lap1 = (int(v) for v in g["cfg"])
    g["params"] = dict(nfeatures=nf, scale=float(g["scale"]), nlevels=nl, ini=ini, mn=mn, lap=(lap0, lap1))
    g["image_name"] = str(g["image"])
    g["level_kps"] = [g["level_kps_%d" % l] for l in range(nl)]
    return g
