"""Shared fixtures.  `-m "not gpu"` covers the oracle against golden vectors, the host logic and the
C-ABI surface; `-m gpu` tests are the parity tests proper and go through the C-ABI library."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CORPUS = ("ellip", "conical", "parab", "xos1", "monocap", "cone_nan", "dubfoc")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    os.environ.setdefault("POLYCAP_B200_QUIET", "1")
    # xraylib is absent from this image: the tests opt in to the built-in optical-constants model (oracle, reference build and
    # product then consume the same amu/scatf tables, so parity does not depend on the provider)
    os.environ.setdefault("POLYCAP_B200_OPTCONST", "builtin")


@pytest.fixture(scope="session")
def harness():
    from oracle import harness as H
    if not H.oracle_available():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    return H


@pytest.fixture(scope="session")
def orc(harness):
    return harness.Oracle()


@pytest.fixture(scope="session")
def ref(harness):
    """the reference's own sources compiled by oracle/Makefile (only where that build exists)"""
    if not harness.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return harness.Ref()


@pytest.fixture(scope="session")
def pc():
    """the product: Python host side over libpolycap_b200.so"""
    import polycap_b200
    return polycap_b200


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, "replay_%s.npz" % name)))


def optic_of(harness, g):
    return harness.Optic(int(g["nmax"]), g["z"], g["cap"], g["ext"], float(g["sig_rough"]), int(g["n_cap"]), float(g["density"]),
                         g["energies"], g["amu"], g["scatf"])


def source_of(harness, g):
    return harness.OrcSource(*[float(v) for v in g["source"]])


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) with NaN == NaN"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    d[both_nan] = 0.0
    d[(a == b)] = 0.0
    return float(np.nanmax(d)) if d.size else 0.0


def has_cuda():
    try:
        import polycap_b200
        return polycap_b200.device_count() > 0
    except Exception:
        return False
