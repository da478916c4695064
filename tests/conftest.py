import glob
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_TCW = "/root/reference/pyfstat/tcw_fstat_map_funcs.py"
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Win:
    """Duck-typed transientWindowRange_t built from a golden fixture."""

    def __init__(self, arr):
        self.type, self.t0, self.t0Band, self.dt0, self.tau, self.tauBand, self.dtau = (int(x) for x in arr)


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    from pyfstat_b200.atoms import AtomBatch

    batch = AtomBatch(z["atoms"], z["n_atoms"], int(z["TAtom"]))
    return z, batch, Win(z["window"])


@pytest.fixture(scope="session")
def ref_tcw():
    """The REAL reference module pyfstat/tcw_fstat_map_funcs.py, loaded standalone (its
    module-scope imports are stdlib + numpy only).  Only available in the build container."""
    if not os.path.exists(REF_TCW):
        pytest.skip("/root/reference not present (GPU box)")
    spec = importlib.util.spec_from_file_location("ref_tcw_fstat_map_funcs", REF_TCW)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def oracle():
    from oracle import tcw_oracle as O

    O.build()
    return O


# the two XLALFastNegExp table geometries on file: SURVEY A.4-1 and round 1's
EXPLUT_GEOMETRIES = [(20.0, 5120), (20.0, 2000)]
EXPLUT_DEFAULT = EXPLUT_GEOMETRIES[0]


@pytest.fixture(params=EXPLUT_GEOMETRIES, ids=lambda g: f"lut{g[1]}")
def explut(request, oracle):
    """Runs the test once per table geometry (the CPU oracle is switched; GPU tests switch their
    handle with ``gpu.set_exp_lut(*explut)``), restoring the default afterwards."""
    oracle.set_exp_lut(*request.param)
    yield request.param
    oracle.set_exp_lut(*EXPLUT_DEFAULT)


def _gpu_usable():
    """False only when the library loads and the driver reports NO CUDA device.  A library that
    is missing or does not load is not a reason to skip: the gpu tests then fail loudly."""
    try:
        from pyfstat_b200 import _lib

        return _lib.load_library().tcw_device_count() > 0
    except Exception:  # noqa: BLE001
        return True


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not errored) on a box without a usable CUDA device."""
    if any(item.get_closest_marker("gpu") for item in items) and not _gpu_usable():
        skip = pytest.mark.skip(reason="no usable CUDA device (pyfstat_b200 has no CPU fallback)")
        for item in items:
            if item.get_closest_marker("gpu"):
                item.add_marker(skip)


@pytest.fixture(scope="session")
def gpu():
    from pyfstat_b200 import _lib

    h = _lib.Handle(0)
    yield h
    h.close()


@pytest.fixture
def gpu_lut(gpu, explut):
    """The session handle switched to the parametrised table geometry (and back)."""
    gpu.set_exp_lut(*explut)
    yield gpu
    gpu.set_exp_lut(*EXPLUT_DEFAULT)
