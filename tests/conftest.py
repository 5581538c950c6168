import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def phe():
    import texasholdempocker_rl_b200 as phe
    return phe


@pytest.fixture(scope="session")
def hostcheck():
    """Host instantiation of the device arithmetic (csrc/phe_hostcheck.cpp), test-only."""
    import ctypes as C
    import importlib
    b = importlib.import_module("texasholdempocker-rl_b200.build")
    L = C.CDLL(b.build_hostcheck())
    assert L.hc_init() == 0
    vp = C.c_void_p
    L.hc_get_image.argtypes = [vp, vp]
    L.hc_eval7_masks.argtypes = [vp, C.c_int64, vp]
    L.hc_deal.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, vp]
    L.hc_equity_mc.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, vp]
    L.hc_colex_unrank7.argtypes = [C.c_uint64, vp]
    L.hc_equity_plan.argtypes = [vp, C.c_int, vp, C.c_int, C.c_uint64, C.c_uint32, vp]
    L.hc_plan_num_deals.argtypes = [C.c_int, vp, C.c_int]
    L.hc_plan_num_deals.restype = C.c_uint64
    return L


@pytest.fixture(scope="session")
def reference_tree():
    """sys.path set up for the UNMODIFIED reference checkout in baseline/_ref (made from /root/reference by
    baseline/make_ref.py when it is missing; it ships to the GPU box) + the phevaluator shim.  Skips when neither exists."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_ref", os.path.join(ROOT, "baseline", "make_ref.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    if not os.path.exists(os.path.join(mk.DST, "MANIFEST.json")) and mk.make() is None:
        pytest.skip("no reference checkout (baseline/_ref absent and /root/reference not present)")
    for p in reversed(mk.paths()):
        if p not in sys.path:
            sys.path.insert(0, p)
    return mk.DST
