"""The C-ABI library loads without a GPU and exports every symbol include/phe_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "phe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phe_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(phe):
    L = phe._native.lib()
    declared = header_symbols()
    assert len(declared) >= 18
    assert sorted(phe._native.SYMBOLS) == declared
    for s in declared:
        assert hasattr(L, s), s


def test_strerror_and_version(phe):
    L = phe._native.lib()
    assert L.phe_version() >= 100
    assert L.phe_strerror(0) == b"ok"
    assert b"plan" in L.phe_strerror(-4)


def test_no_cpu_fallback_without_gpu(phe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(phe._native.PheError):
        phe.evaluate_cards(51, 47, 43, 39, 35, 0, 5)
    with pytest.raises(phe._native.PheError):
        phe.eval7(np.zeros((4, 7), dtype=np.uint8) + np.arange(7, dtype=np.uint8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "texasholdempocker-rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
                assert "phe_oracle" not in src, f


def test_plan_num_deals_host_only(phe):
    from oracle import oracle as O
    for plans in (O.DEFAULT_PLANS, O.DEBUG_PLANS):
        for rnd, (plan, n) in plans.items():
            assert phe.plan_num_deals({0: 2, 1: 5, 2: 6, 3: 7}[rnd], plan) == n
    assert phe.plan_num_deals(2, [(1, 0, True)]) == 0                   # does not reach nine cards
    assert phe.plan_num_deals(7, [(2, 5000, True)]) == 0                # more draws than C(45,2): reference raises
    assert phe.plan_num_deals(5, [(2, 45, True), (1, 0, True), (1, 0, True)]) == 0   # random step must be last


def test_argument_validation_needs_no_gpu(phe):
    """Bad arguments are rejected with PHE_EINVAL / PHE_EPLAN before any CUDA call (so this runs on the CPU box)."""
    L = phe._native.lib()
    buf = np.zeros(64, dtype=np.uint64)
    p = C.c_void_p(buf.ctypes.data)
    assert L.phe_eval7(p, 7, 4, p, None) == -1                      # unknown layout
    assert L.phe_eval7(p, 0, -1, p, None) == -1                     # negative count
    assert L.phe_eval7(None, 0, 0, None, None) == 0                 # empty batch is fine
    assert L.phe_enum7(p, p, 10, 5, None) == -1                     # begin > end
    assert L.phe_enum7(p, p, 0, 133784561, None) == -1              # past C(52,7)
    assert L.phe_enum7_part(p, p, 0, 100, 3, 2, None) == -1         # part >= nparts
    assert L.phe_equity_mc(p, 4, 0, 10, 0, 1, 0, p, None) == -1     # zero opponents
    assert L.phe_equity_mc(p, 4, 9, 10, 0, 1, 0, p, None) == -1     # more than 8 opponents
    assert L.phe_showdown(p, p, None, 4, 10, p, None, None) == -1   # too many players
    assert L.phe_deal(p, 4, 22, 1, 0, 0, p, None) == -1             # more than 21 cards
    plan = np.array([[1, 0, 1]], dtype=np.int32)
    assert L.phe_equity_plan(p, 2, 1, C.c_void_p(plan.ctypes.data), 1, 0, 0, p, None) == -4   # plan does not reach 9 cards
    assert L.phe_settle(p, p, p, p, p, p, p, 4, 12, 4, 25, p, p, p, None, None) == -1         # too many seats
    assert L.phe_enum7_allreduce(p, 0, 10, 3, 2, p, None) == -1     # rank outside the world
    assert L.phe_enum7_allreduce_workspace() > 2 * 8 * 7463 * 4    # two parities x eight ranks x u32 bin slots
