"""Committed golden vectors (tests/golden/oracle_golden.npz, made by tests/golden/make_golden.py).
CPU leg: the oracle still reproduces them bit for bit.  GPU leg: the CUDA path reproduces them
(bit-identical where the path is +,-,*,/,sqrt only; 1e-10 where closures call tanh/log/pow)."""
import os
import sys

import numpy as np
import pytest

from common import ROOT, rel_linf

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"))
EXACT = ("dambreak32", "cap1d", "lake2d")


def _compare(got, exact_everywhere):
    assert sorted(got) == sorted(GOLD.files)
    for k in GOLD.files:
        a, b = np.asarray(got[k]), GOLD[k]
        assert a.shape == b.shape, k
        if exact_everywhere or k.startswith(EXACT) or k.endswith("_active"):
            assert np.array_equal(a, b), f"{k}: rel-Linf {rel_linf(a.astype(float), b.astype(float))}"
        else:
            assert rel_linf(a.astype(float), b.astype(float)) <= 1e-10, k


def test_oracle_reproduces_golden(oracle_lib):
    _compare(make_golden.cases(oracle_lib), exact_everywhere=True)


@pytest.mark.gpu
def test_gpu_reproduces_golden(gpu_lib):
    _compare(make_golden.cases(gpu_lib), exact_everywhere=False)
