"""The assignment algorithm of csrc/assign_core.cuh, compiled for the host with one lane, against
SciPy's linear_sum_assignment (what the reference calls, models/matcher.py:83-86).  The CUDA kernel
runs the same source with 32 lanes (tests/test_gpu_detection.py)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORE = os.path.join(ROOT, "myrtle-vision_b200", "csrc", "assign_core.cuh")

HARNESS = r"""
#include "%s"
extern "C" int match_block_host(const float* cost, int nq, int nt, int ld, int* match, int match_len) {
    static mv_assign::Work w;
    int flag = 0;
    mv_assign::match_block(cost, nq, nt, ld, match, match_len, &flag, w, 0, 1);
    return flag;
}
"""


@pytest.fixture(scope="module")
def host():
    d = tempfile.mkdtemp(prefix="mv_assign_")
    src, so = os.path.join(d, "h.cpp"), os.path.join(d, "h.so")
    open(src, "w").write(HARNESS % CORE)
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


def run(host, cost, nt):
    nq, ld = cost.shape
    match = np.full(ld, -7, dtype=np.int32)
    c = np.ascontiguousarray(cost, dtype=np.float32)
    flag = host.match_block_host(c.ctypes.data_as(ctypes.c_void_p), nq, nt, ld,
                                 match.ctypes.data_as(ctypes.c_void_p), ld)
    return match, flag


@pytest.mark.parametrize("nq,nt,ld", [(100, 20, 20), (100, 1, 24), (100, 100, 100), (16, 40, 48), (1, 1, 1),
                                      (7, 5, 9), (100, 0, 8), (300, 299, 300), (3, 1000, 1000)])
def test_matches_scipy(host, nq, nt, ld):
    rng = np.random.default_rng(nq * 1000 + nt)
    for trial in range(5):
        cost = rng.standard_normal((nq, ld)).astype(np.float32) * 3
        match, flag = run(host, cost, nt)
        assert flag == 0
        assert (match[nt:] == -1).all()
        ri, ci = linear_sum_assignment(cost[:, :nt].astype(np.float64))
        want = np.full(nt, -1, dtype=np.int32)
        want[ci] = ri
        got_cost = sum(float(cost[match[t], t]) for t in range(nt) if match[t] >= 0)
        assert abs(got_cost - float(cost[ri, ci].astype(np.float64).sum())) <= 1e-9 * max(1.0, abs(got_cost))
        assert np.array_equal(match[:nt], want)


def test_detection_shaped_costs(host):
    """L1 + class + GIoU style costs (bounded, clustered) with near-duplicate predictions."""
    rng = np.random.default_rng(5)
    for trial in range(20):
        nt = int(rng.integers(1, 21))
        base = rng.uniform(-3, 3, size=(100, nt)).astype(np.float32)
        base[50:] = base[:50] + rng.uniform(0, 1e-3, size=(50, nt)).astype(np.float32)
        match, flag = run(host, base, nt)
        ri, ci = linear_sum_assignment(base.astype(np.float64))
        want = np.full(nt, -1, dtype=np.int32)
        want[ci] = ri
        assert flag == 0 and np.array_equal(match[:nt], want)


def test_infeasible_block_sets_the_flag(host):
    cost = np.full((4, 3), np.inf, dtype=np.float32)
    _, flag = run(host, cost, 3)
    assert flag == 1


def test_random_shapes_property(host):
    """Any shape, any number of valid columns, costs with many exact ties (small integers): the matching's total
    cost equals SciPy's optimum and it is a valid matching of min(nq, nt) pairs."""
    rng = np.random.default_rng(99)
    for trial in range(300):
        nq, ld = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        nt = int(rng.integers(0, ld + 1))
        cost = rng.integers(-3, 4, size=(nq, ld)).astype(np.float32)
        match, flag = run(host, cost, nt)
        assert flag == 0 and (match[nt:] == -1).all()
        used = match[:nt][match[:nt] >= 0]
        assert len(used) == min(nq, nt) and len(set(used.tolist())) == len(used) and (used < nq).all()
        if nt:
            ri, ci = linear_sum_assignment(cost[:, :nt].astype(np.float64))
            got = sum(float(cost[match[t], t]) for t in range(nt) if match[t] >= 0)
            assert got == float(cost[ri, ci].astype(np.float64).sum())
