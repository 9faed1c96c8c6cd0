"""CPU build of the PRODUCT predicate header (dune-mmesh_b200/csrc/predicates.cuh): every stage of the device
chain (static filter -> interval -> expansions) against the oracle.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "native", "_build", "libhost_predicates.so")


@pytest.fixture(scope="module")
def hp():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           "-I", os.path.join(ROOT, "dune-mmesh_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "host_predicates.cc"), "-o", SO])
    L = C.CDLL(SO)

    def run(which, stage, pts):
        pts = np.ascontiguousarray(pts, np.float64)
        out = np.empty(pts.shape[0], np.int8)
        L.hp_batch(C.c_int(which), C.c_int(stage), pts.ctypes.data_as(C.c_void_p), C.c_long(pts.shape[0]),
                   out.ctypes.data_as(C.c_void_p))
        return out

    return run


def _pts(seed, n, k, dim, grid):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, grid, size=(n, k, dim)) / float(grid)
    pert = rng.uniform(-1e-15, 1e-15, size=(n, k, dim))
    pert[: n // 3] = 0.0
    return (base + pert).reshape(n, -1)


@pytest.mark.parametrize("which,k,dim,grid,name", [(0, 3, 2, 16, "orient2d_batch"), (1, 4, 2, 8, "incircle_batch"),
                                                   (2, 4, 3, 4, "orient3d_batch")])
def test_stages_match_oracle(hp, oracle, which, k, dim, grid, name):
    pts = _pts(20 + which, 6000, k, dim, grid)
    ref = getattr(oracle, name)(pts)
    assert (ref == 0).sum() > 20
    for stage in (1, 2):     # filters may abstain (2) but never contradict
        got = hp(which, stage, pts)
        dec = got != 2
        assert np.array_equal(got[dec], ref[dec])
    assert np.array_equal(hp(which, 3, pts), ref)   # expansions alone
    assert np.array_equal(hp(which, 0, pts), ref)   # full chain


def test_collinear_lattice_triples(hp, oracle):
    # (i, i+k, i+2k) lattice triples perturbed by 1e-15: BASELINE config 4
    rng = np.random.default_rng(5)
    n = 4000
    a = rng.integers(0, 512, size=(n, 2))
    d = rng.integers(-8, 9, size=(n, 2))
    P = np.stack([a, a + d, a + 2 * d], 1) / 1024.0 + rng.uniform(-1e-15, 1e-15, size=(n, 3, 2))
    P[: n // 2] = (np.stack([a, a + d, a + 2 * d], 1) / 1024.0)[: n // 2]
    ref = oracle.orient2d_batch(P.reshape(n, 6))
    assert np.all(ref[: n // 2] == 0)
    assert np.array_equal(hp(0, 0, P.reshape(n, 6)), ref)
