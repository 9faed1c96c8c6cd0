"""Property tests (hypothesis) of the PRODUCT predicate chain compiled for the host (predicates.cuh) and of the oracle
against exact rational arithmetic, over the stated domain: nonzero magnitudes within 1e-50 .. 1e+70 (no overflow /
underflow of the degree-4 polynomial and of its error terms)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import exactref as X

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "native", "_build", "libhost_predicates.so")


def _lib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-I", os.path.join(ROOT, "dune-mmesh_b200", "csrc"),
                               os.path.join(ROOT, "tests", "native", "host_predicates.cc"), "-o", SO])
    return C.CDLL(SO)


def _call(fn, pts, stage=0):
    a = np.ascontiguousarray(pts, np.float64)
    return fn(a.ctypes.data_as(C.c_void_p), C.c_int(stage))


# points built as base + k * ulp-ish step so that near-degenerate configurations are common
coord = st.one_of(
    st.integers(-64, 64).map(lambda k: k / 64.0),
    st.tuples(st.integers(-8, 8), st.integers(-4, 4)).map(lambda t: t[0] / 8.0 + t[1] * 2.0 ** -52),
    st.floats(-1.0, 1.0, allow_nan=False, width=64).filter(lambda v: v == 0.0 or abs(v) >= 1e-9),   # stated domain
)
scale = st.sampled_from([1e-40, 1e-12, 1.0, 3.0, 1e7, 1e40])


@settings(max_examples=400, deadline=None)
@given(st.lists(coord, min_size=6, max_size=6), scale)
def test_orient2d_chain(c, s):
    p = np.array(c).reshape(3, 2) * s
    L = _lib()
    e = X.orient2d(p[0], p[1], p[2])
    assert _call(L.hp_orient2d, p, 0) == e
    assert _call(L.hp_orient2d, p, 3) == e
    for stage in (1, 2):
        r = _call(L.hp_orient2d, p, stage)
        assert r == 2 or r == e


@settings(max_examples=400, deadline=None)
@given(st.lists(coord, min_size=8, max_size=8), scale)
def test_incircle_chain(c, s):
    p = np.array(c).reshape(4, 2) * s
    L = _lib()
    e = X.incircle(p[0], p[1], p[2], p[3])
    assert _call(L.hp_incircle, p, 0) == e
    assert _call(L.hp_incircle, p, 3) == e
    for stage in (1, 2):
        r = _call(L.hp_incircle, p, stage)
        assert r == 2 or r == e


@settings(max_examples=300, deadline=None)
@given(st.lists(coord, min_size=12, max_size=12), scale)
def test_orient3d_chain(c, s):
    p = np.array(c).reshape(4, 3) * s
    L = _lib()
    e = X.orient3d(p[0], p[1], p[2], p[3])
    assert _call(L.hp_orient3d, p, 0) == e
    assert _call(L.hp_orient3d, p, 3) == e
    for stage in (1, 2):
        r = _call(L.hp_orient3d, p, stage)
        assert r == 2 or r == e


@settings(max_examples=300, deadline=None)
@given(st.lists(coord, min_size=8, max_size=8), scale)
def test_oracle_matches_fractions(oracle, c, s):
    p = np.array(c).reshape(4, 2) * s
    assert oracle.orient2d(p[0], p[1], p[2]) == X.orient2d(p[0], p[1], p[2])
    assert oracle.incircle(p[0], p[1], p[2], p[3]) == X.incircle(p[0], p[1], p[2], p[3])
