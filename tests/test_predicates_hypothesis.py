"""Property tests (hypothesis) of the PRODUCT predicate chain compiled for the host (predicates.cuh) and of the oracle
against exact rational arithmetic.  Domain of the product chain: ANY absolute scale (the exact stage rescales its
inputs by a power of two, scales 1e-300 .. 1e+300 are drawn here) as long as the ratio largest / smallest nonzero
magnitude inside one predicate stays below 2^212 (in-circle), 2^300 (orient3d), 2^480 (orient2d); beyond that the exact
stage answers SIGN_DOMAIN (3) instead of a sign -- never a wrong sign."""
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
scale = st.sampled_from([1e-300, 1e-200, 1e-40, 1e-12, 1.0, 3.0, 1e7, 1e40, 1e150, 1e290])


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


def _scaled(p, s):
    """p * s without leaving the doubles: components that would underflow to a denormal / overflow are not drawn"""
    q = p * s
    return q if np.all(np.isfinite(q)) and np.all((q == 0) | (np.abs(q) > 1e-307)) else None


@settings(max_examples=300, deadline=None)
@given(st.lists(coord, min_size=8, max_size=8), st.sampled_from([1e-290, 1e-250, 1e250, 1e290]))
def test_extreme_scales_whole_chain(c, s):
    """orient2d / in-circle at absolute scales where the unscaled degree-2 / degree-4 polynomial over- or underflows"""
    p = _scaled(np.array(c).reshape(4, 2), s)
    if p is None:
        return
    L = _lib()
    assert _call(L.hp_orient2d, p[:3], 0) == X.orient2d(p[0], p[1], p[2])
    assert _call(L.hp_incircle, p, 0) == X.incircle(p[0], p[1], p[2], p[3])
    assert _call(L.hp_incircle, p, 3) == X.incircle(p[0], p[1], p[2], p[3])


def test_exponent_range_is_detected_not_guessed():
    """ratio beyond the exact stage's range => SIGN_DOMAIN (3); just inside => the exact sign"""
    L = _lib()
    big, tiny = 1.0, 2.0 ** -230
    p = np.array([[0.0, 0.0], [big, tiny], [2 * big, 2 * tiny], [big, big]])     # collinear triple + a far point
    assert _call(L.hp_incircle, p, 3) == 3                                          # 2^230 > 2^212
    assert _call(L.hp_orient2d, p[:3], 3) == X.orient2d(p[0], p[1], p[2]) == 0      # 2^230 < 2^480: decided exactly
    q = np.array([[0.0, 0.0], [1.0, 2.0 ** -200], [2.0, 2.0 ** -199], [1.0, 1.0]])
    assert _call(L.hp_incircle, q, 3) == X.incircle(q[0], q[1], q[2], q[3])
    for bad in (np.inf, np.nan):
        r = np.array([[0.0, 0.0], [1.0, bad], [2.0, 1.0]])
        assert _call(L.hp_orient2d, r, 3) == 3
    p3 = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [2.0 ** -310, 2.0 ** -310, 0.0]])
    assert _call(L.hp_orient3d, p3, 3) == 3
