"""CPU build of the PRODUCT clip stages (dune-mmesh_b200/csrc/clip.cuh, the code the projection kernel runs) against the
oracle's reference-structured Sutherland-Hodgman (polygoncutting.hh:17-96): identical point lists, bit for bit, on
random, touching, nested, identical and epsilon-band triangle pairs.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "native", "_build", "libhost_clip.so")


@pytest.fixture(scope="module")
def hc():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           "-I", os.path.join(ROOT, "dune-mmesh_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "host_clip.cc"), "-o", SO])
    L = C.CDLL(SO)

    def run(subj, clip, eps):
        subj = np.ascontiguousarray(subj, np.float64).reshape(-1, 6)
        clip = np.ascontiguousarray(clip, np.float64).reshape(-1, 6)
        n = subj.shape[0]
        out = np.zeros((n, 18))
        cnt = np.zeros(n, np.int32)
        L.hc_clip_batch(subj.ctypes.data_as(C.c_void_p), clip.ctypes.data_as(C.c_void_p), C.c_long(n), C.c_double(eps),
                        out.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
        return out, cnt

    run.lib = L
    return run


def _cases(rng, n):
    """triangle pairs that hit every branch: generic, shared vertices / edges, nested, identical, slivers inside the
    epsilon band, collinear points, and lattice-aligned pairs like the benchmark's self + neighbour pairs"""
    subj = rng.uniform(0, 1, (n, 3, 2))
    clip = rng.uniform(0, 1, (n, 3, 2))
    k = n // 8
    clip[:k] = subj[:k]                                             # identical
    clip[k:2 * k, :2] = subj[k:2 * k, :2]                           # shared edge
    clip[2 * k:3 * k, 0] = subj[2 * k:3 * k, 1]                     # shared vertex
    c = subj[3 * k:4 * k].mean(axis=1, keepdims=True)
    clip[3 * k:4 * k] = c + 0.3 * (subj[3 * k:4 * k] - c)           # nested
    clip[4 * k:5 * k] = subj[4 * k:5 * k] + rng.uniform(-3e-14, 3e-14, (k, 3, 2))    # inside the epsilon band
    g = rng.integers(0, 8, (k, 3, 2)) / 8.0
    subj[5 * k:6 * k] = g
    clip[5 * k:6 * k] = np.roll(g, 1, axis=1) + rng.integers(-1, 2, (k, 1, 2)) / 8.0  # lattice aligned, collinear edges
    subj[6 * k:7 * k, 2] = 0.5 * (subj[6 * k:7 * k, 0] + subj[6 * k:7 * k, 1])        # degenerate (zero area) subject
    return subj.reshape(n, 6), clip.reshape(n, 6)


@pytest.mark.parametrize("eps", [1e-14, 0.0, 1e-3])
def test_product_clip_stages_match_oracle(hc, oracle, eps):
    rng = np.random.default_rng(int(eps * 1e15) + 5)
    subj, clip = _cases(rng, 40000)
    # both orientations of the clip polygon (the kernel gets TDS = ccw order; the other one exercises the empty results)
    clip2 = clip.reshape(-1, 3, 2)[:, ::-1].reshape(-1, 6)
    for cl in (clip, clip2):
        out, cnt = hc(subj, cl, eps)
        ref, rcnt = oracle.sutherland_hodgman_batch(subj, cl, eps)
        assert rcnt.max() <= 9
        assert np.array_equal(cnt, rcnt)
        assert np.array_equal(out.view(np.int64), ref.view(np.int64))          # same bits, same order
        assert len(set(cnt.tolist())) >= 4


def test_clip_stage_on_non_convex_polygons(hc, oracle):
    """Random (mostly non-convex) polygons of 3..6 points cross the clip line several times: the generic emission path of
    clip_stage (more than one crossing of a kind) must reproduce the reference loop as well."""
    rng = np.random.default_rng(77)
    L = hc.lib
    L.hc_clip_edge.restype = C.c_int
    multi = 0
    for it in range(20000):
        n = int(rng.integers(3, 7))
        poly = rng.integers(0, 9, (n, 2)).astype(np.float64) / 8.0 if it % 3 == 0 else rng.uniform(0, 1, (n, 2))
        q1, q2 = rng.uniform(0, 1, 2), rng.uniform(0, 1, 2)
        if it % 3 == 0:
            q1, q2 = np.array([0.5, 0.0]), np.array([0.5, 1.0])        # lattice points exactly on the clip line
        eps = [1e-14, 0.0, 1e-2][it % 3]
        ref, m = oracle.clip_edge(poly, q1, q2, eps)
        out = np.zeros(24)
        poly_c, q1_c, q2_c = np.ascontiguousarray(poly), np.ascontiguousarray(q1), np.ascontiguousarray(q2)
        got = L.hc_clip_edge(poly_c.ctypes.data_as(C.c_void_p), C.c_int(n), q1_c.ctypes.data_as(C.c_void_p),
                             q2_c.ctypes.data_as(C.c_void_p), C.c_double(eps), out.ctypes.data_as(C.c_void_p))
        assert got == m, (it, got, m)
        assert np.array_equal(out[:2 * m].view(np.int64), ref.reshape(-1).view(np.int64)), it
        multi += int(m > n + 1)                                        # more than two crossings happened
    assert multi > 100
