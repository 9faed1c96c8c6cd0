"""CPU build of the PRODUCT distance and RatioIndicator arithmetic (dune-mmesh_b200/csrc/geom.cuh: the device functions
k_distance2d / k_mark2d call) against the oracle: distances bit-exact, marks and longest-edge index identical, incl. the
NaN case maxDist == 0 and degenerate triangles.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "native", "_build", "libhost_geom.so")


@pytest.fixture(scope="module")
def hg():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           "-I", os.path.join(ROOT, "dune-mmesh_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "host_geom.cc"), "-o", SO])
    return C.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_product_segment_distance_matches_oracle(hg, oracle):
    rng = np.random.default_rng(21)
    n = 60000
    p, v0, v1 = rng.uniform(0, 1, (n, 2)), rng.uniform(0, 1, (n, 2)), rng.uniform(0, 1, (n, 2))
    k = n // 6
    p[:k] = v0[:k]                                                       # query on an end point
    p[k:2 * k] = 0.5 * (v0[k:2 * k] + v1[k:2 * k])                       # on the segment
    g = rng.integers(0, 9, (k, 3, 2)) / 8.0                              # lattice: exact perpendicular feet, ties of the branches
    p[2 * k:3 * k], v0[2 * k:3 * k], v1[2 * k:3 * k] = g[:, 0], g[:, 1], g[:, 2]
    v1[3 * k:3 * k + 50] = v0[3 * k:3 * k + 50]                          # zero-length facet (0/0 normal): NaN handling
    out = np.empty(n)
    hg.hg_seg_distance_batch(_p(p), _p(v0), _p(v1), C.c_long(n), _p(out))
    ref = np.array([oracle.point_segment_distance(p[i], v0[i], v1[i]) for i in range(n)])
    assert np.array_equal(out.view(np.int64), ref.view(np.int64))


def test_product_indicator_matches_oracle(hg, oracle):
    rng = np.random.default_rng(22)
    n = 50000
    tri = rng.uniform(0, 1, (n, 3, 2))
    k = n // 5
    tri[:k, 2] = 0.5 * (tri[:k, 0] + tri[:k, 1]) + rng.uniform(-1e-9, 1e-9, (k, 2))      # needles: radius ratio fires
    tri[k:2 * k] = rng.integers(0, 5, (k, 3, 2)) / 4.0                                   # lattice: equal edge lengths, zero area
    tri[2 * k:3 * k] *= 0.01                                                             # small cells: coarsen by minH
    d = rng.uniform(0, 0.5, (n, 3))
    xy = tri.reshape(-1, 2)
    cells = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    perm = np.full(n, 0x24, np.uint8)                                                    # id order == stored order
    dist = d.reshape(-1)
    for minH, maxH, factor, prop, maxd in ((0.05, 0.75, 1.0, 1.0, 0.5), (0.02, 0.3, 2.5, 0.5, 0.5), (0.05, 0.75, 1.0, 1.0, 0.0)):
        prm = oracle.Params.default()
        prm.minH, prm.maxH, prm.factor, prm.distProportion = minH, maxH, factor, prop
        maxDist = prop * maxd                                                            # 0 => l = NaN
        rm, rle, _ = oracle.mark(xy, dist, cells, perm, prm, maxDist)
        prm8 = np.array([prm.minH, prm.maxH, prm.factor, prm.distProportion, prm.edgeRatio, prm.radiusRatio, prm.clip_eps, prm.drop_area])
        mark = np.empty(n, np.int8)
        le = np.empty(n, np.uint8)
        hg.hg_indicator_batch(_p(np.ascontiguousarray(tri.reshape(n, 6))), _p(np.ascontiguousarray(d)), C.c_long(n), _p(prm8),
                              C.c_double(maxDist), _p(mark), _p(le))
        assert np.array_equal(mark, rm) and np.array_equal(le, rle)
        if maxd > 0:
            assert len(set(rm.tolist())) == 3
