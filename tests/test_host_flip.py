"""CPU test of the PRODUCT flip logic (dune-mmesh_b200/csrc/flip.cuh, the functions the three kernels of
mmb_restore_delaunay call) compiled for the host and run phase by phase like the kernels: Delaunay restoration by edge flips
after a vertex move.  Checked against the oracle's sequential Lawson iteration (oracle/mmesh_oracle.c), against
scipy.spatial.Delaunay where no constraint is present (an independent implementation: the Delaunay triangulation of points in
general position is unique), and against the snapshot invariants (neighbour table, corner permutation, edge list)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
from mmesh_b200 import synth  # noqa: E402

SO = os.path.join(ROOT, "tests", "native", "_build", "libhost_flip.so")


def facet_keys(segs, nv):
    v_if = np.zeros(nv, np.uint8)
    if len(segs):
        v_if[np.asarray(segs).ravel()] = 1
        a, b = np.minimum(segs[:, 0], segs[:, 1]).astype(np.uint64), np.maximum(segs[:, 0], segs[:, 1]).astype(np.uint64)
        keys = np.sort((a << np.uint64(32)) | b)
    else:
        keys = np.zeros(0, np.uint64)
    return v_if, np.ascontiguousarray(keys)


@pytest.fixture(scope="module")
def host_flip():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "native", "host_flip.cc"), "-o", SO])
    L = C.CDLL(SO)
    L.hf_edges.restype = C.c_long
    L.hf_perm.restype = C.c_uint8
    P = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)

    def restore(xy, cells, nb, segs, vid=None, max_rounds=1000, perm=None):
        xy = np.ascontiguousarray(xy, np.float64)
        cells = np.ascontiguousarray(cells, np.int32).copy(); nb = np.ascontiguousarray(nb, np.int32).copy()
        nc, nv = cells.shape[0], xy.shape[0]
        v_if, keys = facet_keys(np.asarray(segs).reshape(-1, 2), nv)
        perm = np.zeros(nc, np.uint8) if perm is None else np.ascontiguousarray(perm, np.uint8).copy()
        res = np.zeros(8, np.int64); trace = np.zeros(max(max_rounds, 1), np.int32)
        vid64 = None if vid is None else np.ascontiguousarray(vid, np.int64)
        L.hf_restore(P(xy), C.c_long(nv), P(cells), P(nb), P(perm), C.c_long(nc), P(v_if), P(keys), C.c_int(keys.shape[0]), P(vid64),
                     C.c_int(max_rounds), P(res), P(trace))
        return dict(cells=cells, nb=nb, perm=perm, flips=int(res[0]), rounds=int(res[1]), converged=int(res[2]), illegal=int(res[3]),
                    constrained=int(res[4]), inverted=int(res[5]), broken=int(res[6]), trace=trace[:int(res[1])])

    def restore_active(xy, cells, nb, segs, vid=None, max_rounds=1000, perm=None):
        xy = np.ascontiguousarray(xy, np.float64)
        cells = np.ascontiguousarray(cells, np.int32).copy(); nb = np.ascontiguousarray(nb, np.int32).copy()
        nc, nv = cells.shape[0], xy.shape[0]
        v_if, keys = facet_keys(np.asarray(segs).reshape(-1, 2), nv)
        perm = np.zeros(nc, np.uint8) if perm is None else np.ascontiguousarray(perm, np.uint8).copy()
        res = np.zeros(8, np.int64)
        vid64 = None if vid is None else np.ascontiguousarray(vid, np.int64)
        L.hf_restore_active(P(xy), C.c_long(nv), P(cells), P(nb), P(perm), C.c_long(nc), P(v_if), P(keys), C.c_int(keys.shape[0]), P(vid64),
                            C.c_int(max_rounds), P(res))
        return dict(cells=cells, nb=nb, perm=perm, flips=int(res[0]), rounds=int(res[1]), converged=int(res[2]), illegal=int(res[3]),
                    constrained=int(res[4]), inverted=int(res[5]), broken=int(res[6]), tested=int(res[7]))

    restore.active = restore_active

    def restore_log(xy, cells, nb, segs):
        """(final cells, final nb, edge ids in application order, round offsets)"""
        xy = np.ascontiguousarray(xy, np.float64)
        cells = np.ascontiguousarray(cells, np.int32).copy(); nb = np.ascontiguousarray(nb, np.int32).copy()
        nc, nv = cells.shape[0], xy.shape[0]
        v_if, keys = facet_keys(np.asarray(segs).reshape(-1, 2), nv)
        res = np.zeros(8, np.int64); trace = np.zeros(1000, np.int32); log = np.zeros(4 * nc + 16, np.uint32)
        L.hf_restore_log.restype = C.c_long
        n = L.hf_restore_log(P(xy), C.c_long(nv), P(cells), P(nb), None, C.c_long(nc), P(v_if), P(keys), C.c_int(keys.shape[0]), None,
                             C.c_int(1000), P(res), P(trace), P(log), C.c_long(log.shape[0]))
        off = np.concatenate([[0], np.cumsum(trace[:int(res[1])])]).astype(np.int64)
        assert n == int(res[0]) == off[-1]
        return cells, nb, log[:n], off

    restore.log = restore_log

    def edges(cells, nb, cap):
        out = np.zeros((cap, 4), np.int32)
        n = L.hf_edges(P(np.ascontiguousarray(cells, np.int32)), P(np.ascontiguousarray(nb, np.int32)), C.c_long(cells.shape[0]), P(out), C.c_long(cap))
        return out[:n], n

    restore.edges = edges
    restore.perm = lambda a, b, c: int(L.hf_perm(C.c_int64(a), C.c_int64(b), C.c_int64(c)))
    return restore


def replay_flips(cells, nb, ids):
    """Triangulation_data_structure_2::flip(f, i) (CGAL/Triangulation_data_structure_2.h:885-915) applied to index tables, one flip
    after the other -- what a host TDS does with the log of mmb_flip_log."""
    cells = np.array(cells, np.int32); nb = np.array(nb, np.int32)
    for e in np.asarray(ids, np.int64).tolist():
        f, i = e // 3, e % 3
        n = int(nb[f, i]); ni = int(np.nonzero(nb[n] == f)[0][0])
        ccw, cw = (i + 1) % 3, (i + 2) % 3
        nccw, ncw = (ni + 1) % 3, (ni + 2) % 3
        tr, bl = int(nb[f, ccw]), int(nb[n, nccw])
        tri = int(np.nonzero(nb[tr] == f)[0][0]) if tr >= 0 else -1
        bli = int(np.nonzero(nb[bl] == n)[0][0]) if bl >= 0 else -1
        apex, opp = cells[f, i], cells[n, ni]
        cells[f, cw] = opp; cells[n, ncw] = apex
        nb[f, i] = bl
        if bl >= 0: nb[bl, bli] = f
        nb[f, ccw] = n; nb[n, nccw] = f
        nb[n, ni] = tr
        if tr >= 0: nb[tr, tri] = n
    return cells, nb


def tri_set(cells):
    return set(map(tuple, np.sort(np.asarray(cells, np.int64), axis=1).tolist()))


def moved_mesh(oracle, n, amp, seed, curves=()):
    m = synth.lattice2d(n, n, jitter=0.2, seed=seed, curves=curves, hilbert=True)
    s = synth.shifts_uniform(m, amp * m.h, seed + 100)
    if len(m.segs):                                         # keep the interface where it is: its edges must stay edges of the mesh
        s[np.unique(m.segs.ravel())] = 0.0
    for _ in range(60):                                     # what ensureVertexMovement is for: no cell may invert
        valid, _, _ = oracle.trial_validate(m.xy + s, None, m.cells)
        if valid.all():
            break
        s[np.unique(m.cells[valid == 0].ravel())] *= 0.5
    return m, m.xy + s


def check_snapshot_invariants(host_flip, m, xy, r, oracle):
    cells, nb = r["cells"], r["nb"]
    # every cell counter-clockwise (exact sign)
    pts = xy[cells].reshape(-1, 6)
    assert (oracle.orient2d_batch(pts) > 0).all()
    # neighbour table and edge list = what the snapshot builder derives from the cells
    e_ref, nb_ref, _ = synth.build_edges(cells, m.nv)
    assert np.array_equal(nb, nb_ref)
    e, n = host_flip.edges(cells, nb, e_ref.shape[0] + 8)
    assert n == e_ref.shape[0] == m.edges.shape[0] and np.array_equal(e, e_ref)
    assert np.array_equal(r["perm"], synth.pack_perm(m.vid[cells]))


@pytest.mark.parametrize("n,amp,seed", [(12, 0.2, 1), (40, 0.25, 2), (64, 0.28, 3)])
def test_flips_restore_delaunay_unconstrained(host_flip, oracle, n, amp, seed):
    from scipy.spatial import Delaunay
    m, xy = moved_mesh(oracle, n, amp, seed)
    valid, _, _ = oracle.trial_validate(xy, None, m.cells)
    assert valid.all(), "test input must be a valid triangulation"
    r = host_flip(xy, m.cells, m.nb, m.segs, vid=m.vid, perm=m.perm)
    assert r["converged"] == 1 and r["illegal"] == 0 and r["flips"] > 0 and r["inverted"] == 0 and r["broken"] == 0
    assert r["rounds"] == len(r["trace"]) and (r["trace"] > 0).all() and r["trace"].sum() == r["flips"]
    # oracle: sequential Lawson flips reach the same triangulation
    oc, onb, oflips = oracle.restore_delaunay_2d(xy, m.cells, m.nb)
    assert oflips > 0 and tri_set(oc) == tri_set(r["cells"])
    # independent implementation (Qhull); zero-area simplices along the straight domain boundary are not triangles of a mesh
    sp = Delaunay(xy).simplices
    a = xy[sp]
    area = np.abs((a[:, 1, 0] - a[:, 0, 0]) * (a[:, 2, 1] - a[:, 0, 1]) - (a[:, 2, 0] - a[:, 0, 0]) * (a[:, 1, 1] - a[:, 0, 1]))
    assert tri_set(sp[area > 1e-14]) == tri_set(r["cells"])
    check_snapshot_invariants(host_flip, m, xy, r, oracle)
    # a second call finds nothing to do
    r2 = host_flip(xy, r["cells"], r["nb"], m.segs, vid=m.vid)
    assert r2["flips"] == 0 and r2["converged"] == 1 and np.array_equal(r2["cells"], r["cells"])


def test_flips_respect_the_interface(host_flip, oracle):
    t = np.linspace(0.0, 2.0 * np.pi, 400)
    circle = (np.stack([0.5 + 0.27 * np.cos(t), 0.5 + 0.27 * np.sin(t)], 1), True)
    line = (np.stack([np.linspace(0.0, 1.0, 200), 0.13 + 0.05 * np.sin(2 * np.pi * np.linspace(0.0, 1.0, 200))], 1), False)
    m, xy = moved_mesh(oracle, 48, 0.27, 5, curves=(circle, line))
    assert len(m.segs) > 50
    valid, _, _ = oracle.trial_validate(xy, None, m.cells)
    assert valid.all()
    r = host_flip(xy, m.cells, m.nb, m.segs, vid=m.vid, perm=m.perm)
    assert r["converged"] == 1 and r["flips"] > 0
    oc, onb, oflips = oracle.restore_delaunay_2d(xy, m.cells, m.nb, m.segs)
    assert tri_set(oc) == tri_set(r["cells"])
    check_snapshot_invariants(host_flip, m, xy, r, oracle)
    # every interface segment is still an edge; the only edges failing the in-circle test are interface edges
    e, _ = host_flip.edges(r["cells"], r["nb"], m.edges.shape[0])
    ekeys = set(map(tuple, np.sort(e[:, :2], axis=1).tolist()))
    skeys = set(map(tuple, np.sort(m.segs, axis=1).tolist()))
    assert skeys <= ekeys
    flags, _ = oracle.legality(xy, None, e)
    bad = e[flags == 0]
    assert len(bad) == r["constrained"] and all(tuple(sorted(q[:2])) in skeys for q in bad.tolist())
    # without the constraint some of them flip: the constrained result differs from the plain Delaunay triangulation
    if r["constrained"]:
        r0 = host_flip(xy, m.cells, m.nb, np.zeros((0, 2), np.int32), vid=m.vid)
        assert tri_set(r0["cells"]) != tri_set(r["cells"])


def test_round_limit_and_invalid_input(host_flip, oracle):
    m, xy = moved_mesh(oracle, 32, 0.25, 9)
    full = host_flip(xy, m.cells, m.nb, m.segs)
    assert full["rounds"] >= 2
    part = host_flip(xy, m.cells, m.nb, m.segs, max_rounds=1)
    assert part["rounds"] == 1 and part["converged"] == 0 and 0 < part["flips"] < full["flips"] and part["illegal"] > 0
    # the partial result is a valid triangulation that the next call completes
    rest = host_flip(xy, part["cells"], part["nb"], m.segs)
    assert rest["converged"] == 1 and tri_set(rest["cells"]) == tri_set(full["cells"])
    # an inverted cell: nothing is flipped, the count is reported
    bad = xy.copy()
    c = m.cells[m.nc // 2]
    bad[c[0]] = bad[c[1]] + (bad[c[1]] - bad[c[0]]) + (bad[c[2]] - bad[c[1]]) * 2.0        # drag a corner across the opposite edge
    valid, _, _ = oracle.trial_validate(bad, None, m.cells)
    assert not valid.all()
    r = host_flip(bad, m.cells, m.nb, m.segs)
    assert r["inverted"] > 0 and r["flips"] == 0 and np.array_equal(r["cells"], m.cells)
    # a neighbour table without mirror entries is reported
    nb = m.nb.copy(); c0 = int(np.nonzero((nb >= 0).all(axis=1))[0][0]); nb[nb[c0, 0]][nb[nb[c0, 0]] == c0] = -1
    assert host_flip(xy, m.cells, nb, m.segs)["broken"] > 0


def test_corner_permutation_helper(host_flip):
    import itertools
    for ids in itertools.permutations((11, 42, 97)):
        assert host_flip.perm(*ids) == int(synth.pack_perm(np.array([ids]))[0])


@pytest.mark.parametrize("n,curved,max_rounds", [(40, False, 1000), (64, True, 1000), (48, True, 3), (32, False, 0)])
def test_active_cell_bookkeeping_changes_nothing(host_flip, oracle, n, curved, max_rounds):
    """mmb_restore_delaunay re-tests only the cells around the flips of the previous round (and the cells whose illegal edge had to
    wait).  Same flips, rounds, arrays and counts as sweeping every cell in every round -- with far fewer cells tested."""
    curves = ()
    if curved:
        t = np.linspace(0.0, 2.0 * np.pi, 400)
        curves = ((np.stack([0.5 + 0.27 * np.cos(t), 0.5 + 0.27 * np.sin(t)], 1), True),)
    m, xy = moved_mesh(oracle, n, 0.27, 40 + n, curves=curves)
    a = host_flip(xy, m.cells, m.nb, m.segs, vid=m.vid, perm=m.perm, max_rounds=max_rounds)
    b = host_flip.active(xy, m.cells, m.nb, m.segs, vid=m.vid, perm=m.perm, max_rounds=max_rounds)
    for k in ("flips", "rounds", "converged", "illegal", "constrained", "inverted", "broken"):
        assert a[k] == b[k], k
    assert np.array_equal(a["cells"], b["cells"]) and np.array_equal(a["nb"], b["nb"]) and np.array_equal(a["perm"], b["perm"])
    if max_rounds >= 1000:
        assert b["tested"] < 0.5 * (a["rounds"] + 1) * m.nc


@pytest.mark.parametrize("npts,seed", [(300, 1), (3000, 2), (20000, 3)])
def test_flips_on_irregular_triangulations(host_flip, oracle, npts, seed):
    """Random point clouds: the Delaunay triangulation of an affinely distorted copy (y -> 4 y + x) is a valid triangulation of the
    original points with the same hull, but far from Delaunay -- long chains of dependent flips, vertices of every degree.
    Restoring it must give Qhull's Delaunay triangulation of the original points, through the plain and the active-cell rounds."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0.0, 1.0, (npts, 2))
    warped = np.stack([xy[:, 0], 4.0 * xy[:, 1] + xy[:, 0]], 1)
    cells = Delaunay(warped).simplices.astype(np.int32)
    a = xy[cells]
    area2 = (a[:, 1, 0] - a[:, 0, 0]) * (a[:, 2, 1] - a[:, 0, 1]) - (a[:, 2, 0] - a[:, 0, 0]) * (a[:, 1, 1] - a[:, 0, 1])
    cells[area2 < 0] = cells[area2 < 0][:, [0, 2, 1]]                     # counter-clockwise, as a TDS stores them
    _, nb, _ = synth.build_edges(cells, npts)
    valid, _, _ = oracle.trial_validate(xy, None, cells)
    assert valid.all()
    want = tri_set(Delaunay(xy).simplices)
    assert tri_set(cells) != want
    r = host_flip(xy, cells, nb, np.zeros((0, 2), np.int32))
    assert r["converged"] == 1 and r["flips"] > npts // 2 and tri_set(r["cells"]) == want
    e_ref, nb_ref, _ = synth.build_edges(r["cells"], npts)
    assert np.array_equal(r["nb"], nb_ref)
    ra = host_flip.active(xy, cells, nb, np.zeros((0, 2), np.int32))
    assert (ra["flips"], ra["rounds"]) == (r["flips"], r["rounds"]) and np.array_equal(ra["cells"], r["cells"]) and np.array_equal(ra["nb"], r["nb"])
    oc, _, oflips = oracle.restore_delaunay_2d(xy, cells, nb)
    assert tri_set(oc) == want


def test_flips_on_irregular_triangulations_with_constraints(host_flip, oracle):
    """The same distorted triangulations with a random tenth of their interior edges declared interface: the constrained edges stay,
    the result equals the oracle's sequential iteration with the same constraints, every edge still failing the in-circle test is a
    constraint, and the count of those is what the pass reports."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(8)
    npts = 4000
    xy = rng.uniform(0.0, 1.0, (npts, 2))
    warped = np.stack([xy[:, 0], 4.0 * xy[:, 1] + xy[:, 0]], 1)
    cells = Delaunay(warped).simplices.astype(np.int32)
    a = xy[cells]
    area2 = (a[:, 1, 0] - a[:, 0, 0]) * (a[:, 2, 1] - a[:, 0, 1]) - (a[:, 2, 0] - a[:, 0, 0]) * (a[:, 1, 1] - a[:, 0, 1])
    cells[area2 < 0] = cells[area2 < 0][:, [0, 2, 1]]
    edges, nb, _ = synth.build_edges(cells, npts)
    inner = np.nonzero(edges[:, 3] >= 0)[0]
    segs = np.ascontiguousarray(edges[rng.choice(inner, len(inner) // 10, replace=False), :2])
    r = host_flip(xy, cells, nb, segs)
    ra = host_flip.active(xy, cells, nb, segs)
    assert r["converged"] == 1 and r["flips"] > 0 and r["constrained"] > 0
    assert (ra["flips"], ra["rounds"], ra["constrained"]) == (r["flips"], r["rounds"], r["constrained"]) and np.array_equal(ra["cells"], r["cells"])
    oc, _, _ = oracle.restore_delaunay_2d(xy, cells, nb, segs)
    assert tri_set(oc) == tri_set(r["cells"])
    e, _, _ = synth.build_edges(r["cells"], npts)
    ekeys = set(map(tuple, np.sort(e[:, :2], axis=1).tolist()))
    skeys = set(map(tuple, np.sort(segs, axis=1).tolist()))
    assert skeys <= ekeys
    flags, nill = oracle.legality(xy, None, e)
    bad = e[flags == 0]
    assert nill == r["constrained"] and all(tuple(sorted(q[:2])) in skeys for q in bad.tolist())


def test_flip_log_replays_to_the_same_tables(host_flip, oracle):
    """The log of flips (cell, slot) in round order -- what mmb_flip_log hands to the host -- replayed with TDS::flip semantics on the
    original tables gives the restored tables slot for slot; shuffling the flips INSIDE a round changes nothing (they touch disjoint
    cells), which is what allows the device to log them in any order."""
    rng = np.random.default_rng(4)
    m, xy = moved_mesh(oracle, 40, 0.27, 90)
    cells, nb, ids, off = host_flip.log(xy, m.cells, m.nb, m.segs)
    assert len(ids) > 100 and len(off) > 3
    rc, rnb = replay_flips(m.cells, m.nb, ids)
    assert np.array_equal(rc, cells) and np.array_equal(rnb, nb)
    shuffled = np.concatenate([rng.permutation(ids[off[r]:off[r + 1]]) for r in range(len(off) - 1)])
    rc2, rnb2 = replay_flips(m.cells, m.nb, shuffled)
    assert np.array_equal(rc2, cells) and np.array_equal(rnb2, nb)
