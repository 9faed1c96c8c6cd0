"""The Python mirror of the reference's pybind API (mmesh_b200.grid.MMesh), written like the reference's own Python tests
(python/dune/mmesh/test/distance.py, doc/examples/moving.py:231-252)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_distance_integral_and_user_loop(oracle):
    from mmesh_b200 import synth
    from mmesh_b200.grid import MMesh
    m = synth.config1(n=64)
    grid = MMesh(m)
    # python/dune/mmesh/test/distance.py:24-28: the integral of the P1 distance over the unit square is 0.25
    dist = grid.distance()
    vol = np.abs(synth.cross2(m.xy[m.cells[:, 1]] - m.xy[m.cells[:, 0]], m.xy[m.cells[:, 2]] - m.xy[m.cells[:, 0]])) / 2
    integral = float(np.sum(vol * dist.array()[m.cells].mean(axis=1)))
    assert abs(integral - 0.25) < 1e-8
    assert abs(dist(7) - dist.array()[m.cells[7]].mean()) < 1e-15
    # doc/examples/moving.py:231-252: markElements -> ensure movement -> (adapt) -> move -> postAdapt
    grid.indicator().init()
    grid.indicator().maxH = 0.45 * grid.indicator().maxH        # users tune the members (test-femadaptation.cc:219-220)
    shifts = synth.shifts_uniform(m, 0.3 * m.h, 30)
    marked = grid.markElements()
    d = oracle.distance_2d(m.xy, m.segs)
    prm = oracle.Params.from_buffer_copy(bytes(grid._ctx.get_params()))
    rm, _, rc = oracle.mark(m.xy, d, m.cells, m.perm, prm, oracle.distance_max(d))
    assert [grid.getMark(c) for c in range(0, m.nc, 97)] == [int(v) for v in rm[::97]]
    assert marked == (rc[0] + rc[1] > 0 or bool(np.any(oracle.mark_interface_2d(m.xy, m.segs, prm, oracle.distance_max(d)) != 0)))
    cand = grid.adaptCandidates()                                  # mmesh.hh:826-908
    im = oracle.mark_interface_2d(m.xy, m.segs, prm, oracle.distance_max(d))
    want = oracle.adapt_candidates_2d(m.xy, m.cells, m.perm, m.nb, rm, _, m.segs, im, None)
    assert cand["insert_cells"].tolist() == want["insert_cells"].tolist() and len(want["insert_cells"]) > 0
    assert cand["remove_vertices"].tolist() == want["remove_vertices"].tolist()
    assert cand["insert_segments"].tolist() == want["insert_segs"].tolist()
    assert cand["remove_segment_vertices"].tolist() == want["remove_seg_vertices"].tolist()
    change = grid.ensureVertexMovement(shifts)
    valid, _, ninv = oracle.trial_validate(m.xy, shifts, m.cells)
    assert change == (ninv > 0 and len(grid.removed) > 0) and np.array_equal(grid.coordinates(), m.xy)
    grid.moveVertices(shifts)
    assert np.array_equal(grid.coordinates(), oracle.move(m.xy, shifts))
    grid.postAdapt()
    assert not grid.preAdapt()
    # interface movement: interface-vertex-indexed shifts
    from mmesh_b200 import capi
    niv = int((m.vflags & 1).sum())
    ish = np.tile([0.0, 1e-4 * m.h], (niv, 1))
    moved = oracle.move(m.xy, shifts)
    if oracle.trial_validate(moved, None, m.cells)[2] > 0:      # the moved mesh has inverted cells: mmesh.hh:926-930 throws
        with pytest.raises(capi.GridError):
            grid.ensureInterfaceMovement(ish)
    grid2 = MMesh(m)
    assert len(grid2.ensureInterfaceMovement(ish)) == 0
    grid2.moveInterface(ish)
    iv = np.nonzero(m.vflags & 1)[0]
    expect = m.xy.copy(); expect[iv] = expect[iv] + ish
    assert np.array_equal(grid2.coordinates(), expect)
    with pytest.raises(AssertionError):
        grid.moveVertices(shifts[:-1])                          # wrong size, as the assert at mmesh.hh:1423


def test_restore_delaunay_through_the_grid_class(oracle):
    """MMesh.restoreDelaunay: the snapshot object receives the flipped topology; the same triangulation as the oracle's sequential
    Lawson iteration; the next user loop runs on it."""
    from mmesh_b200 import synth
    from mmesh_b200.grid import MMesh
    m = synth.config1(n=48)
    s = synth.shifts_uniform(m, 0.25 * m.h, 77)
    s[np.unique(m.segs.ravel())] = 0.0
    for _ in range(60):
        valid, _, ninv = oracle.trial_validate(m.xy + s, None, m.cells)
        if ninv == 0:
            break
        s[np.unique(m.cells[valid == 0].ravel())] *= 0.5
    m.xy = np.ascontiguousarray(m.xy + s)
    cells0, nb0 = m.cells.copy(), m.nb.copy()
    grid = MMesh(m)
    res = grid.restoreDelaunay()
    assert res["converged"] == 1 and res["n_flips"] > 0
    oc, _, _ = oracle.restore_delaunay_2d(m.xy, cells0, nb0, m.segs)
    tri = lambda c: set(map(tuple, np.sort(np.asarray(c, np.int64), axis=1).tolist()))
    assert tri(m.cells) == tri(oc) and tri(m.cells) != tri(cells0)
    e_ref, nb_ref, _ = synth.build_edges(m.cells, m.nv)
    assert np.array_equal(m.edges, e_ref) and np.array_equal(m.nb, nb_ref) and np.array_equal(m.perm, synth.pack_perm(m.vid[m.cells]))
    grid.indicator().init()
    grid.markElements()
    d = oracle.distance_2d(m.xy, m.segs)
    prm = oracle.Params.from_buffer_copy(bytes(grid._ctx.get_params()))
    rm, _, _ = oracle.mark(m.xy, d, m.cells, m.perm, prm, oracle.distance_max(d))
    assert [grid.getMark(c) for c in range(m.nc)] == [int(v) for v in rm]
    assert grid.restoreDelaunay()["n_flips"] == 0
