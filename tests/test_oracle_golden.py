"""Pins the oracle's geometry/indicator/clipping restatement to the values and tolerances the reference's own
tests assert (tests/golden/reference_goldens.json, transcribed by tests/golden/make_goldens.py)."""
import json
import os

import numpy as np
import pytest

import exactref as X

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_goldens.json")))


def _perm_for(ids):
    from mmesh_b200.synth import pack_perm
    return pack_perm(np.asarray([ids]))[0]


def test_simple2d_golden(oracle):
    g = G["simple2d"]
    xy = np.array(g["vertices"], float)
    tris = np.array(g["triangles"], np.int32)
    for e, exp in g["elements"].items():
        cell = tris[int(e)]
        perm = _perm_for(cell)           # persistent id == vertex index in this fixture
        ctr, vol, ec, cc = oracle.cell_geometry_2d(xy, cell, perm)
        sub = [int(cell[(perm >> (2 * k)) & 3]) for k in range(3)]
        assert sub == exp["sub_indices"]                                   # test-mmesh-2d.cc:148-151,208-211
        assert np.allclose(ctr, exp["center"], rtol=0, atol=1e-15)         # :143,203 (printed to 17 digits)
        assert vol == exp["volume"]                                        # :145,205
        assert np.allclose(ec, exp["edge_centers"], rtol=0, atol=1e-15)    # :158-162,218-223
        # circumcenter is equidistant from the three corners
        d = [np.hypot(*(cc - xy[v])) for v in sub]
        assert max(d) - min(d) < 1e-14
    # all four triangles are positively oriented in the order listed and their volumes sum to 1 (test-grid.cc)
    valid, sign, ninv = oracle.trial_validate(xy, None, tris)
    assert ninv == 0 and np.all(sign == 1)
    vols = [oracle.cell_geometry_2d(xy, t, _perm_for(t))[1] for t in tris]
    assert abs(sum(vols) - 1.0) < 1e-15


def test_twocell3d_golden(oracle):
    g = G["twocell3d"]
    xyz = np.array(g["vertices"], float)
    tets = np.array(g["tets"], np.int32)
    for e, exp in g["elements"].items():
        cell = tets[int(e)]
        perm = _perm_for(cell)
        ctr, vol, cc = oracle.cell_geometry_3d(xyz, cell, perm)
        assert np.allclose(ctr, exp["center"], rtol=0, atol=1e-15)         # test-mmesh-3d.cc:137,191
        assert abs(vol - exp["volume"]) < 1e-16                            # :138,192
        d = [np.linalg.norm(cc - xyz[v]) for v in cell]
        assert max(d) - min(d) < 1e-14
    # 3-D validity: both orientations
    v, s, n = oracle.trial_validate(xyz, None, tets, dim=3)
    v2, s2, n2 = oracle.trial_validate(xyz, None, tets[:, [0, 1, 3, 2]], dim=3)
    assert np.array_equal(s, -s2) and np.all(s != 0) and n + n2 == 2


def test_distance_to_straight_interface(oracle):
    """test-distance.cc:18-37: dist(v) = |y - 0.5| to rel 1e-8; python/dune/mmesh/test/distance.py: integral 0.25."""
    from mmesh_b200 import synth
    m = synth.config1(n=40)
    d = oracle.distance_2d(m.xy, m.segs)
    exact = np.abs(m.xy[:, 1] - 0.5)
    nz = exact > 0
    assert np.max(np.abs(d[nz] - exact[nz]) / exact[nz]) <= G["tolerances"]["distance_2d_rel"]["value"]
    assert np.all(d[~nz] == 0)
    assert oracle.distance_max(d) == d.max()
    # P1 integral of the distance = sum_K |K| mean(d) = 0.25
    vols = np.abs(synth.cross2(m.xy[m.cells[:, 1]] - m.xy[m.cells[:, 0]], m.xy[m.cells[:, 2]] - m.xy[m.cells[:, 0]])) / 2
    integ = np.sum(vols * d[m.cells].mean(axis=1))
    assert abs(integ - 0.25) < G["tolerances"]["distance_integral"]["tol"]


def test_distance_3d_estimate(oracle):
    """test-distance.cc:63: the 3-D estimate is within 5 % of |z - 0.5| for a facet-resolved plane (away from it)."""
    from mmesh_b200 import synth
    m = synth.lattice3d(8, jitter=0.0)
    d = oracle.distance_3d(m.xyz, m.tris)
    exact = np.abs(m.xyz[:, 2] - 0.5)
    far = exact > 0
    assert np.max(np.abs(d[far] - exact[far]) / exact[far]) <= G["tolerances"]["distance_3d_rel"]["value"]


def test_clip_area_vs_exact(oracle):
    """test-intersectionvolume.cc:27-61: |SH area - exact area| <= 1e-10 on random ccw triangle pairs."""
    rng = np.random.default_rng(42)
    worst = 0.0
    for it in range(400):
        p1 = rng.uniform(0, 1, size=(3, 2))
        p2 = rng.uniform(0, 1, size=(3, 2)) if it % 2 == 0 else np.vstack([p1[:2], rng.uniform(0, 1, size=(1, 2))])
        if oracle.polygon_area(p1) < 0:
            p1 = p1[::-1].copy()
        if oracle.polygon_area(p2) < 0:
            p2 = p2[::-1].copy()
        poly = oracle.sutherland_hodgman(p1, p2)
        a = oracle.polygon_area(poly) if len(poly) else 0.0
        exact = float(X.clip_area_exact(p1, p2))
        worst = max(worst, abs(a - exact))
        assert abs(oracle.intersection_volume(p1, p2) - abs(a)) == 0.0
    assert worst <= G["tolerances"]["clip_area_abs"]["value"]


def test_conservation_and_exactness(oracle):
    """movement.cc:89-94 (sum of old∩new areas = |new| to 1e-8) and test-femadaptation.cc:286-287 (projection
    reproduces polynomials of the ansatz degree): constants for P0, linears for DG-P1; mass to 1e-13."""
    from mmesh_b200 import synth
    from projection_cases import diagonal_flip_case
    case = diagonal_flip_case(n=24, seed=7)
    prm = oracle.Params.default()
    # P0 constant + mass
    u_old = np.full(case["n_old"], 2.5)
    u_new, mass, cov, npc = oracle.project(case["xy"], case["cells"], case["perm"], case["new_off"], case["new_cell"],
                                           case["old_xy"], case["old_idx"], u_old, case["cells"].shape[0], 1, prm)
    assert np.max(np.abs(u_new - 2.5)) < 1e-12
    assert abs(cov - 1.0) < 1e-8 and abs(mass - 2.5) < 2.5e-13
    # per new cell: sum of intersection volumes = |K|
    vols = np.array([oracle.intersection_volume(case["old_xy"][p], case["new6"][i])
                     for i in range(case["new_cell"].shape[0]) for p in range(case["new_off"][i], case["new_off"][i + 1])])
    per_cell = np.add.reduceat(vols, case["new_off"][:-1])
    assert np.max(np.abs(per_cell - case["new_vol"])) < G["tolerances"]["conservation_abs"]["value"]
    # P0 non-constant: mass conserved
    u_old = case["old_centroid"][:, 0] ** 2 + case["old_centroid"][:, 1]
    mass_old = float(np.sum(case["old_vol"] * u_old))
    _, mass, _, _ = oracle.project(case["xy"], case["cells"], case["perm"], case["new_off"], case["new_cell"],
                                   case["old_xy"], case["old_idx"], u_old, case["cells"].shape[0], 1, prm)
    assert abs(mass - mass_old) <= 1e-13 * abs(mass_old)
    # DG-P1: linear function reproduced nodally, mass conserved
    f = lambda p: 0.3 + 1.7 * p[..., 0] - 0.9 * p[..., 1]
    dofs_old = f(case["old_xy"].reshape(-1, 3, 2)[case["first_pair_of_old"]])          # nodal values on cached corners
    u_new, mass, _, _ = oracle.project(case["xy"], case["cells"], case["perm"], case["new_off"], case["new_cell"],
                                       case["old_xy"], case["old_idx"], dofs_old, case["cells"].shape[0], 3, prm)
    expect = f(case["new_corners_idorder"])
    assert np.max(np.abs(u_new - expect)) < 1e-12
    mass_old = float(np.sum(case["old_vol"] * dofs_old.mean(axis=1)))
    assert abs(mass - mass_old) <= 1e-13 * abs(mass_old)


def test_indicator_canonical_params(oracle):
    """RatioIndicator with the parameter set of test-femadaptation.cc:279-281: marks are -1/0/+1, coarsening has
    priority, and the longest edge is the first strict maximum (longestedgerefinement.hh:49)."""
    from mmesh_b200 import synth
    m = synth.config1(n=40)
    p = G["indicator_params_femadaptation"]
    prm = oracle.Params.default(minH=p["minH"], maxH=p["maxH"], factor=p["factor"])
    d = oracle.distance_2d(m.xy, m.segs)
    for step in range(3):
        marks, le, (nref, ncoa) = oracle.mark(m.xy, d, m.cells, m.perm, prm, prm.distProportion * oracle.distance_max(d))
        assert set(np.unique(marks)).issubset({-1, 0, 1})
        assert nref == (marks == 1).sum() and ncoa == (marks == -1).sum()
        # recompute the longest edge independently
        k = np.stack([np.take_along_axis(m.cells, ((m.perm >> (2 * i)) & 3)[:, None].astype(np.int64), 1)[:, 0] for i in range(3)], 1)
        c = m.xy[k]
        L = np.stack([np.hypot(*(c[:, 1] - c[:, 0]).T), np.hypot(*(c[:, 2] - c[:, 0]).T), np.hypot(*(c[:, 2] - c[:, 1]).T)], 1)
        assert np.array_equal(le, np.argmax(L, axis=1))
        prm.maxH *= p["scale_maxH"]
        prm.minH *= p["scale_minH"]
    # with minH tiny and maxH tiny every cell wants refinement unless a shape criterion asks for coarsening
    prm = oracle.Params.default(minH=1e-9, maxH=1e-6, factor=1.0)
    marks, _, _ = oracle.mark(m.xy, d, m.cells, m.perm, prm, oracle.distance_max(d))
    assert set(np.unique(marks)).issubset({-1, 1}) and (marks == 1).sum() > 0.9 * m.nc
    # NaN semantics: maxDist == 0 => l = NaN => only the shape criteria can fire
    marks0, _, _ = oracle.mark(m.xy, d * 0, m.cells, m.perm, prm, 0.0)
    assert set(np.unique(marks0)).issubset({-1, 0})
    # interface elements: -1 if len < minH_, +1 if len > maxH_ (quirk 4)
    prm = oracle.Params.default(minH=0.02, maxH=0.03)
    im = oracle.mark_interface_2d(m.xy, m.segs, prm, 1.0)
    ln = np.hypot(*(m.xy[m.segs[:, 1]] - m.xy[m.segs[:, 0]]).T)
    assert np.array_equal(im, np.where(ln < 0.02, -1, np.where(ln > 0.03, 1, 0)))


def test_interface_p1_transfer_pins(oracle):
    """MMeshInterfaceGrid::adapt(handle) with DG-P1 data (interface/grid.hh:388-415): the oracle's definition must reproduce
    linear functions along the interface under refinement (prolong) and coarsening (restrict), conserve the mass of
    discontinuous data, and use the reference's father choice (the longer element)."""
    rng = np.random.default_rng(3)
    f = lambda p: 1.5 + 2.0 * p[..., 0] - 0.7 * p[..., 1]                 # linear in the plane => linear along every segment
    for trial in range(50):
        a, b = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2)
        k = int(rng.integers(2, 6))
        t = np.sort(np.concatenate([[0.0, 1.0], rng.uniform(0.05, 0.95, k - 1)]))
        pts = a + t[:, None] * (b - a)
        fine = np.concatenate([pts[:-1], pts[1:]], axis=1)                 # k segments tiling (a, b)
        coarse = np.array([[a[0], a[1], b[0], b[1]]])
        # refine: every new (fine) element has the coarse one as its single, longer child -> prolong
        got = oracle.project_interface_p1(fine, np.arange(k + 1), np.repeat(coarse, k, 0), np.zeros(k, np.int32), f(coarse.reshape(1, 2, 2)))
        assert np.max(np.abs(got - f(fine.reshape(k, 2, 2)))) <= 1e-13
        # coarsen: the new (coarse) element has the k shorter fine ones as children -> restrict
        got = oracle.project_interface_p1(coarse, [0, k], fine, np.arange(k, dtype=np.int32), f(fine.reshape(k, 2, 2)))
        assert np.max(np.abs(got - f(coarse.reshape(1, 2, 2)))) <= 1e-12
        u = rng.uniform(-1, 1, (k, 2))                                     # discontinuous data: the mass is conserved
        got = oracle.project_interface_p1(coarse, [0, k], fine, np.arange(k, dtype=np.int32), u)
        lens = np.linalg.norm(fine[:, 2:] - fine[:, :2], axis=1)
        assert abs(np.linalg.norm(b - a) * got.mean() - (lens * u.mean(axis=1)).sum()) <= 1e-13
