"""GPU parity tests: every kernel of the hot path, called through the C-ABI (mmesh_b200.capi -> libmmesh_b200.so),
against the CPU oracle on the same seeded inputs.  Integer / flag / mark outputs must be bit-exact; distances are
asserted bit-exact too (same arithmetic, conservative culling); projected DOFs within 1e-12 relative, mass 1e-13.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mesh1(n=100):
    from mmesh_b200 import synth
    return synth.config1(n=n)


def _upload(ctx, m):
    ctx.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs)


def test_library_fails_loudly_without_fallback():
    """The product path has no CPU fallback: a bad device index is an error, not a silent CPU run."""
    from mmesh_b200 import capi
    with pytest.raises(capi.MMeshError):
        capi.Context(dim=2, device=977)


def test_move_and_trial_validity_2d(gpu_ctx_factory, oracle):
    from mmesh_b200 import synth
    m = _mesh1()
    ctx = gpu_ctx_factory(2)
    _upload(ctx, m)
    # large shifts: many inverted cells, so both branches of `signedVolume_ <= 0` are exercised
    for amp, seed in ((0.3, 30), (0.9, 31)):
        s = synth.shifts_uniform(m, amp * m.h, seed)
        valid, sign, ninv = ctx.trial_move_validate(s)
        rv, rs, rn = oracle.trial_validate(m.xy, s, m.cells)
        assert np.array_equal(valid, rv) and np.array_equal(sign, rs) and ninv == rn
        assert np.array_equal(ctx.invalid_cells(), np.nonzero(rv == 0)[0])
        # trial move must leave the coordinates untouched (restore_mode = exact copy)
        assert np.array_equal(ctx.download_coords(), m.xy)
    assert rn > 0
    # real move: bit-identical to x (+) s
    ctx.move_vertices(s)
    assert np.array_equal(ctx.download_coords(), oracle.move(m.xy, s))
    # and the reference's move-back quirk (x (+) s) (+) (-s)
    ctx.scale_shifts(-1.0)
    ctx.move_vertices(None)
    assert np.array_equal(ctx.download_coords(), oracle.move(oracle.move(m.xy, s), -s))


def test_validity_zero_area_cells(gpu_ctx_factory, oracle):
    """Cells with exactly zero / tiny signed area: valid_fp follows the inexact test, orient_sign the exact one."""
    m = _mesh1(20)
    xy = m.xy.copy()
    # collapse some vertices onto neighbours / make cells exactly degenerate
    c = m.cells
    xy[c[::7, 2]] = 0.5 * (xy[c[::7, 0]] + xy[c[::7, 1]])
    ctx = gpu_ctx_factory(2)
    ctx.upload_mesh(xy, m.cells, m.perm, m.edges, m.segs)
    s = np.zeros_like(xy)
    valid, sign, ninv = ctx.trial_move_validate(s)
    rv, rs, rn = oracle.trial_validate(xy, s, m.cells)
    assert np.array_equal(valid, rv) and np.array_equal(sign, rs) and ninv == rn
    assert (rs == 0).sum() > 0


def test_legality(gpu_ctx_factory, oracle):
    m = _mesh1()
    ctx = gpu_ctx_factory(2)
    _upload(ctx, m)
    flags, nill = ctx.legality()
    rf, rn = oracle.legality(m.xy, None, m.edges)
    assert np.array_equal(flags, rf) and nill == rn and nill > 0
    assert (flags == 2).sum() == 400


def test_distance_bit_exact(gpu_ctx_factory, oracle):
    from mmesh_b200 import synth
    for m in (_mesh1(), synth.config2(n=120)):
        ctx = gpu_ctx_factory(2)
        _upload(ctx, m)
        d, mx = ctx.distance_update()
        rd = oracle.distance_2d(m.xy, m.segs)
        assert np.array_equal(d, rd)
        assert mx == oracle.distance_max(rd)


def test_distance_no_interface(gpu_ctx_factory, oracle):
    m = _mesh1(10)
    ctx = gpu_ctx_factory(2)
    ctx.upload_mesh(m.xy, m.cells, m.perm, m.edges, None)
    d, mx = ctx.distance_update()
    assert np.all(d == 1e100) and mx == 1e100            # distance.hh:45


def test_marks_bit_exact(gpu_ctx_factory, oracle):
    from mmesh_b200 import synth, capi
    for m in (_mesh1(), synth.config2(n=120)):
        ctx = gpu_ctx_factory(2)
        _upload(ctx, m)
        prm = ctx.indicator_init()
        oprm = oracle.indicator_init_2d(m.xy, m.edges, m.segs, oracle.Params.default())
        assert (prm.minH, prm.maxH, prm.factor) == (oprm.minH, oprm.maxH, oprm.factor)
        rd = oracle.distance_2d(m.xy, m.segs)
        for scale_min, scale_max, prop in ((1.0, 1.0, 1.0), (3.0, 0.45, 0.5), (0.2, 0.3, 0.25)):
            p = capi.Params.from_buffer_copy(bytes(prm))
            p.minH, p.maxH, p.distProportion = prm.minH * scale_min, prm.maxH * scale_max, prop
            ctx.set_params(p)
            marks, le, counts = ctx.mark_elements(run_distance=True)
            op = oracle.Params.from_buffer_copy(bytes(p))
            rm, rle, rc = oracle.mark(m.xy, rd, m.cells, m.perm, op, prop * oracle.distance_max(rd))
            assert np.array_equal(marks, rm) and np.array_equal(le, rle) and tuple(counts) == tuple(rc)
            assert np.array_equal(ctx.marked_cells(+1), np.nonzero(rm == 1)[0])
            assert np.array_equal(ctx.marked_cells(-1), np.nonzero(rm == -1)[0])
            assert np.array_equal(ctx.mark_interface(), oracle.mark_interface_2d(m.xy, m.segs, op, prop * oracle.distance_max(rd)))
        assert len(set(np.unique(rm))) >= 2


def test_marks_nan_semantics(gpu_ctx_factory, oracle):
    """maxDist == 0 => l = NaN: device must keep IEEE NaN comparisons."""
    from mmesh_b200 import capi
    m = _mesh1(30)
    ctx = gpu_ctx_factory(2)
    _upload(ctx, m)
    prm = ctx.indicator_init()
    ctx.distance_set(np.zeros(m.nv))
    marks, le, counts = ctx.mark_elements(run_distance=False)
    rm, rle, rc = oracle.mark(m.xy, np.zeros(m.nv), m.cells, m.perm, oracle.Params.from_buffer_copy(bytes(prm)), 0.0)
    assert np.array_equal(marks, rm) and np.array_equal(le, rle)


def test_projection_p0_p1(gpu_ctx_factory, oracle):
    from mmesh_b200 import capi
    from projection_cases import diagonal_flip_case
    case = diagonal_flip_case(n=64, seed=9)
    ctx = gpu_ctx_factory(2)
    ctx.upload_mesh(case["xy"], case["cells"], case["perm"], None, None)
    ctx.upload_pairs(case["new_off"], case["new_cell"], case["old_xy"], case["old_idx"])
    rng = np.random.default_rng(3)
    for translate in (0, 1):
        prm = ctx.get_params()
        prm.clip_translate = translate
        ctx.set_params(prm)
        oprm = oracle.Params.from_buffer_copy(bytes(prm))
        u_old = rng.uniform(-1, 2, case["n_old"])
        got, mass, cov, npc = ctx.project(1, u_old)
        ref, rmass, rcov, rnpc = oracle.project(case["xy"], case["cells"], case["perm"], case["new_off"], case["new_cell"],
                                                case["old_xy"], case["old_idx"], u_old, case["cells"].shape[0], 1, oprm)
        assert npc == rnpc
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
        if translate == 0:
            assert np.array_equal(got, ref)      # same accumulation order => same bits
        assert abs(mass - rmass) <= 1e-13 * abs(rmass) and abs(cov - rcov) <= 1e-13 * rcov
        mass_old = float(np.sum(case["old_vol"] * u_old))
        assert abs(mass - mass_old) <= 1e-13 * abs(mass_old)
        d_old = rng.uniform(-1, 2, (case["n_old"], 3))
        got, mass, cov, npc = ctx.project(3, d_old)
        ref, rmass, rcov, rnpc = oracle.project(case["xy"], case["cells"], case["perm"], case["new_off"], case["new_cell"],
                                                case["old_xy"], case["old_idx"], d_old, case["cells"].shape[0], 3, oprm)
        assert npc == rnpc
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
        assert abs(mass - rmass) <= 1e-13 * abs(rmass)
    # P1 reproduces linears, conserves mass
    f = lambda p: 0.3 + 1.7 * p[..., 0] - 0.9 * p[..., 1]
    d_old = f(case["old_xy"].reshape(-1, 3, 2)[case["first_pair_of_old"]])
    got, mass, _, _ = ctx.project(3, d_old)
    assert np.max(np.abs(got - f(case["new_corners_idorder"]))) < 1e-12
    assert abs(mass - float(np.sum(case["old_vol"] * d_old.mean(axis=1)))) <= 1e-13 * abs(mass)
    # intersection volumes (cachingentity.hh:154-167) and the conservation check of movement.cc:89-94
    vols = ctx.intersection_volumes()
    ref = np.array([oracle.intersection_volume(case["old_xy"][p], case["new6"][i])
                    for i in range(case["new_cell"].shape[0]) for p in range(case["new_off"][i], case["new_off"][i + 1])])
    assert np.array_equal(vols, ref)
    assert np.max(np.abs(np.add.reduceat(vols, case["new_off"][:-1]) - case["new_vol"])) < 1e-8


def test_projection_stress_pairs(gpu_ctx_factory, oracle):
    """The benchmark's pair set-up (self + face neighbours, moved mesh) incl. dropped slivers and empty cuts."""
    from mmesh_b200 import synth
    m = _mesh1(60)
    s = synth.shifts_uniform(m, 0.3 * m.h, 30)
    new_off, new_cell, old_xy, old_idx = synth.pairs_self_and_neighbours(m, m.xy - s)
    ctx = gpu_ctx_factory(2)
    _upload(ctx, m)
    ctx.upload_pairs(new_off, new_cell, old_xy, old_idx)
    prm = ctx.get_params()
    prm.drop_area = 1e-8
    ctx.set_params(prm)
    oprm = oracle.Params.from_buffer_copy(bytes(prm))
    rng = np.random.default_rng(4)
    for ncomp in (1, 3):
        u = rng.uniform(0, 1, (m.nc, ncomp)) if ncomp == 3 else rng.uniform(0, 1, m.nc)
        got, mass, cov, npc = ctx.project(ncomp, u)
        ref, rmass, rcov, rnpc = oracle.project(m.xy, m.cells, m.perm, new_off, new_cell, old_xy, old_idx, u, m.nc, ncomp, oprm)
        assert npc == rnpc
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
        assert abs(mass - rmass) <= 1e-13 * abs(rmass) and abs(cov - rcov) <= 1e-13 * rcov


def test_predicate_batches_config4(gpu_ctx_factory, oracle):
    """BASELINE config 4: cocircular/collinear lattice perturbed by 1e-15 -- the exact fallback decides everything."""
    from mmesh_b200 import synth
    ctx = gpu_ctx_factory(2)
    for perturb in (0.0, 1e-15):
        m = synth.config4(n=128, perturb=perturb)
        tri = np.concatenate([m.xy[m.cells[:, 0]], m.xy[m.cells[:, 1]], m.xy[m.cells[:, 2]]], 1)
        got, nex = ctx.predicate_batch(0, tri)
        assert np.array_equal(got, oracle.orient2d_batch(tri))
        e = m.edges[m.edges[:, 3] >= 0]
        quad = np.concatenate([m.xy[e[:, 0]], m.xy[e[:, 1]], m.xy[e[:, 2]], m.xy[e[:, 3]]], 1)
        got, nex = ctx.predicate_batch(1, quad)
        ref = oracle.incircle_batch(quad)
        assert np.array_equal(got, ref)
        if perturb == 0.0:
            lat = m.lat
            diag = np.abs(lat[e[:, 0]] - lat[e[:, 1]]).sum(1) == 2
            assert np.all(ref[diag] == 0) and nex == 0        # cocircular quads: decided by the interval stage (exact zeros)
        else:
            assert nex > 0
        # collinear lattice triples (i, i+k, i+2k)
        rng = np.random.default_rng(6)
        a = rng.integers(0, 64, size=(20000, 2))
        d = rng.integers(-3, 4, size=(20000, 2))
        P = np.stack([a, a + d, a + 2 * d], 1) / 128.0
        if perturb:
            P = P + rng.uniform(-perturb, perturb, P.shape)
        got, _ = ctx.predicate_batch(0, P.reshape(-1, 6))
        ref = oracle.orient2d_batch(P.reshape(-1, 6))
        assert np.array_equal(got, ref)
        if perturb == 0.0:
            assert np.all(ref == 0)
    # the mesh-level kernels on the perturbed lattice
    _upload(ctx, m)
    valid, sign, ninv = ctx.trial_move_validate(np.zeros_like(m.xy))
    rv, rs, rn = oracle.trial_validate(m.xy, None, m.cells)
    assert np.array_equal(valid, rv) and np.array_equal(sign, rs)
    flags, nill = ctx.legality()
    rf, rnill = oracle.legality(m.xy, None, m.edges)
    assert np.array_equal(flags, rf) and nill == rnill


def test_orient3d_batch(gpu_ctx_factory, oracle):
    rng = np.random.default_rng(8)
    base = rng.integers(0, 6, size=(30000, 4, 3)) / 8.0
    pts = base + np.where(rng.uniform(size=(30000, 1, 1)) < 0.5, 0.0, rng.uniform(-1e-15, 1e-15, size=(30000, 4, 3)))
    ctx = gpu_ctx_factory(3)
    got, nex = ctx.predicate_batch(2, pts.reshape(-1, 12))
    ref = oracle.orient3d_batch(pts.reshape(-1, 12))
    assert np.array_equal(got, ref) and (ref == 0).sum() > 100


def test_3d_path(gpu_ctx_factory, oracle):
    from mmesh_b200 import synth, capi
    m = synth.lattice3d(12)
    ctx = gpu_ctx_factory(3)
    ctx.upload_mesh(m.xyz, m.cells, m.perm, None, m.tris)
    s = synth.shifts_uniform(m, 0.05 * m.h, 50)
    valid, sign, ninv = ctx.trial_move_validate(s)
    rv, rs, rn = oracle.trial_validate(m.xyz, s, m.cells, dim=3)
    assert np.array_equal(valid, rv) and np.array_equal(sign, rs) and ninv == rn == 0
    # an invalid 3-D cell is the GridError of mmesh.hh:1104-1107
    big = synth.shifts_uniform(m, 0.9 * m.h, 51)
    with pytest.raises(capi.GridError) as ei:
        ctx.trial_move_validate(big)
    rv, rs, rn = oracle.trial_validate(m.xyz, big, m.cells, dim=3)
    assert np.array_equal(ei.value.valid_fp, rv) and np.array_equal(ei.value.orient_sign, rs) and ei.value.n_invalid == rn
    d, mx = ctx.distance_update()
    rd = oracle.distance_3d(m.xyz, m.tris)
    assert np.array_equal(d, rd) and mx == oracle.distance_max(rd)
    prm = ctx.get_params()
    prm.minH, prm.maxH, prm.factor = 0.25 * m.h, 1.3 * m.h, 1.5
    ctx.set_params(prm)
    marks, le, counts = ctx.mark_elements(run_distance=False)
    rm, rle, rc = oracle.mark(m.xyz, rd, m.cells, m.perm, oracle.Params.from_buffer_copy(bytes(prm)), oracle.distance_max(rd), dim=3)
    assert np.array_equal(marks, rm) and np.array_equal(le, rle) and tuple(counts) == tuple(rc)
    assert 0 < rc[0] < m.nc
    # insertion candidates of adapt() in 3-D (mmesh.hh:828-851): edge-id dedup over all cells around an edge
    want_cells, want_ev = oracle.refine_candidates_3d(m.cells, m.perm, rm, rle, m.tris)
    got_cells = ctx.refinement_candidates()
    assert got_cells.tolist() == want_cells.tolist() and 0 < len(want_cells) < rc[0]
    pts, ev = ctx.refinement_points(got_cells)
    assert np.array_equal(ev, want_ev)
    a, b = m.xyz[ev[:, 0]], m.xyz[ev[:, 1]]
    assert np.array_equal(pts, a + 0.5 * (b - a))
    ctx.move_vertices(s)
    assert np.array_equal(ctx.download_coords(), oracle.move(m.xyz, s))
    # P0 trace / skeleton transfer
    rng = np.random.default_rng(2)
    nf = m.tris.shape[0]
    inside = rng.integers(0, m.nc, nf).astype(np.int32)
    outside = rng.integers(-1, m.nc, nf).astype(np.int32)
    u = rng.uniform(size=m.nc)
    ug = rng.uniform(size=nf)
    tin, tout, sk = ctx.trace_skeleton_p0(inside, outside, u, ug)
    assert np.array_equal(tin, u[inside]) and np.array_equal(tout, np.where(outside >= 0, u[np.maximum(outside, 0)], 0.0))
    assert np.array_equal(sk, ug)


def test_adapt_step_composition(gpu_ctx_factory, oracle):
    """mmb_adapt_step == markElements -> ensureVertexMovement -> adapt(handle) -> moveVertices -> legality."""
    from mmesh_b200 import synth
    m = _mesh1(80)
    s = synth.shifts_uniform(m, 0.3 * m.h, 30)
    new_off, new_cell, old_xy, old_idx = synth.pairs_self_and_neighbours(m, m.xy - s)
    ctx = gpu_ctx_factory(2)
    _upload(ctx, m)
    prm = ctx.indicator_init()
    prm.move_policy = 1                       # always move (the 0.3 h shifts invert a few cells; policy 0 is tested below)
    ctx.set_params(prm)
    ctx.upload_pairs(new_off, new_cell, old_xy, old_idx)
    rng = np.random.default_rng(5)
    dofs = rng.uniform(size=(m.nc, 3))
    ctx.project(3, dofs, download=False)      # makes the DOFs resident
    ctx.upload_shifts(s)
    res = ctx.adapt_step(None, ncomp=3)
    assert res.moved == 1
    oprm = oracle.Params.from_buffer_copy(bytes(prm))
    rd = oracle.distance_2d(m.xy, m.segs)
    rm, _, rc = oracle.mark(m.xy, rd, m.cells, m.perm, oprm, oracle.distance_max(rd))
    rv, rs, rn = oracle.trial_validate(m.xy, s, m.cells)
    _, rmass, rcov, rnpc = oracle.project(m.xy, m.cells, m.perm, new_off, new_cell, old_xy, old_idx, dofs, m.nc, 3, oprm)
    moved = oracle.move(m.xy, s)
    rf, rnill = oracle.legality(moved, None, m.edges)
    assert res.n_invalid == rn and res.n_orient_nonpos == (rs <= 0).sum()
    assert (res.n_refine, res.n_coarsen) == tuple(rc)
    assert res.n_illegal == rnill and res.n_pieces == rnpc
    assert res.max_dist == oracle.distance_max(rd)
    assert abs(res.mass_new - rmass) <= 1e-13 * abs(rmass) and abs(res.covered_area - rcov) <= 1e-13 * rcov
    assert np.array_equal(ctx.download_coords(), moved)
    assert ctx.launch_count > 0


def test_adapt_step_host_buffers(gpu_ctx_factory, oracle):
    """mmb_adapt_step_host (pipelined copies, chunked projection) returns the same fields as the separate calls."""
    from mmesh_b200 import synth
    for n in (80, 200):                                   # 200: > 2^16 new cells => the chunked projection path
        m = _mesh1(n)
        s = synth.shifts_uniform(m, 0.3 * m.h, 30)
        new_off, new_cell, old_xy, old_idx = synth.pairs_self_and_neighbours(m, m.xy - s)
        ctx = gpu_ctx_factory(2)
        _upload(ctx, m)
        prm = ctx.indicator_init()
        prm.drop_area = 1e-12
        prm.move_policy = 1
        ctx.set_params(prm)
        ctx.upload_pairs(new_off, new_cell, old_xy, old_idx)
        rng = np.random.default_rng(5)
        oprm = oracle.Params.from_buffer_copy(bytes(prm))
        rd = oracle.distance_2d(m.xy, m.segs)
        rm, rle, rc = oracle.mark(m.xy, rd, m.cells, m.perm, oprm, oracle.distance_max(rd))
        rv, rs, rn = oracle.trial_validate(m.xy, s, m.cells)
        rf, rnill = oracle.legality(oracle.move(m.xy, s), None, m.edges)
        for ncomp in (3, 1):
            ctx.upload_coords(m.xy)
            dofs = rng.uniform(size=(m.nc, ncomp)) if ncomp == 3 else rng.uniform(size=m.nc)
            out = ctx.adapt_step_host(s, ncomp, dofs)
            ref, rmass, rcov, rnpc = oracle.project(m.xy, m.cells, m.perm, new_off, new_cell, old_xy, old_idx, dofs, m.nc, ncomp, oprm)
            assert np.array_equal(out["marks"], rm) and np.array_equal(out["longest"], rle)
            assert np.array_equal(out["valid_fp"], rv) and np.array_equal(out["orient_sign"], rs)
            assert np.array_equal(out["edge_flags"], rf)
            assert np.max(np.abs(out["dofs_new"] - ref)) <= 1e-12 * np.max(np.abs(ref))
            r = out["result"]
            assert (r.n_refine, r.n_coarsen) == tuple(rc) and r.n_invalid == rn and r.n_illegal == rnill and r.n_pieces == rnpc
            assert abs(r.mass_new - rmass) <= 1e-13 * abs(rmass) and abs(r.covered_area - rcov) <= 1e-13 * rcov
            assert np.array_equal(ctx.download_coords(), oracle.move(m.xy, s))


def test_step_withholds_move_when_cells_invert(gpu_ctx_factory, oracle):
    """Reference order of the user loop (doc/examples/moving.py:231-252): ensureVertexMovement == true => vertices are removed
    BEFORE moveVertices.  move_policy 0 (default): a step whose trial move inverts a cell leaves the coordinates alone
    (decided on the device), reports moved == 0 and its edge flags describe the unmoved mesh; a step without inverted
    cells moves.  Resident and host-buffer steps agree."""
    from mmesh_b200 import synth
    m = _mesh1(60)
    big, small = synth.shifts_uniform(m, 0.45 * m.h, 30), synth.shifts_uniform(m, 0.05 * m.h, 31)
    assert oracle.trial_validate(m.xy, big, m.cells)[2] > 0 and oracle.trial_validate(m.xy, small, m.cells)[2] == 0
    for host in (False, True):
        ctx = gpu_ctx_factory(2)
        _upload(ctx, m)
        ctx.indicator_init()
        assert ctx.get_params().move_policy == 0
        step = (lambda s: ctx.adapt_step_host(s, 0, None)["result"]) if host else (lambda s: ctx.adapt_step(s, 0))
        r = step(big)
        assert r.moved == 0 and r.n_invalid == oracle.trial_validate(m.xy, big, m.cells)[2]
        assert np.array_equal(ctx.download_coords(), m.xy)
        assert r.n_illegal == oracle.legality(m.xy, None, m.edges)[1]
        r = step(small)
        moved = oracle.move(m.xy, small)
        assert r.moved == 1 and r.n_invalid == 0 and np.array_equal(ctx.download_coords(), moved)
        assert r.n_illegal == oracle.legality(moved, None, m.edges)[1]
        prm = ctx.get_params(); prm.move_policy = 1; ctx.set_params(prm)
        r = step(big)
        assert r.moved == 1 and np.array_equal(ctx.download_coords(), oracle.move(moved, big))


def test_interface_projection_and_refinement_points(gpu_ctx_factory, oracle):
    """interface/grid.hh:388-415 (1-D father/child by length) and longestedgerefinement.hh:41-57 (edge midpoints)."""
    rng = np.random.default_rng(12)
    nnew, nold = 5000, 7000
    cnt = rng.integers(1, 4, nnew)
    new_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    P = int(new_off[-1])
    new_len = rng.uniform(0.5, 1.5, nnew)
    old_len = rng.uniform(0.2, 2.0, P)
    old_idx = rng.integers(0, nold, P).astype(np.int32)
    dofs = rng.uniform(-1, 1, nold)
    ctx = gpu_ctx_factory(2)
    got = ctx.project_interface_p0(new_len, new_off, old_len, old_idx, dofs)
    assert np.array_equal(got, oracle.project_interface_p0(new_len, new_off, old_len, old_idx, dofs))
    m = _mesh1(60)
    _upload(ctx, m)
    prm = ctx.indicator_init()
    prm.maxH *= 0.3
    ctx.set_params(prm)
    marks, le, counts = ctx.mark_elements(run_distance=True)
    cells = ctx.marked_cells(+1)
    assert len(cells) > 0
    pts, ev = ctx.refinement_points(cells)
    EV = np.array([[0, 1], [0, 2], [1, 2]])
    for i, c in enumerate(cells[:2000]):
        _, _, ec, _ = oracle.cell_geometry_2d(m.xy, m.cells[c], m.perm[c])
        assert np.array_equal(pts[i], ec[le[c]])
        k = [int(m.cells[c][(m.perm[c] >> (2 * t)) & 3]) for t in range(3)]
        assert list(ev[i]) == [k[EV[le[c]][0]], k[EV[le[c]][1]]]


def test_projection_ragged_csr(gpu_ctx_factory, oracle):
    """Ragged pair lists: new cells with 0..9 old children (more than one round of the per-cell loop), arbitrary old
    triangles (clockwise ones too, degenerate ones, far-away ones), non-contiguous new cells; and an empty pair set."""
    from mmesh_b200 import synth
    m = _mesh1(40)
    rng = np.random.default_rng(21)
    ctx = gpu_ctx_factory(2)
    _upload(ctx, m)
    new_cell = np.sort(rng.choice(m.nc, size=m.nc // 3, replace=False)).astype(np.int32)
    cnt = rng.integers(0, 10, new_cell.shape[0])
    new_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    P = int(new_off[-1])
    ctr = m.xy[m.cells[np.repeat(new_cell, cnt)]].mean(axis=1)
    old = ctr[:, None, :] + rng.uniform(-1.5 * m.h, 1.5 * m.h, size=(P, 3, 2))
    old[::17, 2] = 0.5 * (old[::17, 0] + old[::17, 1])            # degenerate (zero area) old cells
    old[::23] += 0.4                                              # far away: empty intersection
    old_xy = old.reshape(P, 6)
    old_idx = rng.integers(0, 500, P).astype(np.int32)
    ctx.upload_pairs(new_off, new_cell, old_xy, old_idx)
    prm = ctx.get_params()
    prm.drop_area = 1e-9
    ctx.set_params(prm)
    oprm = oracle.Params.from_buffer_copy(bytes(prm))
    for ncomp in (1, 3):
        dofs = rng.uniform(-1, 1, (500, ncomp)) if ncomp == 3 else rng.uniform(-1, 1, 500)
        ctx.project(ncomp, dofs * 0 + 7.0, download=False)        # stale values must not leak into untouched cells... 
        got, mass, cov, npc = ctx.project(ncomp, dofs)
        ref, rmass, rcov, rnpc = oracle.project(m.xy, m.cells, m.perm, new_off, new_cell, old_xy, old_idx, dofs, m.nc, ncomp, oprm)
        sel = new_cell
        assert npc == rnpc
        g, r = (got[sel], ref[sel])
        assert np.max(np.abs(g - r)) <= 1e-12 * max(1.0, np.max(np.abs(r)))
        if ncomp == 1:
            assert np.array_equal(g, r)
        assert abs(mass - rmass) <= 1e-13 * max(abs(rmass), 1e-300) + 1e-18 and abs(cov - rcov) <= 1e-13 * rcov
    vols = ctx.intersection_volumes()
    ref_v = np.array([oracle.intersection_volume(old_xy[p], m.xy[m.cells[new_cell[i]]].reshape(6))
                      for i in range(new_cell.shape[0]) for p in range(new_off[i], new_off[i + 1])])
    assert np.array_equal(vols, ref_v)
    # empty pair set
    ctx.upload_pairs(np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros((0, 6)), np.zeros(0, np.int32))
    _, mass, cov, npc = ctx.project(1, np.ones(4))
    assert (mass, cov, npc) == (0.0, 0.0, 0)


def test_malformed_snapshots_are_rejected(gpu_ctx_factory):
    """Out-of-range indices / malformed CSR are status codes (MMB_ERR_INVALID_ARG), never device faults."""
    from mmesh_b200 import capi
    m = _mesh1(10)
    ctx = gpu_ctx_factory(2)
    bad = m.cells.copy(); bad[5, 1] = m.nv
    with pytest.raises(capi.MMeshError) as e:
        ctx.upload_mesh(m.xy, bad, m.perm, m.edges, m.segs)
    assert e.value.status == capi.MMB_ERR_INVALID_ARG
    badp = m.perm.copy(); badp[3] = 0
    with pytest.raises(capi.MMeshError):
        ctx.upload_mesh(m.xy, m.cells, badp, m.edges, m.segs)
    bade = m.edges.copy(); bade[7, 3] = -2
    with pytest.raises(capi.MMeshError):
        ctx.upload_mesh(m.xy, m.cells, m.perm, bade, m.segs)
    _upload(ctx, m)
    with pytest.raises(capi.MMeshError):      # state error: marks before any distance update
        ctx.mark_elements(run_distance=False)
    with pytest.raises(capi.MMeshError):      # CSR end does not match the pair count
        ctx.upload_pairs(np.array([0, 2], np.int64), np.array([0], np.int32), np.zeros((3, 6)), np.zeros(3, np.int32))
    ctx.upload_pairs(np.array([0, 1], np.int64), np.array([0], np.int32), np.zeros((1, 6)), np.array([9], np.int32))
    with pytest.raises(capi.MMeshError):      # old index 9 but only 4 old DOFs
        ctx.project(1, np.ones(4))
    with pytest.raises(capi.MMeshError):      # shifts == NULL without resident shifts
        ctx.adapt_step(None, 0)


def test_refinement_candidates_like_adapt(gpu_ctx_factory, oracle):
    """MMesh::adapt() insertion candidates (mmesh.hh:828-851): per marked cell in index order, longest edge, skipped when it
    is an interface edge or its edge id was inserted before; emulated here with a python set like `inserted_`."""
    from mmesh_b200 import synth
    for m in (_mesh1(70), synth.config2(n=90)):
        ctx = gpu_ctx_factory(2)
        _upload(ctx, m)
        ctx.upload_neighbours(m.nb)
        prm = ctx.indicator_init()
        prm.maxH *= 0.28
        prm.minH *= 0.2
        ctx.set_params(prm)
        marks, le, counts = ctx.mark_elements(run_distance=True)
        iface = {tuple(sorted(e)) for e in m.segs.tolist()}
        EV = [(0, 1), (0, 2), (1, 2)]
        inserted, expect = set(), []
        for c in np.nonzero(marks == 1)[0]:
            k = [int(m.cells[c][(int(m.perm[c]) >> (2 * t)) & 3]) for t in range(3)]
            e = tuple(sorted((k[EV[le[c]][0]], k[EV[le[c]][1]])))
            if e in iface:
                continue
            if e not in inserted:
                inserted.add(e)
                expect.append(int(c))
        got = ctx.refinement_candidates()
        assert counts[0] > 100 and len(expect) < counts[0]          # some edges are shared by two marked cells
        assert got.tolist() == expect
        pts, ev = ctx.refinement_points(got)
        assert len({tuple(sorted(x)) for x in ev.tolist()}) == len(got)      # unique edges


def test_adapt_candidate_lists_match_oracle(gpu_ctx_factory, oracle):
    """Both candidate loops of MMesh::adapt() (mmesh.hh:826-908) against the oracle: insert_ cells, remove_ (cell, vertex)
    pairs with the libstdc++-pinned coarsening choice, interface insert_/remove_ lists.  Bit-exact index lists."""
    from mmesh_b200 import synth
    rng = np.random.default_rng(11)
    for m, fmax, fmin in ((_mesh1(70), 0.4, 1.6), (synth.config2(n=90), 0.5, 1.4)):
        ctx = gpu_ctx_factory(2)
        _upload(ctx, m)
        ctx.upload_neighbours(m.nb)
        level = rng.integers(0, 3, m.nv).astype(np.int32)
        prm = ctx.indicator_init()
        prm.maxH *= fmax
        prm.minH *= fmin
        ctx.set_params(prm)
        marks, le, counts = ctx.mark_elements(run_distance=True)
        imark = ctx.mark_interface()
        assert counts[0] > 50 and counts[1] > 50
        for lev in (None, level):
            if lev is not None:
                ctx.upload_vertex_levels(lev)
            want = oracle.adapt_candidates_2d(m.xy, m.cells, m.perm, m.nb, marks, le, m.segs, imark, lev)
            assert ctx.refinement_candidates().tolist() == want["insert_cells"].tolist()
            cells, verts = ctx.coarsening_candidates()
            assert cells.tolist() == want["remove_cells"].tolist()
            assert verts.tolist() == want["remove_vertices"].tolist()
            ins, rs, rv = ctx.interface_candidates()
            assert ins.tolist() == want["insert_segs"].tolist()
            assert rs.tolist() == want["remove_segs"].tolist()
            assert rv.tolist() == want["remove_seg_vertices"].tolist()
            assert len(want["remove_cells"]) > 20


def test_component_numbers_like_adapt(gpu_ctx_factory, oracle):
    """getComponentNumber_ fix point (mmesh.hh:1194-1218, 1919-1941) on the device against the oracle's sequential loop"""
    from test_oracle_components import random_zones
    m = _mesh1(40)
    ctx = gpu_ctx_factory(2)
    _upload(ctx, m)
    rng = np.random.default_rng(12)
    for nz, size in ((0, 3), (1, 3), (50, 4), (3000, 8), (20000, 3)):
        off, cells = random_zones(rng, m.nc, nz, max_size=size)
        base = int(rng.integers(1, 9))
        zn, cn, cc = ctx.component_numbers(off, cells, base)
        rz, rc, rcc = oracle.component_numbers(off, cells, m.nc, base)
        assert np.array_equal(zn, rz) and np.array_equal(cn, rc) and cc == rcc
    # the zones adapt_() really builds: stars of removed vertices and the two cells of split edges (overlapping on purpose)
    vs = rng.choice(m.nv, 300, replace=False)
    star = [np.nonzero((m.cells == v).any(axis=1))[0] for v in vs]
    off = np.concatenate([[0], np.cumsum([len(s) for s in star])]).astype(np.int64)
    cells = np.concatenate(star).astype(np.int32)
    zn, cn, cc = ctx.component_numbers(off, cells, 1)
    rz, rc, rcc = oracle.component_numbers(off, cells, m.nc, 1)
    assert np.array_equal(zn, rz) and np.array_equal(cn, rc) and cc == rcc and len(set(rz.tolist())) < len(rz)


def test_predicate_exponent_range(gpu_ctx_factory, oracle):
    """The exact stage rescales by a power of two: any absolute scale is decided exactly; an exponent RATIO beyond what
    expansions can hold is reported (MMB_ERR_CAPACITY), never answered with a guessed sign."""
    from mmesh_b200 import capi
    ctx = gpu_ctx_factory(2)
    rng = np.random.default_rng(8)
    base = rng.integers(0, 8, size=(4000, 4, 2)) / 8.0
    base[1000:] += rng.uniform(-1e-15, 1e-15, size=(3000, 4, 2))
    for scale in (1e-290, 1e-150, 1.0, 1e150, 1e290):
        q = (base * scale).reshape(-1, 8)
        if scale < 1e-200:
            q[np.abs(q) < 1e-307] = 0.0                       # keep the inputs normal doubles
        si, _ = ctx.predicate_batch(1, q)
        assert np.array_equal(si, oracle.incircle_batch(q))
        so, _ = ctx.predicate_batch(0, q[:, :6])
        assert np.array_equal(so, oracle.orient2d_batch(q[:, :6]))
    # a rectangle (cocircular corners) whose sides differ by 2^230: filter and interval abstain, the exact stage cannot hold it
    a, b = 1.0 / 3.0, (1.0 / 3.0) * 2.0 ** -230
    bad = np.array([[0.0, 0.0, a, 0.0, a, b, 0.0, b]])
    with pytest.raises(capi.MMeshError) as ei:
        ctx.predicate_batch(1, bad)
    assert ei.value.status == capi.MMB_ERR_CAPACITY
    b2 = a * 2.0 ** -200                                      # the same rectangle inside the range: decided exactly (0)
    ok = np.array([[0.0, 0.0, a, 0.0, a, b2, 0.0, b2]])
    sg, nex = ctx.predicate_batch(1, ok)
    assert sg[0] == 0 == oracle.incircle_batch(ok)[0] and nex == 1

def test_interface_projection_p1(gpu_ctx_factory, oracle):
    """interface/grid.hh:388-415 with DG-P1 data on the segments: device == oracle bit for bit (mixed prolong / restrict
    rows, ragged child lists, degenerate lengths)"""
    ctx = gpu_ctx_factory(2)
    m = _mesh1(20)
    _upload(ctx, m)
    rng = np.random.default_rng(6)
    nnew, n_old = 5000, 3000
    cnt = rng.integers(0, 5, nnew)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    P = int(off[-1])
    new_xy = rng.uniform(0, 1, (nnew, 4))
    old_xy = rng.uniform(0, 1, (P, 4))
    row = np.repeat(np.arange(nnew), cnt)
    sub = rng.random(P) < 0.5                                            # half of the children lie inside their new element
    ta, tb = np.sort(rng.uniform(0, 1, (2, P)), axis=0)
    d = new_xy[row, 2:] - new_xy[row, :2]
    old_xy[sub, :2] = (new_xy[row, :2] + ta[:, None] * d)[sub]
    old_xy[sub, 2:] = (new_xy[row, :2] + tb[:, None] * d)[sub]
    old_xy[::97, 2:] = old_xy[::97, :2]                                  # zero-length children
    old_idx = rng.integers(0, n_old, P).astype(np.int32)
    dofs = rng.uniform(-1, 1, (n_old, 2))
    got = ctx.project_interface_p1(new_xy, off, old_xy, old_idx, dofs)
    ref = oracle.project_interface_p1(new_xy, off, old_xy, old_idx, dofs)
    assert np.array_equal(got, ref, equal_nan=True)


def test_3d_indicator_init_and_interface_marks(gpu_ctx_factory, oracle):
    """RatioIndicator::init (ratioindicator.hh:62-91) and the interface marks of markElements (mmesh.hh:745-748) in 3-D"""
    from mmesh_b200 import synth
    m = synth.lattice3d(14)
    ctx = gpu_ctx_factory(3)
    ctx.upload_mesh(m.xyz, m.cells, m.perm, None, m.tris)
    prm = ctx.indicator_init()
    oprm = oracle.indicator_init_3d(m.xyz, m.cells, m.tris, oracle.Params.default())
    assert (prm.minH, prm.maxH, prm.factor) == (oprm.minH, oprm.maxH, oprm.factor) and prm.maxH > prm.minH > 0
    d, mx = ctx.distance_update()
    for scale in (1.0, 0.4, 0.3):
        p = ctx.get_params(); p.maxH = prm.maxH * scale; ctx.set_params(p)
        got = ctx.mark_interface()
        ref = oracle.mark_interface_3d(m.xyz, m.tris, oracle.Params.from_buffer_copy(bytes(p)), p.distProportion * mx)
        assert np.array_equal(got, ref)
    assert 0 < (ref == 1).sum() and set(np.unique(ref)).issubset({0, 1})
    # no interface: the bulk edges define the parameters (:84-88)
    ctx2 = gpu_ctx_factory(3)
    ctx2.upload_mesh(m.xyz, m.cells, m.perm, None, None)
    p2 = ctx2.indicator_init()
    o2 = oracle.indicator_init_3d(m.xyz, m.cells, np.zeros((0, 3), np.int32), oracle.Params.default())
    assert (p2.minH, p2.maxH, p2.factor) == (o2.minH, o2.maxH, o2.factor)


def test_patch_mesh_equals_full_upload(gpu_ctx_factory, oracle):
    """mmb_patch_mesh (incremental snapshot update after a small TDS edit): edge-midpoint insertions patched into the resident
    snapshot give the same fields as a context that uploads the edited mesh from scratch -- with the interface unchanged
    (hints of untouched vertices survive) and with a changed interface (hints rebuilt)."""
    from mmesh_b200 import synth
    m = synth.config1(n=40)
    rng = np.random.default_rng(17)
    nv0, nc0 = m.nv, m.nc
    cells = m.cells.astype(np.int64).copy()
    ids = m.vid.astype(np.int64)
    key = {tuple(sorted(t)): i for i, t in enumerate(cells.tolist())}
    inner = np.nonzero(m.edges[:, 3] >= 0)[0]
    rng.shuffle(inner)
    touched = np.zeros(nc0, bool)
    new_pts, new_ids, app = [], [], []
    segset = {tuple(sorted(s)) for s in m.segs.tolist()}
    for e in inner:
        if len(new_pts) == 60:
            break
        a, b, c, d = (int(t) for t in m.edges[e])
        if tuple(sorted((a, b))) in segset:
            continue                                              # bulk insertions only: the interface stays as it is
        L, R = key[tuple(sorted((a, b, c)))], key[tuple(sorted((a, b, d)))]
        if touched[L] or touched[R]:
            continue
        touched[[L, R]] = True
        mid = nv0 + len(new_pts)
        new_pts.append(m.xy[a] + 0.5 * (m.xy[b] - m.xy[a])); new_ids.append(int(ids.max()) + 1 + len(new_ids))
        cells[L] = (a, mid, c); cells[R] = (b, mid, d)            # two of the four new cells replace the old ones in place ...
        app += [(mid, b, c), (mid, a, d)]                         # ... two are appended
    xy1 = np.concatenate([m.xy, np.array(new_pts)])
    id1 = np.concatenate([ids, np.array(new_ids, np.int64)])
    cells1 = np.concatenate([cells, np.array(app, np.int64)]).astype(np.int32)
    perm1 = synth.pack_perm(id1[cells1])
    edges1, _, _ = synth.build_edges(cells1, xy1.shape[0])
    ne0, ne1 = m.edges.shape[0], edges1.shape[0]
    common = min(ne0, ne1)
    ech = np.concatenate([np.nonzero((edges1[:common] != m.edges[:common]).any(axis=1))[0], np.arange(common, ne1)]).astype(np.int32)
    cch = np.concatenate([np.nonzero(touched)[0], np.arange(nc0, cells1.shape[0])]).astype(np.int32)
    vch = np.arange(nv0, xy1.shape[0], dtype=np.int32)
    assert ne1 >= ne0 and len(new_pts) == 60

    def fields(ctx):
        prm = ctx.indicator_init()
        d, mx = ctx.distance_update()
        marks, le, counts = ctx.mark_elements(run_distance=False)
        valid, sign, ninv = ctx.trial_move_validate(np.zeros_like(xy1))
        flags, nill = ctx.legality()
        return dict(prm=bytes(prm), d=d, mx=mx, marks=marks, le=le, valid=valid, sign=sign, flags=flags, nill=nill, imarks=ctx.mark_interface())

    fresh = gpu_ctx_factory(2)
    fresh.upload_mesh(xy1, cells1, perm1, edges1, m.segs)
    want = fields(fresh)
    ctx = gpu_ctx_factory(2)
    _upload(ctx, m)
    ctx.distance_update()                                          # hints of the old mesh
    ctx.patch_mesh(xy1.shape[0], vch, xy1[vch], cells1.shape[0], cch, cells1[cch], perm1[cch], ne1, ech, edges1[ech])
    got = fields(ctx)
    for k in want:
        assert np.array_equal(got[k], want[k]) if isinstance(want[k], np.ndarray) else got[k] == want[k], k
    assert np.array_equal(got["d"], oracle.distance_2d(xy1, m.segs))
    # a patch that also changes the interface (one facet dropped) and moves a vertex
    segs2 = m.segs[:-1]
    xy2 = xy1.copy(); xy2[5] += 1e-3 * m.h
    fresh2 = gpu_ctx_factory(2)
    fresh2.upload_mesh(xy2, cells1, perm1, edges1, segs2)
    want2 = fields(fresh2)
    ctx.patch_mesh(xy2.shape[0], np.array([5], np.int32), xy2[[5]], cells1.shape[0], np.zeros(0, np.int32), np.zeros((0, 3), np.int32), np.zeros(0, np.uint8),
                   ne1, None, None, facets=segs2)
    got2 = fields(ctx)
    for k in want2:
        assert np.array_equal(got2[k], want2[k]) if isinstance(want2[k], np.ndarray) else got2[k] == want2[k], k


@pytest.mark.gpu
def test_restore_delaunay_on_device(gpu_ctx_factory, oracle):
    """mmb_restore_delaunay: after a vertex move the resident snapshot is flipped back to the (constrained)
    Delaunay triangulation.  The device rounds are deterministic, so cells / neighbours / corner permutations equal the host
    build of the same phase functions entry by entry (tests/native/host_flip.cc); the triangle set equals the oracle's
    sequential Lawson iteration and, without constraints, scipy's Qhull triangulation; the rebuilt edge list equals what the
    snapshot builder derives from the new cells, and the legality pass afterwards finds only the interface edges."""
    from mmesh_b200 import synth
    from test_host_flip import moved_mesh, tri_set, facet_keys, SO as FLIP_SO, ROOT as TROOT
    import ctypes as C, subprocess
    os.makedirs(os.path.dirname(FLIP_SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(TROOT, "include"),
                           os.path.join(TROOT, "tests", "native", "host_flip.cc"), "-o", FLIP_SO])
    L = C.CDLL(FLIP_SO)
    P = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    t = np.linspace(0.0, 2.0 * np.pi, 800)
    circle = (np.stack([0.5 + 0.27 * np.cos(t), 0.5 + 0.27 * np.sin(t)], 1), True)
    for n, curves, use_vid in ((96, (), True), (160, (circle,), True), (64, (), False)):
        m, xy = moved_mesh(oracle, n, 0.27, 11 + n, curves=curves)
        ctx = gpu_ctx_factory(2)
        ids = m.vid if use_vid else np.arange(m.nv)                    # without vertex ids the persistent id is the index
        perm0 = synth.pack_perm(ids[m.cells])
        ctx.upload_mesh(xy, m.cells, perm0, m.edges, m.segs)
        with pytest.raises(Exception):
            ctx.restore_delaunay()                                     # neighbours are required
        ctx.upload_neighbours(m.nb)
        _, nill0 = ctx.legality()
        res = ctx.restore_delaunay(vertex_id=m.vid if use_vid else None)
        cells, perm, nb, edges = ctx.download_topology()
        assert res["converged"] == 1 and res["n_illegal_left"] == 0 and 0 < res["n_flips"] and res["rounds"] < 64
        # host build of the same phases: identical arrays
        hc, hnb, hperm = m.cells.copy(), m.nb.copy(), perm0.copy()
        v_if, keys = facet_keys(m.segs.reshape(-1, 2), m.nv)
        r = np.zeros(8, np.int64)
        vid64 = np.ascontiguousarray(m.vid, np.int64) if use_vid else None
        L.hf_restore(P(np.ascontiguousarray(xy)), C.c_long(m.nv), P(hc), P(hnb), P(hperm), C.c_long(m.nc), P(v_if), P(keys), C.c_int(keys.shape[0]),
                     P(vid64), C.c_int(1000), P(r), None)
        assert (res["n_flips"], res["rounds"], res["n_constrained_illegal"]) == (int(r[0]), int(r[1]), int(r[4]))
        assert np.array_equal(cells, hc) and np.array_equal(nb, hnb) and np.array_equal(perm, hperm)
        # the flip log: per round the same set of (cell, slot) as the host build, and replaying it with TDS::flip semantics on
        # the original tables gives the device's tables slot for slot
        from test_host_flip import replay_flips
        fids, off, tot = ctx.flip_log()
        assert tot == res["n_flips"] == len(fids) and len(off) == res["rounds"] + 1 and off[0] == 0 and off[-1] == tot
        hlog = np.zeros(4 * m.nc + 16, np.uint32); htrace = np.zeros(1000, np.int32); hr = np.zeros(8, np.int64)
        lc, lnb, lperm = m.cells.copy(), m.nb.copy(), perm0.copy()
        L.hf_restore_log.restype = C.c_long
        hn = L.hf_restore_log(P(np.ascontiguousarray(xy)), C.c_long(m.nv), P(lc), P(lnb), P(lperm), C.c_long(m.nc), P(v_if), P(keys), C.c_int(keys.shape[0]),
                              P(vid64), C.c_int(1000), P(hr), P(htrace), P(hlog), C.c_long(hlog.shape[0]))
        hoff = np.concatenate([[0], np.cumsum(htrace[:int(hr[1])])])
        assert hn == tot and np.array_equal(hoff, off)
        for r_ in range(len(off) - 1):
            assert np.array_equal(np.sort(fids[off[r_]:off[r_ + 1]]), np.sort(hlog[off[r_]:off[r_ + 1]]))
        if m.nc <= 30000:
            rc_, rnb_ = replay_flips(m.cells, m.nb, fids)
            assert np.array_equal(rc_, cells) and np.array_equal(rnb_, nb)
        # oracle (sequential Lawson) and, unconstrained, Qhull
        oc, _, oflips = oracle.restore_delaunay_2d(xy, m.cells, m.nb, m.segs)
        assert oflips > 0 and tri_set(oc) == tri_set(cells)
        if not curves:
            from scipy.spatial import Delaunay
            sp = Delaunay(xy).simplices
            a = xy[sp]
            area = np.abs((a[:, 1, 0] - a[:, 0, 0]) * (a[:, 2, 1] - a[:, 0, 1]) - (a[:, 2, 0] - a[:, 0, 0]) * (a[:, 1, 1] - a[:, 0, 1]))
            assert tri_set(sp[area > 1e-14]) == tri_set(cells)
        assert np.array_equal(perm, synth.pack_perm(ids[cells]))
        e_ref, nb_ref, _ = synth.build_edges(cells, m.nv)
        assert np.array_equal(edges, e_ref) and np.array_equal(nb, nb_ref)
        # the hot path keeps working on the new topology: legality finds only the constrained edges, marks equal the oracle's
        flags, nill = ctx.legality()
        oflags, onill = oracle.legality(xy, None, edges)
        assert np.array_equal(flags, oflags) and nill == onill == res["n_constrained_illegal"] and nill0 > nill
        valid, sign, ninv = ctx.trial_move_validate(np.zeros_like(xy))
        assert ninv == 0 and (sign > 0).all()
        if len(m.segs):
            prm = ctx.indicator_init()
            d, mx = ctx.distance_update()
            marks, le, counts = ctx.mark_elements(run_distance=False)
            od = oracle.distance_2d(xy, m.segs)
            prop = oracle.Params.from_buffer_copy(bytes(prm)).distProportion
            omarks, ole, oc3 = oracle.mark(xy, od, cells, perm, oracle.Params.from_buffer_copy(bytes(prm)), prop * oracle.distance_max(od))
            assert np.array_equal(d, od) and np.array_equal(marks, omarks) and np.array_equal(le, ole) and tuple(counts) == tuple(oc3)
        # a second call has nothing to do; an inverted cell is refused
        again = ctx.restore_delaunay(vertex_id=m.vid if use_vid else None)
        assert again["n_flips"] == 0 and again["converged"] == 1 and again["rounds"] == 0
        bad = xy.copy(); c = cells[m.nc // 2]
        bad[c[0]] = bad[c[1]] + (bad[c[1]] - bad[c[0]]) + (bad[c[2]] - bad[c[1]]) * 2.0
        ctx.upload_coords(bad)
        with pytest.raises(Exception, match="positively oriented"):
            ctx.restore_delaunay()


def test_distance_wide_path_equidistant_interface(gpu_ctx_factory, oracle):
    """Vertices about equally far from a long stretch of interface (the centre of a closed curve): the warp's candidate list
    overflows and its vertices are served by k_distance2d_wide, one warp per vertex.  The facets need not be mesh edges for the
    distance field, so the case is built directly: 1500 segments on a circle, query vertices at the centre, on the circle,
    and everywhere else.  Distances and Distance::maximum must carry the brute-force bits, cold, warm and after a move."""
    rng = np.random.default_rng(23)
    ncirc = 1500
    t = np.linspace(0.0, 2.0 * np.pi, ncirc, endpoint=False)
    circ = np.stack([0.5 + 0.3 * np.cos(t), 0.5 + 0.3 * np.sin(t)], 1)
    centre = 0.5 + 1e-4 * rng.standard_normal((1024, 2))
    ring2 = 0.5 + 0.3 * (1.0 + 1e-3 * rng.standard_normal((512, 1))) * np.stack([np.cos(rng.uniform(0, 6.3, 512)), np.sin(rng.uniform(0, 6.3, 512))], 1)
    anywhere = rng.uniform(0.0, 1.0, (2048, 2))
    xy = np.ascontiguousarray(np.concatenate([circ, centre, anywhere, ring2, np.full((37, 2), 0.5)]))
    segs = np.stack([np.arange(ncirc), (np.arange(ncirc) + 1) % ncirc], 1).astype(np.int32)
    nv = xy.shape[0]
    cells = np.array([[0, 1, 2]], np.int32); perm = np.array([0x24], np.uint8)
    ctx = gpu_ctx_factory(2)
    ctx.upload_mesh(xy, cells, perm, np.zeros((0, 4), np.int32), segs)
    ref = oracle.distance_2d(xy, segs)
    d, mx = ctx.distance_update()                                       # cold
    assert np.array_equal(d, ref) and mx == oracle.distance_max(ref)
    d, mx = ctx.distance_update()                                       # warm
    assert np.array_equal(d, ref) and mx == oracle.distance_max(ref)
    tm_names = None
    ctx.timing_enable(True); ctx.timing_reset(); ctx.distance_update(); tm_names = ctx.timing_get(); ctx.timing_enable(False)
    assert "distance2d_wide" in tm_names
    s = 0.05 * rng.standard_normal((nv, 2)); s[:ncirc] *= 0.02            # the interface wobbles, the queries move a lot
    ctx.move_vertices(s)
    xy2 = oracle.move(xy, s)
    ref2 = oracle.distance_2d(xy2, segs)
    d2, mx2 = ctx.distance_update()
    assert np.array_equal(d2, ref2) and mx2 == oracle.distance_max(ref2)
    assert (ref[ncirc:ncirc + 1024] > 0.29).all()                          # the centre really is equidistant
