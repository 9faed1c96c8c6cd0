"""Parity at BASELINE.json's full size (config 3: 16 777 216 triangles): full-array equality where the oracle is fast
enough with all host threads, seeded samples where it is not (brute-force distance), and size-independent
properties (identity projection reproduces the DOFs and conserves mass)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    from mmesh_b200 import synth
    m = synth.config3()
    s = synth.shifts_uniform(m, 0.3 * m.h, 30)
    return m, s


def test_full_size_parity(big, gpu_ctx_factory, oracle):
    from mmesh_b200 import synth
    m, s = big
    oracle.set_threads(os.cpu_count() or 1)
    try:
        ctx = gpu_ctx_factory(2)
        ctx.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs)
        prm = ctx.indicator_init()
        oprm = oracle.Params.from_buffer_copy(bytes(prm))
        # --- distance: EVERY vertex against the oracle (the culled variant returns the brute-force bits; it is pinned
        #     against the reference's brute-force loop on a seeded sample here and on whole meshes in the CPU suite) ---
        d, mx = ctx.distance_update()
        rd = oracle.distance_2d_culled(m.xy, m.segs)
        assert np.array_equal(d, rd)
        rng = np.random.default_rng(99)
        near = np.argsort(d)[:: max(1, m.nv // 1500)][:1500]
        sample = np.unique(np.concatenate([rng.integers(0, m.nv, 2500), near, [int(np.argmax(d))]])).astype(np.int32)
        assert np.array_equal(rd[sample], oracle.distance_2d_subset(m.xy, sample, m.segs))
        assert mx == d.max() and np.all(d[m.segs.ravel()] == 0.0)
        # second call starts from the hints of the first: identical field
        d2, mx2 = ctx.distance_update()
        assert np.array_equal(d, d2) and mx2 == mx
        # --- marks and longest edges of ALL 16.8 M cells (tuned parameters: refinement and coarsening marks both occur) ---
        prm.maxH *= 0.5
        ctx.set_params(prm)
        oprm = oracle.Params.from_buffer_copy(bytes(prm))
        marks, le, counts = ctx.mark_elements(run_distance=False)
        rm, rle, rc = oracle.mark(m.xy, rd, m.cells, m.perm, oprm, prm.distProportion * oracle.distance_max(rd))
        assert np.array_equal(marks, rm) and np.array_equal(le, rle) and tuple(counts) == tuple(rc)
        assert rc[0] > 1000 and rc[1] > 1000
        # --- trial validity and legality: full arrays ---
        valid, sign, ninv = ctx.trial_move_validate(s)
        rv, rs, rn = oracle.trial_validate(m.xy, s, m.cells)
        assert np.array_equal(valid, rv) and np.array_equal(sign, rs) and ninv == rn
        flags, nill = ctx.legality()
        rf, rnill = oracle.legality(m.xy, None, m.edges)
        assert np.array_equal(flags, rf) and nill == rnill
        # --- identity projection: every cell is its own old child => DOFs reproduced, mass conserved ---
        ident = np.arange(m.nc, dtype=np.int32)
        ctx.upload_pairs(np.arange(m.nc + 1, dtype=np.int64), ident, synth.cached_corners(m, m.xy, ident), ident)
        k = np.stack([np.take_along_axis(m.cells, ((m.perm >> (2 * i)) & 3)[:, None].astype(np.int64), 1)[:, 0] for i in range(3)], 1)
        pk = m.xy[k]
        dofs = np.sin(np.pi * pk[..., 0]) * np.sin(np.pi * pk[..., 1]) + 2.0
        got, mass, cov, npc = ctx.project(3, dofs)
        assert npc == m.nc and abs(cov - 2.0) < 1e-12
        assert np.max(np.abs(got - dofs)) <= 1e-12 * np.max(np.abs(dofs))
        vol = np.abs(synth.cross2(pk[:, 1] - pk[:, 0], pk[:, 2] - pk[:, 0])) * 0.5
        mass_old = float(np.sum(vol * dofs.mean(axis=1), dtype=np.longdouble))
        assert abs(mass - mass_old) <= 1e-13 * abs(mass_old)
        p0 = dofs[:, 0].copy()
        got0, mass0, _, _ = ctx.project(1, p0)
        assert np.array_equal(got0, p0 * (vol / vol))          # weight |T|/|K| == 1 exactly for the identity cut
        # --- real move ---
        ctx.move_vertices(s)
        assert np.array_equal(ctx.download_coords(), m.xy + s)
    finally:
        oracle.set_threads(1)


def test_full_size_projection_vs_oracle(big, gpu_ctx_factory, oracle):
    """Benchmark pair set (self + face neighbours): DOFs on sampled ranges, piece count and mass on the full set."""
    from mmesh_b200 import synth
    m, s = big
    oracle.set_threads(os.cpu_count() or 1)
    try:
        new_off, new_cell, old_xy, old_idx = synth.pairs_self_and_neighbours(m, m.xy - s)
        ctx = gpu_ctx_factory(2)
        ctx.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs)
        ctx.upload_pairs(new_off, new_cell, old_xy, old_idx)
        prm = ctx.get_params()
        oprm = oracle.Params.from_buffer_copy(bytes(prm))
        rng = np.random.default_rng(3)
        dofs = rng.uniform(0.5, 1.5, (m.nc, 3))
        got, mass, cov, npc = ctx.project(3, dofs)
        ref, rmass, rcov, rnpc = oracle.project(m.xy, m.cells, m.perm, new_off, new_cell, old_xy, old_idx, dofs, m.nc, 3, oprm)
        assert npc == rnpc
        assert abs(mass - rmass) <= 1e-13 * abs(rmass) and abs(cov - rcov) <= 1e-13 * rcov
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(got - ref)) <= 1e-12 * scale
        # the pipelined host-buffer step (chunked DOF upload / projection / download) returns the very same DOFs
        out = ctx.adapt_step_host(s, 3, dofs)
        assert np.array_equal(out["dofs_new"], got)
        assert out["result"].n_pieces == rnpc and abs(out["result"].mass_new - rmass) <= 1e-13 * abs(rmass)
        rv, rs, rn = oracle.trial_validate(m.xy, s, m.cells)
        assert np.array_equal(out["valid_fp"], rv) and np.array_equal(out["orient_sign"], rs)
    finally:
        oracle.set_threads(1)


def test_mass_conservation_at_16M(gpu_ctx_factory, oracle):
    """North-star: mass conserved to 1e-13 on a 16 M-cell mesh.  Exactly conservative set-up (every quad of a jittered
    2896^2 lattice re-triangulated with the other diagonal => 16.8 M new cells, 33.5 M pairs, cell area 6e-8).
    The reference arithmetic (absolute coordinates in lineIntersectionPoint, area cut-off 1e-8) and the scale-aware
    mode (clip relative to the new cell, no cut-off) are both measured; the latter must meet 1e-13."""
    from mmesh_b200 import synth
    c = synth.flip_case(2896)
    assert c["cells"].shape[0] == 2 * 2896 * 2896
    ctx = gpu_ctx_factory(2)
    ctx.upload_mesh(c["xy"], c["cells"], c["perm"], None, None)
    ctx.upload_pairs(c["new_off"], c["new_cell"], c["old_xy"], c["old_idx"])
    u0 = np.sin(3.0 * c["old_centroid"][:, 0]) + c["old_centroid"][:, 1] ** 2 + 1.0
    m0 = float(np.sum(c["old_vol"] * u0, dtype=np.longdouble))
    pk = c["old_xy"].reshape(-1, 3, 2)
    first = np.arange(0, c["old_idx"].shape[0], 1)
    # nodal DG-P1 values of a smooth function on the cached corners of every old cell
    firstpair = np.full(c["n_old"], -1, np.int64)
    firstpair[c["old_idx"][::-1]] = np.arange(c["old_idx"].shape[0])[::-1]
    corners = pk[firstpair]
    u1 = np.sin(3.0 * corners[..., 0]) + corners[..., 1] ** 2 + 1.0
    m1 = float(np.sum(c["old_vol"] * u1.mean(axis=1), dtype=np.longdouble))
    errs = {}
    for translate, drop in ((0, 1e-8), (1, 0.0)):
        prm = ctx.get_params()
        prm.clip_translate, prm.drop_area = translate, drop
        ctx.set_params(prm)
        _, mass, cov, npc = ctx.project(1, u0)
        _, massp1, _, _ = ctx.project(3, u1)
        errs[translate] = (abs(mass - m0) / m0, abs(massp1 - m1) / m1, abs(cov - 1.0), npc)
    print("mass conservation at 16.8M cells: reference arithmetic", errs[0], " scale-aware", errs[1])
    assert errs[1][0] <= 1e-13 and errs[1][1] <= 1e-13 and errs[1][2] <= 1e-13
    assert errs[1][3] == 2 * c["cells"].shape[0]                    # every (old, new) pair contributes exactly one piece
    assert errs[0][0] <= 1e-3                                       # reference constants: pieces below the absolute 1e-8 area cut-off are lost
    # the oracle agrees with the device on a sample of new cells in the scale-aware mode
    oprm = oracle.Params.from_buffer_copy(bytes(ctx.get_params()))
    sel = np.arange(0, 200000, dtype=np.int32)
    got, _, _, _ = ctx.project(1, u0)
    ref, _, _, _ = oracle.project(c["xy"], c["cells"], c["perm"], c["new_off"][:200001], sel, c["old_xy"][:400000], c["old_idx"][:400000],
                                  u0, c["cells"].shape[0], 1, oprm)
    assert np.array_equal(got[:200000], ref[:200000])


def test_config2_rotating_circle_100_steps(gpu_ctx_factory, oracle):
    """BASELINE config 2 at its named size: 708^2 lattice (1 002 528 triangles), closed circular interface, 100 move / adapt
    steps with DG-P1 projection of the self + face-neighbour pairs.  Checked against the oracle:
      * every step: the counts of the fused step, on the host-tracked coordinates (moveVertices is one IEEE add per
        component, so the CPU follows the GPU state bit for bit), for steps 1, 50 and 100 all per-cell / per-edge fields;
      * steps 1, 50, 100: the projected DOFs of a seeded sample of 20 000 new cells from the DOFs the GPU held before
        the step (the DOF history itself is 100 chained projections, which only the step-wise check can pin);
      * after step 100: coordinates, distances of all vertices, marks of all cells."""
    from mmesh_b200 import synth
    oracle.set_threads(os.cpu_count() or 1)
    try:
        m = synth.config2()
        assert m.nc == 1002528
        ctx = gpu_ctx_factory(2)
        ctx.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs)
        prm = ctx.indicator_init()
        prm.clip_translate, prm.drop_area = 1, 0.0
        ctx.set_params(prm)
        oprm = oracle.Params.from_buffer_copy(bytes(prm))
        k = np.stack([np.take_along_axis(m.cells, ((m.perm >> (2 * i)) & 3)[:, None].astype(np.int64), 1)[:, 0] for i in range(3)], 1)
        pk = m.xy[k]
        dofs = np.ascontiguousarray(np.sin(np.pi * pk[..., 0]) * np.sin(np.pi * pk[..., 1]) + 2.0)
        xy = m.xy.copy()
        rng = np.random.default_rng(7)
        sample = np.sort(rng.choice(m.nc, 20000, replace=False)).astype(np.int32)
        dtheta = 2 * np.pi / 1000
        import copy
        mm = copy.copy(m)
        new_off, new_cell, _, old_idx = synth.pairs_self_and_neighbours(m, m.xy)      # the pair topology is the same every step
        for step in range(1, 101):
            mm.xy = xy
            s = synth.shifts_rotation(mm, dtheta, r_in=0.35, r_out=0.5)
            moved = oracle.move(xy, s)
            # new cells = the moved cells, old children = pre-move copies of the cell and its face neighbours
            old_xy = synth.cached_corners(m, xy, old_idx)
            ctx.upload_coords(moved)                       # adapt(handle) sees the new cells; the step below moves them back first
            ctx.upload_pairs(new_off, new_cell, old_xy, old_idx)
            got, mass, cov, npc = ctx.project(3, dofs)
            if step in (1, 50, 100):
                rows = np.concatenate([np.arange(new_off[c], new_off[c + 1]) for c in sample])
                off = np.concatenate([[0], np.cumsum(np.diff(new_off)[sample])]).astype(np.int64)
                ref, _, _, _ = oracle.project(moved, m.cells, m.perm, off, sample, old_xy.reshape(-1, 6)[rows], old_idx[rows], dofs, m.nc, 3, oprm)
                ref = ref.reshape(-1, 3)
                assert np.max(np.abs(got[sample] - ref[sample])) <= 1e-12 * np.max(np.abs(ref[sample]))
            ctx.upload_coords(xy)
            res = ctx.adapt_step(s, ncomp=0)               # marks + validity on the old state, moveVertices, legality
            rv, rs, rn = oracle.trial_validate(xy, s, m.cells)
            assert res.n_invalid == rn == 0 and res.moved == 1
            if step in (1, 50, 100):
                f = ctx.download_fields(0)
                rd = oracle.distance_2d_culled(xy, m.segs)
                rm, rle, rc = oracle.mark(xy, rd, m.cells, m.perm, oprm, prm.distProportion * oracle.distance_max(rd))
                rf, rnill = oracle.legality(moved, None, m.edges)
                assert np.array_equal(f["marks"], rm) and np.array_equal(f["longest"], rle) and (res.n_refine, res.n_coarsen) == tuple(rc)
                assert np.array_equal(f["valid_fp"], rv) and np.array_equal(f["orient_sign"], rs)
                assert np.array_equal(f["edge_flags"], rf) and res.n_illegal == rnill
                assert res.max_dist == oracle.distance_max(rd)
            xy, dofs = moved, got
        assert np.array_equal(ctx.download_coords(), xy)
        d, mx = ctx.distance_update()
        assert np.array_equal(d, oracle.distance_2d_culled(xy, m.segs))
        # (the self + face-neighbour pair set does not tile the rotated cells, so the field decays from step to step: only
        # finiteness and positivity are asserted for the chained result; every single projection was pinned above)
        assert np.isfinite(dofs).all() and 0.0 < dofs.mean() < 3.0
    finally:
        oracle.set_threads(1)


def test_config4_degenerate_lattice_full_sign_arrays(gpu_ctx_factory, oracle):
    """BASELINE config 4 at its named size: 1025 x 1025 exact lattice (every quad cocircular, every row collinear) and the
    same lattice with every coordinate perturbed by U(-1e-15, 1e-15): orientation of every lattice triangle and of 2^20
    collinear triples, in-circle of every quad with both diagonals -- the full sign arrays, bit for bit against the exact
    oracle (whose signs are pinned to exact rational arithmetic in the CPU suite)."""
    from mmesh_b200 import synth
    oracle.set_threads(os.cpu_count() or 1)
    try:
        ctx = gpu_ctx_factory(2)
        n = 1024
        for perturb in (0.0, 1e-15):
            m = synth.config4(n=n, perturb=perturb)
            assert m.nv == 1025 * 1025
            tri = m.xy[m.cells].reshape(-1, 6)
            sg, nex = ctx.predicate_batch(0, tri)
            assert np.array_equal(sg, oracle.orient2d_batch(tri))
            assert (sg > 0).all() if perturb == 0.0 else True
            # collinear triples (i, i + k, i + 2k) along lattice rows
            rng = np.random.default_rng(4)
            row = rng.integers(0, n + 1, 1 << 20); kk = rng.integers(1, n // 2, 1 << 20); i0 = (rng.random(1 << 20) * (n + 1 - 2 * kk)).astype(np.int64)
            vmap = np.lexsort((m.lat[:, 0], m.lat[:, 1]))           # vertex index of lattice point (i, j) = vmap[j * (n + 1) + i]
            tri3 = np.stack([m.xy[vmap[row * (n + 1) + i0 + t * kk]] for t in range(3)], 1).reshape(-1, 6)
            sg3, nex3 = ctx.predicate_batch(0, tri3)
            r3 = oracle.orient2d_batch(tri3)
            assert np.array_equal(sg3, r3)
            if perturb == 0.0:
                assert (r3 == 0).all()
            else:
                assert len(set(np.unique(r3).tolist())) == 3 and nex3 > 0
            # in-circle of every quad with both diagonals: (a, b, c; d) for every inner edge of the mesh, and the other diagonal
            e = m.edges[m.edges[:, 3] >= 0]
            q1 = m.xy[e].reshape(-1, 8)
            q2 = m.xy[e[:, [2, 3, 1, 0]]].reshape(-1, 8)            # the quad's other diagonal (c, d) with apices b, a
            for q in (q1, q2):
                si, nei = ctx.predicate_batch(1, q)
                ri = oracle.incircle_batch(q)
                assert np.array_equal(si, ri)
            if perturb > 0:
                assert nei > 0
    finally:
        oracle.set_threads(1)


def test_config5_3d_8m_tets(gpu_ctx_factory, oracle):
    """BASELINE config 5 at its named size: 111^3 lattice, 7 986 000 Kuhn tetrahedra, interface = the facets on z = 0.5:
    orient3d validity of ALL cells for the trial move, sampled distances (the 3-D reference loop is brute force), marks of
    ALL cells, the real move, and the P0 trace / skeleton transfer on every interface facet."""
    from mmesh_b200 import synth
    oracle.set_threads(os.cpu_count() or 1)
    try:
        m = synth.lattice3d(110)
        assert m.nc == 110 ** 3 * 6
        ctx = gpu_ctx_factory(3)
        ctx.upload_mesh(m.xyz, m.cells, m.perm, None, m.tris)
        from mmesh_b200 import capi
        # shifts U(-0.2h, 0.2h) on the 0.15h-jittered lattice invert a few hundred tetrahedra: that is the
        # GridError of mmesh.hh:1104-1107 (no vertex removal in 3-D); the error carries the full flag arrays
        big = synth.shifts_uniform(m, 0.2 * m.h, 50)
        rv, rs, rn = oracle.trial_validate(m.xyz, big, m.cells, dim=3)
        assert rn > 0
        with pytest.raises(capi.GridError) as ei:
            ctx.trial_move_validate(big)
        assert np.array_equal(ei.value.valid_fp, rv) and np.array_equal(ei.value.orient_sign, rs) and ei.value.n_invalid == rn
        s = synth.shifts_uniform(m, 0.1 * m.h, 50)                 # a movement the reference accepts
        valid, sign, ninv = ctx.trial_move_validate(s)
        rv, rs, rn = oracle.trial_validate(m.xyz, s, m.cells, dim=3)
        assert np.array_equal(valid, rv) and np.array_equal(sign, rs) and ninv == rn == 0
        d, mx = ctx.distance_update()
        rng = np.random.default_rng(55)
        sample = np.unique(np.concatenate([rng.integers(0, m.nv, 3000), [int(np.argmax(d)), int(np.argmin(d))]]))
        rd = oracle.distance_3d_subset(m.xyz, sample, m.tris)      # brute force of the sample against all 24 200 facets
        assert np.array_equal(d[sample], rd) and mx == d.max()
        prm = ctx.get_params()
        prm.minH, prm.maxH, prm.factor = 0.25 * m.h, 1.3 * m.h, 1.5
        ctx.set_params(prm)
        marks, le, counts = ctx.mark_elements(run_distance=False)
        rm, rle, rc = oracle.mark(m.xyz, d, m.cells, m.perm, oracle.Params.from_buffer_copy(bytes(prm)), prm.distProportion * mx, dim=3)
        assert np.array_equal(marks, rm) and np.array_equal(le, rle) and tuple(counts) == tuple(rc) and 0 < rc[0] < m.nc
        ctx.move_vertices(s)
        assert np.array_equal(ctx.download_coords(), oracle.move(m.xyz, s))
        nf = m.tris.shape[0]
        assert nf == 110 * 110 * 2
        inside = rng.integers(0, m.nc, nf).astype(np.int32)
        outside = rng.integers(-1, m.nc, nf).astype(np.int32)
        cen = m.xyz[m.cells].mean(axis=1)
        u = (cen ** 2).sum(axis=1)                                 # P0 data x^2 + y^2 + z^2 at the centroids
        ug = rng.uniform(size=nf)
        tin, tout, sk = ctx.trace_skeleton_p0(inside, outside, u, ug)
        assert np.array_equal(tin, u[inside]) and np.array_equal(tout, np.where(outside >= 0, u[np.maximum(outside, 0)], 0.0)) and np.array_equal(sk, ug)
    finally:
        oracle.set_threads(1)
