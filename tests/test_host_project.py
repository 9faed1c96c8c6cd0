"""CPU run of the PRODUCT DG-P1 projection phases (dune-mmesh_b200/csrc/project.cuh: tile of pair slots, transposed
columns, dense list of non-empty polygons, per-pair partial sums) against the oracle, thread by thread without a GPU.
Without FMA the DOFs must be bit-identical to the oracle (same arithmetic, same accumulation order); with FMA in the
quadrature they must agree to 1e-12 and the piece count / covered area must not change at all."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "native", "_build", "libhost_project.so")


@pytest.fixture(scope="module")
def hp():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           "-I", os.path.join(ROOT, "dune-mmesh_b200", "csrc"),
                           os.path.join(ROOT, "tests", "native", "host_project.cc"), "-o", SO])
    L = C.CDLL(SO)

    def run(nt, fma, ragged, xy, cells, perm, new_off, new_cell, old_xy, old_idx, dofs_old, n_new_cells, eps=1e-14, drop=1e-8, translate=0):
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        xy = np.ascontiguousarray(xy, np.float64)
        cv = [np.ascontiguousarray(cells[:, k], np.int32) for k in range(3)]
        perm = np.ascontiguousarray(perm, np.uint8)
        new_off = np.ascontiguousarray(new_off, np.int64); new_cell = np.ascontiguousarray(new_cell, np.int32)
        old_xy = np.ascontiguousarray(old_xy, np.float64); old_idx = np.ascontiguousarray(old_idx, np.int32)
        dofs_old = np.ascontiguousarray(dofs_old, np.float64)
        out = np.zeros((n_new_cells, 3))
        mass, cov, npc = C.c_double(), C.c_double(), C.c_long()
        L.hp_project(C.c_int(nt), C.c_int(fma), C.c_int(ragged), P(xy), P(cv[0]), P(cv[1]), P(cv[2]), P(perm), P(new_off), P(new_cell),
                     C.c_long(new_cell.shape[0]), P(old_xy), P(old_idx), P(dofs_old), P(out), C.c_double(eps), C.c_double(drop),
                     C.c_int(translate), C.byref(mass), C.byref(cov), C.byref(npc))
        return out, mass.value, cov.value, npc.value

    return run


def _self_neighbour_case(n=37, seed=30):
    from mmesh_b200 import synth
    m = synth.config1(n=n)
    s = synth.shifts_uniform(m, 0.3 * m.h, seed)
    new_off, new_cell, old_xy, old_idx = synth.pairs_self_and_neighbours(m, m.xy - s)
    k = np.stack([np.take_along_axis(m.cells, ((m.perm >> (2 * i)) & 3)[:, None].astype(np.int64), 1)[:, 0] for i in range(3)], 1)
    pk = (m.xy - s)[k]
    dofs = np.ascontiguousarray(np.sin(np.pi * pk[..., 0]) * np.sin(np.pi * pk[..., 1]) + 2.0)
    return m, new_off, new_cell, old_xy, old_idx, dofs


@pytest.mark.parametrize("nt", [128, 256])
@pytest.mark.parametrize("translate,drop", [(0, 1e-8), (1, 0.0), (0, 1e-5)])
def test_tile_projection_matches_oracle_bitwise(hp, oracle, nt, translate, drop):
    m, new_off, new_cell, old_xy, old_idx, dofs = _self_neighbour_case()
    prm = oracle.Params.default(); prm.clip_translate = translate; prm.drop_area = drop
    ref, rmass, rcov, rnpc = oracle.project(m.xy, m.cells, m.perm, new_off, new_cell, old_xy, old_idx, dofs, m.nc, 3, prm)
    for ragged in (0, 1):            # transposed columns (uniform tile) and the identity map must give the same bits
        got, mass, cov, npc = hp(nt, 0, ragged, m.xy, m.cells, m.perm, new_off, new_cell, old_xy, old_idx, dofs, m.nc, drop=drop, translate=translate)
        assert np.array_equal(got, ref.reshape(-1, 3))
        assert npc == rnpc and abs(cov - rcov) <= 1e-13 * abs(rcov) and abs(mass - rmass) <= 1e-13 * abs(rmass)
    assert rnpc > 0
    # FMA in the quadrature: same pieces, DOFs within 1e-12 of the oracle
    got, mass, cov, npc = hp(nt, 1, 0, m.xy, m.cells, m.perm, new_off, new_cell, old_xy, old_idx, dofs, m.nc, drop=drop, translate=translate)
    assert npc == rnpc and cov == hp(nt, 0, 0, m.xy, m.cells, m.perm, new_off, new_cell, old_xy, old_idx, dofs, m.nc, drop=drop, translate=translate)[2]
    scale = np.abs(ref).max()
    assert np.abs(got - ref.reshape(-1, 3)).max() <= 1e-12 * scale
    assert abs(mass - rmass) <= 1e-13 * abs(rmass)


@pytest.mark.parametrize("nt", [128, 256])
def test_tile_projection_ragged_csr(hp, oracle, nt):
    """pair lists of every length (0 .. 700 pairs per new cell: several chunks per tile, cells straddling chunks)"""
    m, new_off, new_cell, old_xy, old_idx, dofs = _self_neighbour_case(n=21)
    rng = np.random.default_rng(5)
    nnew = 300
    cnt = rng.integers(0, 9, nnew)
    cnt[17] = 700; cnt[18] = 0; cnt[40] = 129; cnt[299] = 300
    off = np.zeros(nnew + 1, np.int64); off[1:] = np.cumsum(cnt)
    cells_sel = np.sort(rng.choice(m.nc, nnew, replace=False)).astype(np.int32)
    P = int(off[-1])
    # children: the cell's own old copy, its neighbours' and random far cells (empty intersections)
    src = rng.integers(0, old_idx.shape[0], P)
    near = rng.random(P) < 0.6
    rowstart = np.repeat(new_off[cells_sel], cnt)
    src[near] = (rowstart + rng.integers(0, 4, P))[near]
    oxy, oidx = old_xy.reshape(-1, 6)[src], old_idx[src]
    prm = oracle.Params.default()
    ref, rmass, rcov, rnpc = oracle.project(m.xy, m.cells, m.perm, off, cells_sel, oxy, oidx, dofs, m.nc, 3, prm)
    got, mass, cov, npc = hp(nt, 0, 0, m.xy, m.cells, m.perm, off, cells_sel, oxy, oidx, dofs, m.nc)
    assert np.array_equal(got[cells_sel], ref.reshape(-1, 3)[cells_sel])
    assert npc == rnpc and rnpc > 100


def test_tile_projection_flip_case_conservative(hp, oracle):
    from projection_cases import diagonal_flip_case
    case = diagonal_flip_case(n=24, seed=7)
    rng = np.random.default_rng(2)
    dofs = rng.uniform(1, 2, (case["n_old"], 3))
    prm = oracle.Params.default(); prm.clip_translate = 1; prm.drop_area = 0.0
    nc = case["cells"].shape[0]
    ref, rmass, rcov, rnpc = oracle.project(case["xy"], case["cells"], case["perm"], case["new_off"], case["new_cell"], case["old_xy"],
                                            case["old_idx"], dofs, nc, 3, prm)
    got, mass, cov, npc = hp(128, 0, 0, case["xy"], case["cells"], case["perm"], case["new_off"], case["new_cell"], case["old_xy"],
                             case["old_idx"], dofs, nc, drop=0.0, translate=1)
    assert np.array_equal(got, ref.reshape(-1, 3)) and npc == rnpc
    assert abs(cov - case["new_vol"].sum()) <= 1e-13 * cov


def test_tile_projection_generic_path(hp, oracle):
    """degenerate pairs (NaN / huge coordinates, zero-area cells, clip edges of zero length): whatever the tile's capacity
    test decides, slots or generic path, the result equals the oracle's reference-structured loop"""
    m, new_off, new_cell, old_xy, old_idx, dofs = _self_neighbour_case(n=15)
    oxy = old_xy.reshape(-1, 6).copy()
    rng = np.random.default_rng(11)
    k = rng.choice(oxy.shape[0], 200, replace=False)
    oxy[k[:50]] += rng.uniform(-3e-14, 3e-14, (50, 6))
    oxy[k[50:100], 4:6] = 0.5 * (oxy[k[50:100], 0:2] + oxy[k[50:100], 2:4])      # zero-area old cells
    oxy[k[100:150]] *= 1e8                                                          # far away / huge
    oxy[k[150:200], 2:4] = oxy[k[150:200], 0:2]                                    # repeated corner
    prm = oracle.Params.default()
    ref, rmass, rcov, rnpc = oracle.project(m.xy, m.cells, m.perm, new_off, new_cell, oxy, old_idx, dofs, m.nc, 3, prm)
    got, mass, cov, npc = hp(128, 0, 0, m.xy, m.cells, m.perm, new_off, new_cell, oxy, old_idx, dofs, m.nc)
    assert npc == rnpc
    fin = np.isfinite(ref.reshape(-1, 3)).all(axis=1)
    assert np.array_equal(got[fin], ref.reshape(-1, 3)[fin]) and np.array_equal(np.isfinite(got).all(axis=1), fin)
