"""CPU test of the PRODUCT host function buildProjectionPairs (dune-mmesh_b200/host/pairs.hh): the (new cell, old children)
CSR of MMesh::adapt(handle) derived from an old and a new snapshot, on generated insertion / removal cases.  Checked
against a brute-force overlap search (oracle intersection volumes), the tiling property sum |old ^ new| = |new|
(movement.cc:89-94) and the transfer pins (constants / linears reproduced, mass conserved) through the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from adapt_cases import adapt_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "native", "_build", "libhost_pairs.so")


@pytest.fixture(scope="module")
def build_pairs():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "native", "host_pairs.cc"), "-o", SO])
    L = C.CDLL(SO)
    L.hpair_build.restype = C.c_long

    def run(case, comp_old=None):
        old = case["old"]
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        xo, co, po = np.ascontiguousarray(old.xy), np.ascontiguousarray(old.cells, np.int32), np.ascontiguousarray(old.perm, np.uint8)
        ido = np.ascontiguousarray(case["old_id"], np.int64)
        xn, cn, pn = case["new_xy"], case["new_cells"], np.ascontiguousarray(case["new_perm"], np.uint8)
        idn = np.ascontiguousarray(case["new_id"], np.int64)
        nnew, ncomp = C.c_long(), C.c_long()
        err = C.create_string_buffer(256)
        comp = None if comp_old is None else np.ascontiguousarray(comp_old, np.int32)
        npairs = L.hpair_build(C.c_long(old.nv), P(xo), P(ido), C.c_long(old.nc), P(co), P(po), C.c_long(xn.shape[0]), P(xn), P(idn),
                               C.c_long(cn.shape[0]), P(cn), P(pn), None if comp is None else P(comp), C.byref(nnew), C.byref(ncomp), err, 256)
        if npairs < 0:
            raise RuntimeError(err.value.decode())
        out = dict(new_off=np.zeros(nnew.value + 1, np.int64), new_cell=np.zeros(nnew.value, np.int32), old_xy=np.zeros((npairs, 6)),
                   old_idx=np.zeros(npairs, np.int32), comp_old=np.zeros(old.nc, np.int32), comp_new=np.zeros(cn.shape[0], np.int32),
                   ncomp=ncomp.value)
        L.hpair_get(P(out["new_off"]), P(out["new_cell"]), P(out["old_xy"]), P(out["old_idx"]), P(out["comp_old"]), P(out["comp_new"]))
        return out

    return run


@pytest.mark.parametrize("seed,n_ins,n_rem", [(3, 6, 6), (4, 25, 0), (5, 0, 20), (6, 40, 40)])
def test_pairs_from_snapshots(build_pairs, oracle, seed, n_ins, n_rem):
    case = adapt_case(n=14, seed=seed, n_insert=n_ins, n_remove=n_rem)
    old = case["old"]
    pr = build_pairs(case)
    # isNew / vanished exactly as generated
    assert np.array_equal(np.sort(pr["new_cell"]), np.nonzero(case["created"])[0])
    assert np.array_equal(pr["comp_old"] >= 0, case["vanished"])
    assert (np.diff(pr["new_cell"]) > 0).all()                        # element order (mmesh.hh:1143)
    # children of a new cell = all vanished cells of its component, lowest index first (seed of MMeshConnectedComponent)
    for i, K in enumerate(pr["new_cell"]):
        ch = pr["old_idx"][pr["new_off"][i]:pr["new_off"][i + 1]]
        comp = pr["comp_new"][K]
        members = np.nonzero(pr["comp_old"] == comp)[0]
        assert np.array_equal(np.sort(ch), members) and ch[0] == members.min() and len(set(ch.tolist())) == len(ch)
    # cached corners = id-sorted corners of the old cell
    k = np.stack([np.take_along_axis(old.cells, ((old.perm >> (2 * t)) & 3)[:, None].astype(np.int64), 1)[:, 0] for t in range(3)], 1)
    assert np.array_equal(pr["old_xy"].reshape(-1, 3, 2), old.xy[k[pr["old_idx"]]])
    # brute force: every vanished cell that overlaps a new cell is among its children, and the children tile the new cell
    van = np.nonzero(case["vanished"])[0]
    newc = case["new_xy"][case["new_cells"]]
    for i, K in enumerate(pr["new_cell"]):
        new6 = newc[K].reshape(6)
        area = abs(0.5 * ((new6[2] - new6[0]) * (new6[5] - new6[1]) - (new6[4] - new6[0]) * (new6[3] - new6[1])))
        ch = set(pr["old_idx"][pr["new_off"][i]:pr["new_off"][i + 1]].tolist())
        tot = 0.0
        for o in van:
            v = oracle.intersection_volume(old.xy[k[o]].reshape(6), new6)
            if v > 1e-13:
                assert o in ch
            if o in ch:
                tot += v
        assert abs(tot - area) <= 1e-12 * max(area, 1e-30) + 1e-15
    # transfer through the oracle with the built pairs: constants and linears reproduced, mass conserved
    prm = oracle.Params.default(); prm.clip_translate = 1; prm.drop_area = 0.0
    pk = old.xy[k]
    lin = 1.0 + 2.0 * pk[..., 0] - 3.0 * pk[..., 1]                   # DG-P1 nodal values of a linear function
    ncn = case["new_cells"].shape[0]
    got, mass, cov, npc = oracle.project(case["new_xy"], case["new_cells"], case["new_perm"], pr["new_off"], pr["new_cell"], pr["old_xy"],
                                         pr["old_idx"], lin, ncn, 3, prm)
    kn = np.stack([np.take_along_axis(case["new_cells"], ((case["new_perm"] >> (2 * t)) & 3)[:, None].astype(np.int64), 1)[:, 0] for t in range(3)], 1)
    pkn = case["new_xy"][kn]
    want = 1.0 + 2.0 * pkn[..., 0] - 3.0 * pkn[..., 1]
    sel = pr["new_cell"]
    assert np.abs(got.reshape(-1, 3)[sel] - want[sel]).max() <= 1e-11
    old_area = np.abs(0.5 * ((pk[:, 1, 0] - pk[:, 0, 0]) * (pk[:, 2, 1] - pk[:, 0, 1]) - (pk[:, 2, 0] - pk[:, 0, 0]) * (pk[:, 1, 1] - pk[:, 0, 1])))
    mass_old = (old_area * lin.mean(axis=1))[case["vanished"]].sum()
    assert abs(mass - mass_old) <= 1e-12 * abs(mass_old)


def test_pairs_with_explicit_component_numbers(build_pairs):
    """ElementInfo::componentNumber exported by the adapter: face-adjacent zones stay separate components"""
    case = adapt_case(n=10, seed=8, n_insert=10, n_remove=10)
    pr0 = build_pairs(case)
    # numbers from the default grouping (shifted, as the reference counts from 1): the same result must come back
    comp = pr0["comp_old"] + 1
    pr1 = build_pairs(case, comp_old=comp)
    for key in ("new_off", "new_cell", "old_idx", "comp_new"):
        assert np.array_equal(pr0[key], pr1[key])
    comp_bad = comp.copy(); comp_bad[np.nonzero(case["vanished"])[0][0]] = 0
    with pytest.raises(RuntimeError):
        build_pairs(case, comp_old=comp_bad)


def test_pairs_no_change(build_pairs):
    case = adapt_case(n=6, seed=1, n_insert=0, n_remove=0)
    pr = build_pairs(case)
    assert pr["new_cell"].size == 0 and pr["old_idx"].size == 0 and pr["ncomp"] == 0
