"""The C++ partitioner (dune-mmesh_b200/host/partition.hh: Hilbert ranges + one-ring halo, what a C++ caller of the sharded
step uses) equals the Python partitioner of the bench driver (mmesh_b200.partition.make_shard) array for array."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "native", "_build", "libhost_partition.so")


@pytest.fixture(scope="module")
def cpp_shard():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", os.path.join(ROOT, "tests", "native", "host_partition.cc"), "-o", SO])
    L = C.CDLL(SO)

    def run(m, world, rank, pairs=None):
        P = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        x, cells, perm = np.ascontiguousarray(m.xy), np.ascontiguousarray(m.cells, np.int32), np.ascontiguousarray(m.perm, np.uint8)
        edges, ec, segs = np.ascontiguousarray(m.edges, np.int32), np.ascontiguousarray(m.edge_cell, np.int32), np.ascontiguousarray(m.segs, np.int32)
        sizes = (C.c_long * 9)()
        if pairs is None:
            no = ncl = oxy = oi = None; nrows = 0
        else:
            no, ncl = np.ascontiguousarray(pairs[0], np.int64), np.ascontiguousarray(pairs[1], np.int32)
            oxy, oi = np.ascontiguousarray(pairs[2], np.float64), np.ascontiguousarray(pairs[3], np.int32)
            nrows = ncl.shape[0]
        rc = L.hpart_make(C.c_long(m.nv), P(x), C.c_long(m.nc), P(cells), P(perm), C.c_long(edges.shape[0]), P(edges), P(ec), C.c_long(segs.shape[0]),
                          P(segs), C.c_int(world), C.c_int(rank), C.c_long(nrows), P(no), P(ncl), P(oxy), P(oi), sizes)
        assert rc == 0
        nlv, nlc, nle, lo, hi, nsend, nrecv, nr, npairs = (int(v) for v in sizes)
        out = dict(lverts=np.zeros(nlv, np.int32), lcells=np.zeros(nlc, np.int32), xy=np.zeros((nlv, 2)), cells=np.zeros((nlc, 3), np.int32),
                   perm=np.zeros(nlc, np.uint8), edges=np.zeros((nle, 4), np.int32), segs=np.zeros((segs.shape[0], 2), np.int32),
                   vowner=np.zeros(nlv, np.int32), send_idx=np.zeros(nsend, np.int32), recv_idx=np.zeros(nrecv, np.int32),
                   send_counts=np.zeros(world, np.int64), recv_counts=np.zeros(world, np.int64), new_off=np.zeros(nr + 1 if pairs else 0, np.int64),
                   new_cell=np.zeros(nr, np.int32), old_xy=np.zeros((npairs, 6)), old_idx=np.zeros(npairs, np.int32), own_lo=lo, own_hi=hi)
        L.hpart_get(*[P(out[k]) for k in ("lverts", "lcells", "xy", "cells", "perm", "edges", "segs", "vowner", "send_idx", "recv_idx", "send_counts",
                                           "recv_counts", "new_off", "new_cell", "old_xy", "old_idx")])
        return out

    return run


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_cpp_partitioner_equals_python(cpp_shard, world):
    from mmesh_b200 import partition, synth
    m = synth.config3(nx=48, ny=24)
    s = synth.shifts_uniform(m, 0.3 * m.h, 30)
    pairs = synth.pairs_self_and_neighbours(m, m.xy - s)
    for rank in range(world):
        ref = partition.make_shard(m, world, rank, *pairs)
        got = cpp_shard(m, world, rank, pairs)
        assert np.array_equal(got["lverts"], ref.lverts) and np.array_equal(got["lcells"], ref.lcells)
        assert np.array_equal(got["xy"], ref.xy) and np.array_equal(got["cells"], ref.cells) and np.array_equal(got["perm"], ref.perm)
        assert np.array_equal(got["edges"], ref.edges) and np.array_equal(got["segs"], ref.segs)
        assert (got["own_lo"], got["own_hi"]) == (ref.own_lo, ref.own_hi)
        assert np.array_equal(got["vowner"], ref.vowner)
        assert np.array_equal(got["send_idx"], ref.send_idx) and np.array_equal(got["recv_idx"], ref.recv_idx)
        assert np.array_equal(got["send_counts"], ref.send_counts) and np.array_equal(got["recv_counts"], ref.recv_counts)
        assert np.array_equal(got["new_off"], ref.new_off) and np.array_equal(got["new_cell"], ref.new_cell)
        assert np.array_equal(got["old_xy"], ref.old_xy) and np.array_equal(got["old_idx"], ref.old_idx)


def test_partition_rows_of_a_sparse_pair_list(cpp_shard):
    """a real adapt() lists only the isNew cells: rows are selected by cell id, not sliced by it (both partitioners)"""
    from mmesh_b200 import partition, synth
    m = synth.config3(nx=32, ny=16)
    s = synth.shifts_uniform(m, 0.3 * m.h, 30)
    new_off, new_cell, old_xy, old_idx = synth.pairs_self_and_neighbours(m, m.xy - s)
    keep = np.nonzero(np.arange(m.nc) % 5 == 2)[0]                     # every fifth cell is "new"
    cnt = np.diff(new_off)[keep]
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    rows = np.concatenate([np.arange(new_off[k], new_off[k + 1]) for k in keep])
    sparse = (off, new_cell[keep], old_xy.reshape(-1, 6)[rows], old_idx[rows])
    world = 3
    seen = 0
    for rank in range(world):
        ref = partition.make_shard(m, world, rank, *sparse)
        got = cpp_shard(m, world, rank, sparse)
        assert np.array_equal(got["new_off"], ref.new_off) and np.array_equal(got["new_cell"], ref.new_cell)
        assert np.array_equal(got["old_idx"], ref.old_idx) and np.array_equal(got["old_xy"], ref.old_xy)
        assert np.array_equal(ref.lcells[ref.new_cell], keep[(keep >= ref.c0) & (keep < ref.c1)])
        seen += ref.new_cell.size
    assert seen == keep.size
