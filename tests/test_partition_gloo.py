"""Host-side logic of the multi-GPU path on CPU: Hilbert-range shards with one-ring halos (mmesh_b200.partition)
and the halo shift exchange (mmesh_b200.multigpu.exchange_rows) over a world_size-2/3 gloo group."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shards_cover_mesh_and_halo_is_one_ring():
    from mmesh_b200 import partition, synth
    m = synth.config1(n=40)
    world = 3
    owned_cells = np.zeros(m.nc, int)
    owned_verts = np.zeros(m.nv, int)
    owned_edges = 0
    for r in range(world):
        sh = partition.make_shard(m, world, r)
        owned_cells[sh.lcells[sh.own_lo:sh.own_hi]] += 1
        owned_verts[sh.lverts[sh.owned_vertex]] += 1
        owned_edges += sh.edges.shape[0]
        # local numbering preserves the global (Hilbert) order
        assert np.all(np.diff(sh.lcells) > 0) and np.all(np.diff(sh.lverts) > 0)
        # local cell table refers to the same coordinates as the global one
        assert np.array_equal(sh.xy[sh.cells], m.xy[m.cells[sh.lcells]])
        # one-ring: every global cell sharing a vertex with an owned cell is local
        vm = np.zeros(m.nv, bool)
        vm[m.cells[sh.c0:sh.c1].ravel()] = True
        assert np.array_equal(np.nonzero(vm[m.cells].any(1))[0], sh.lcells)
        # owned edges see both apices locally
        assert np.array_equal(sh.xy[sh.edges[:, 2]], m.xy[m.edges[np.searchsorted(m.edge_cell, sh.c0):np.searchsorted(m.edge_cell, sh.c1), 2]])
        # recv list = exactly the non-owned local vertices; send/recv counts are consistent across ranks
        assert set(sh.recv_idx.tolist()) == set(np.nonzero(~sh.owned_vertex)[0].tolist())
        for p in range(world):
            if p != r:
                shp = partition.make_shard(m, world, p)
                assert sh.send_counts[p] == shp.recv_counts[r]
    assert np.all(owned_cells == 1) and np.all(owned_verts == 1) and owned_edges == m.edges.shape[0]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from mmesh_b200 import multigpu, partition, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = synth.config1(n=32)
    s = synth.shifts_uniform(m, 0.3 * m.h, 30)
    sh = partition.make_shard(m, world, rank)
    local = s[sh.lverts].copy()
    local[~sh.owned_vertex] = np.nan                       # a rank only knows the shifts of its own vertices
    send = torch.from_numpy(np.ascontiguousarray(local[sh.send_idx]))          # halo_pack (host mirror of k_gather2)
    recv = torch.empty((sh.recv_idx.shape[0], 2), dtype=torch.float64)
    multigpu.exchange_rows(dist, send, recv, sh.send_counts, sh.recv_counts, rank, world)
    local[sh.recv_idx] = recv.numpy()                                          # halo_unpack (k_scatter2)
    ok = bool(np.array_equal(local, s[sh.lverts]))
    # the two scalar reductions of a step
    mx = torch.tensor([np.float64(rank + 0.5).view(np.int64)], dtype=torch.int64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    vec = torch.tensor([float(sh.own_hi - sh.own_lo)], dtype=torch.float64)
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    ok = ok and mx.numpy().view(np.float64)[0] == world - 0.5 and vec.item() == m.nc
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
