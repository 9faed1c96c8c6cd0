"""Hilbert-range sharding of a 2-D mesh snapshot with a one-ring halo (host logic; numpy only).

The reference's parallel scheme (dune/mmesh/misc/partitionhelper.hh:175-236) gives every rank the whole
triangulation, marks contiguous ranges of the iteration order as interior and adds one face-neighbour ghost
layer.  Here each rank stores only:
  * its contiguous range of the (Hilbert-sorted) cell order                       -> "owned" cells,
  * every cell sharing a vertex with an owned cell                                -> one-ring halo cells,
  * all vertices of those cells plus the (small, replicated) interface vertices   -> local vertices.
A vertex is owned by the rank owning its lowest-index incident cell.  Per step only two things cross NVLink:
the shifts of non-owned local vertices (halo exchange) and the scalar reductions.
"""
import numpy as np


def cell_bounds(nc, world):
    return np.array([(r * nc) // world for r in range(world + 1)], dtype=np.int64)


def vertex_owner_cell(cells, nv):
    """Lowest-index incident cell per vertex (interface-only vertices cannot occur: every vertex has a cell)."""
    nc = cells.shape[0]
    owner = np.full(nv, nc, dtype=np.int64)
    # assign in descending cell order: the last assignment (lowest cell) wins
    idx = cells[::-1].ravel()
    val = np.repeat(np.arange(nc - 1, -1, -1, dtype=np.int64), cells.shape[1])
    owner[idx] = val
    return owner


def local_sets(cells, segs, nv, c0, c1):
    """(local cell ids ascending, local vertex ids ascending) of the rank owning cells [c0, c1)."""
    vmask = np.zeros(nv, bool)
    vmask[cells[c0:c1].ravel()] = True
    touch = vmask[cells].any(axis=1)
    lcells = np.nonzero(touch)[0]
    lv = np.zeros(nv, bool)
    lv[cells[lcells].ravel()] = True
    if segs is not None and len(segs):
        lv[segs.ravel()] = True                                # interface is replicated
    return lcells, np.nonzero(lv)[0]


class Shard:
    """Everything one rank uploads, in local numbering (global order preserved => Hilbert locality preserved)."""
    pass


def make_shard(m, world, rank, new_off=None, new_cell=None, old_xy=None, old_idx=None, all_halo_lists=True):
    nv, nc = m.nv, m.nc
    bounds = cell_bounds(nc, world)
    c0, c1 = int(bounds[rank]), int(bounds[rank + 1])
    owner_cell = vertex_owner_cell(m.cells, nv)
    vowner = (np.searchsorted(bounds, owner_cell, side="right") - 1).astype(np.int32)
    lcells, lverts = local_sets(m.cells, m.segs, nv, c0, c1)
    g2l_v = np.full(nv, -1, np.int64)
    g2l_v[lverts] = np.arange(lverts.shape[0])
    g2l_c = np.full(nc, -1, np.int64)
    g2l_c[lcells] = np.arange(lcells.shape[0])
    sh = Shard()
    sh.rank, sh.world, sh.c0, sh.c1 = rank, world, c0, c1
    sh.lcells, sh.lverts = lcells, lverts
    sh.xy = np.ascontiguousarray(m.xy[lverts])
    sh.cells = np.ascontiguousarray(g2l_v[m.cells[lcells]].astype(np.int32))
    sh.perm = np.ascontiguousarray(m.perm[lcells])
    sh.own_lo = int(np.searchsorted(lcells, c0))
    sh.own_hi = int(np.searchsorted(lcells, c1))
    assert sh.own_hi - sh.own_lo == c1 - c0
    # owned edges = edges whose left (lower-index) cell is owned: a contiguous range of the edge array
    e0, e1 = int(np.searchsorted(m.edge_cell, c0)), int(np.searchsorted(m.edge_cell, c1))
    eg = m.edges[e0:e1]
    el = g2l_v[np.maximum(eg, 0)].astype(np.int32)
    el[eg < 0] = -1
    assert (el[eg >= 0] >= 0).all()
    sh.edges = np.ascontiguousarray(el)
    sh.segs = np.ascontiguousarray(g2l_v[m.segs].astype(np.int32)) if len(m.segs) else np.zeros((0, 2), np.int32)
    sh.vowner = vowner[lverts]
    sh.owned_vertex = sh.vowner == rank
    # halo lists.  recv: my non-owned local vertices grouped by owner rank (then ascending global id).
    nonown = np.nonzero(~sh.owned_vertex)[0]
    order = np.lexsort((lverts[nonown], sh.vowner[nonown]))
    sh.recv_idx = nonown[order].astype(np.int32)
    sh.recv_counts = np.bincount(sh.vowner[sh.recv_idx], minlength=world).astype(np.int64)
    # send: for every peer p, the vertices I own that p holds as non-owned, in p's receive order.
    send_idx, send_counts = [], np.zeros(world, np.int64)
    if all_halo_lists:
        for p in range(world):
            if p == rank:
                continue
            _, pverts = local_sets(m.cells, m.segs, nv, int(bounds[p]), int(bounds[p + 1]))
            mine = pverts[vowner[pverts] == rank]            # ascending global id == p's receive order within owner
            send_idx.append(g2l_v[mine])
            send_counts[p] = mine.shape[0]
            assert (g2l_v[mine] >= 0).all()
    sh.send_idx = (np.concatenate(send_idx) if send_idx else np.zeros(0, np.int64)).astype(np.int32)
    sh.send_counts = send_counts
    # projection pairs of the owned new cells: rows are selected by the cell they belong to, because the CSR of a real
    # adapt() lists only the isNew cells (rows != cells); new_cell is ascending (element order, mmesh.hh:1143)
    if new_off is not None:
        new_cell = np.asarray(new_cell)
        assert (np.diff(new_cell) > 0).all(), "new_cell must be ascending"
        r0, r1 = int(np.searchsorted(new_cell, c0)), int(np.searchsorted(new_cell, c1))
        p0, p1 = int(new_off[r0]), int(new_off[r1])
        sh.new_off = (new_off[r0:r1 + 1] - p0).astype(np.int64)
        sh.new_cell = g2l_c[new_cell[r0:r1]].astype(np.int32)
        sh.old_xy = np.ascontiguousarray(old_xy[p0:p1])
        oi = g2l_c[old_idx[p0:p1]]
        assert (oi >= 0).all() and (sh.new_cell >= 0).all()
        sh.old_idx = oi.astype(np.int32)
    return sh
