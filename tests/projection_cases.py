"""Projection set-ups shared by the CPU and GPU tests."""
import numpy as np

from mmesh_b200 import synth


def diagonal_flip_case(n=24, seed=7, Lx=1.0, Ly=1.0, jitter=0.2):
    """Old mesh: jittered lattice.  New mesh: same vertices, every quad re-triangulated with the OTHER diagonal.
    Each new triangle's connected component is its quad (two old children, connectedcomponent.hh:104-127), so
    old∩new tiles every new cell exactly: the conservative set-up of movement.cc:89-94."""
    old = synth.lattice2d(n, n, Lx, Ly, jitter=jitter, seed=seed, hilbert=True)
    nv = old.nv
    # recover the quads from the old mesh: an inner edge whose two triangles form a lattice quad diagonal
    lat = old.lat
    e = old.edges
    inner = e[:, 3] >= 0
    a, b, c, d = (e[inner, k] for k in range(4))
    dl = np.abs(lat[a] - lat[b])
    isdiag = (dl[:, 0] == 1) & (dl[:, 1] == 1)
    a, b, c, d = a[isdiag], b[isdiag], c[isdiag], d[isdiag]
    assert a.shape[0] == n * n
    # old children: (a,b,c) ccw and (b,a,d) ccw; new triangles: (c,a,d) and (d,b,c) (ccw, other diagonal c-d)
    key_old = {}
    oc = old.cells
    srt = np.sort(oc, axis=1)
    lookup = {tuple(r): i for i, r in enumerate(srt.tolist())}
    o1 = np.array([lookup[tuple(sorted(t))] for t in np.stack([a, b, c], 1).tolist()], np.int32)
    o2 = np.array([lookup[tuple(sorted(t))] for t in np.stack([a, b, d], 1).tolist()], np.int32)
    new_cells = np.empty((2 * n * n, 3), np.int32)
    new_cells[0::2] = np.stack([c, a, d], 1)
    new_cells[1::2] = np.stack([d, b, c], 1)
    perm = synth.pack_perm(old.vid[new_cells])
    nnew = new_cells.shape[0]
    new_off = np.arange(0, 2 * nnew + 1, 2, dtype=np.int64)
    old_idx = np.empty(2 * nnew, np.int32)
    old_idx[0::4], old_idx[1::4], old_idx[2::4], old_idx[3::4] = o1, o2, o1, o2
    old_xy = synth.cached_corners(old, old.xy, old_idx)
    xy = old.xy
    pc = xy[new_cells]
    new_vol = np.abs(synth.cross2(pc[:, 1] - pc[:, 0], pc[:, 2] - pc[:, 0])) / 2
    po = xy[oc]
    old_vol = np.abs(synth.cross2(po[:, 1] - po[:, 0], po[:, 2] - po[:, 0])) / 2
    k = np.stack([np.take_along_axis(new_cells, ((perm >> (2 * i)) & 3)[:, None].astype(np.int64), 1)[:, 0] for i in range(3)], 1)
    first_pair = np.full(old.nc, -1, np.int64)
    first_pair[old_idx[::-1]] = np.arange(old_idx.shape[0])[::-1]
    return dict(xy=xy, cells=new_cells, perm=perm, new_off=new_off, new_cell=np.arange(nnew, dtype=np.int32),
                old_xy=old_xy, old_idx=old_idx, n_old=old.nc, new6=pc.reshape(nnew, 6), new_vol=new_vol,
                old_vol=old_vol, old_centroid=po.mean(axis=1), new_corners_idorder=xy[k], first_pair_of_old=first_pair,
                old_mesh=old)
