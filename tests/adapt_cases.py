"""Old / new snapshot pairs of one adapt step, generated host-side (no CGAL): edge-midpoint insertions
(refineEdge, mmesh.hh:768-790 -> insertInEdge_) and vertex removals with a fan re-triangulation of the star polygon
(removeVertex -> removeAndGiveNewElements, cgal/triangulationwrapper.hh:58-63).  Vertices do not move in between, so
old and new cells tile the same holes: the conservative set-up of movement.cc:89-94."""
import numpy as np

from mmesh_b200 import synth


def _area2(p, q, r):
    return (q[0] - p[0]) * (r[1] - p[1]) - (r[0] - p[0]) * (q[1] - p[1])


def adapt_case(n=12, seed=3, n_insert=6, n_remove=6, shuffle=True):
    rng = np.random.default_rng(seed)
    old = synth.lattice2d(n, n, 1.0, 1.0, jitter=0.2, seed=seed, hilbert=True)
    xy = old.xy
    cells = old.cells.astype(np.int64)
    vid = old.vid.astype(np.int64)
    nv = old.nv
    cells_of_v = [[] for _ in range(nv)]
    for c, tri in enumerate(cells):
        for v in tri:
            cells_of_v[v].append(c)
    touched = np.zeros(old.nc, bool)
    removed_cells, new_tris, new_pts, new_ids, removed_vertices = [], [], [], [], []
    next_id = int(vid.max()) + 1
    # edge-midpoint insertions on inner edges
    inner = np.nonzero(old.edges[:, 3] >= 0)[0]
    rng.shuffle(inner)
    cell_by_key = {tuple(sorted(t)): i for i, t in enumerate(cells.tolist())}
    done = 0
    for e in inner:
        if done == n_insert:
            break
        a, b, c, d = (int(t) for t in old.edges[e])
        L, R = cell_by_key[tuple(sorted((a, b, c)))], cell_by_key[tuple(sorted((a, b, d)))]
        if touched[L] or touched[R]:
            continue
        touched[[L, R]] = True
        m = nv + len(new_pts)
        new_pts.append(xy[a] + 0.5 * (xy[b] - xy[a]))
        new_ids.append(next_id); next_id += 1
        removed_cells += [L, R]
        new_tris += [(a, m, c), (m, b, c), (b, m, d), (m, a, d)]
        done += 1
    # vertex removals (interior vertices whose star is untouched); hole re-triangulated as a fan that keeps all areas positive
    interior = np.nonzero((old.vflags & 2) == 0)[0]
    rng.shuffle(interior)
    done = 0
    for v in interior:
        if done == n_remove:
            break
        star = cells_of_v[v]
        if any(touched[c] for c in star):
            continue
        nxt = {}
        for c in star:                                     # ccw link: in cell (v, p, q) ccw the link edge is p -> q
            t = cells[c].tolist()
            k = t.index(v)
            nxt[t[(k + 1) % 3]] = t[(k + 2) % 3]
        start = next(iter(nxt))
        ring = [start]
        while nxt[ring[-1]] != start:
            ring.append(nxt[ring[-1]])
        if len(ring) != len(star):
            continue
        ok_fan = None
        for s in range(len(ring)):
            r = ring[s:] + ring[:s]
            if all(_area2(xy[r[0]], xy[r[i]], xy[r[i + 1]]) > 1e-12 for i in range(1, len(r) - 1)):
                ok_fan = r
                break
        if ok_fan is None:
            continue
        touched[star] = True
        removed_cells += star
        removed_vertices.append(int(v))
        new_tris += [(ok_fan[0], ok_fan[i], ok_fan[i + 1]) for i in range(1, len(ok_fan) - 1)]
        done += 1
    # new snapshot: surviving vertices re-indexed, inserted ones appended; cells = kept + created, optionally shuffled
    keep_v = np.ones(nv, bool); keep_v[removed_vertices] = False
    all_xy = np.concatenate([xy, np.array(new_pts).reshape(-1, 2)])
    all_id = np.concatenate([vid, np.array(new_ids, np.int64)])
    keep_all = np.concatenate([keep_v, np.ones(len(new_pts), bool)])
    newidx = -np.ones(all_xy.shape[0], np.int64); newidx[keep_all] = np.arange(int(keep_all.sum()))
    kept_cells = np.ones(old.nc, bool); kept_cells[removed_cells] = False
    ncells = np.concatenate([cells[kept_cells], np.array(new_tris, np.int64).reshape(-1, 3)])
    created = np.concatenate([np.zeros(int(kept_cells.sum()), bool), np.ones(len(new_tris), bool)])
    if shuffle:
        p = rng.permutation(ncells.shape[0])
        ncells, created = ncells[p], created[p]
        rot = rng.integers(0, 3, ncells.shape[0])          # TDS start vertex is arbitrary
        ncells = np.take_along_axis(ncells, (np.arange(3)[None, :] + rot[:, None]) % 3, axis=1)
    new_xy, new_id = all_xy[keep_all], all_id[keep_all]
    new_cells = newidx[ncells].astype(np.int32)
    assert (new_cells >= 0).all()
    new_perm = synth.pack_perm(new_id[new_cells])
    return dict(old=old, old_id=vid, new_xy=np.ascontiguousarray(new_xy), new_id=new_id, new_cells=np.ascontiguousarray(new_cells),
                new_perm=new_perm, created=created, vanished=~kept_cells, n_insert=len(new_pts), n_remove=len(removed_vertices))
