"""ctypes binding of the CPU oracle (oracle/mmesh_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmmesh_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "mmesh_oracle.c")
    hdr = os.path.join(_HERE, "mmesh_oracle.h")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Params(C.Structure):
    _fields_ = [("minH", C.c_double), ("maxH", C.c_double), ("factor", C.c_double), ("distProportion", C.c_double),
                ("edgeRatio", C.c_double), ("radiusRatio", C.c_double), ("clip_eps", C.c_double),
                ("drop_area", C.c_double), ("clip_translate", C.c_int32), ("pad_", C.c_int32)]

    @staticmethod
    def default(minH=0.0, maxH=0.0, factor=1.0, distProportion=1.0):
        return Params(minH, maxH, factor, distProportion, 4.0, 30.0, 1e-14, 1e-8, 0, 0)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        for name in ("orc_orient2d", "orc_orient3d", "orc_incircle", "orc_orient2d_filter", "orc_orient3d_filter",
                     "orc_incircle_filter", "orc_sutherland_hodgman", "orc_cutset", "orc_num_threads"):
            getattr(_lib, name).restype = C.c_int
        for name in ("orc_signed_area", "orc_signed_volume", "orc_point_segment_distance", "orc_distance_max",
                     "orc_polygon_area", "orc_intersection_volume"):
            getattr(_lib, name).restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def set_threads(n):
    lib().orc_set_threads(C.c_int(int(n)))


def orient2d(p, q, r):
    p, q, r = _f64(p), _f64(q), _f64(r)
    return lib().orc_orient2d(_p(p), _p(q), _p(r))


def orient3d(p, q, r, s):
    p, q, r, s = _f64(p), _f64(q), _f64(r), _f64(s)
    return lib().orc_orient3d(_p(p), _p(q), _p(r), _p(s))


def incircle(p, q, r, t):
    p, q, r, t = _f64(p), _f64(q), _f64(r), _f64(t)
    return lib().orc_incircle(_p(p), _p(q), _p(r), _p(t))


def orient2d_filter(p, q, r):
    p, q, r = _f64(p), _f64(q), _f64(r)
    return lib().orc_orient2d_filter(_p(p), _p(q), _p(r))


def orient3d_filter(p, q, r, s):
    p, q, r, s = _f64(p), _f64(q), _f64(r), _f64(s)
    return lib().orc_orient3d_filter(_p(p), _p(q), _p(r), _p(s))


def incircle_filter(p, q, r, t):
    p, q, r, t = _f64(p), _f64(q), _f64(r), _f64(t)
    return lib().orc_incircle_filter(_p(p), _p(q), _p(r), _p(t))


def _batch(fn, pts, k):
    pts = _f64(pts).reshape(-1, k)
    out = np.empty(pts.shape[0], dtype=np.int8)
    fn(_p(pts), C.c_int64(pts.shape[0]), _p(out))
    return out


def orient2d_batch(pts):
    return _batch(lib().orc_orient2d_batch, pts, 6)


def orient3d_batch(pts):
    return _batch(lib().orc_orient3d_batch, pts, 12)


def incircle_batch(pts):
    return _batch(lib().orc_incircle_batch, pts, 8)


def move(x, shift):
    x = _f64(x).copy()
    s = _f64(shift)
    lib().orc_move(_p(x), _p(s), C.c_int64(x.size))
    return x


def trial_validate(x, shift, cells, dim=2):
    x = _f64(x)
    s = None if shift is None else _f64(shift)
    cells = _i32(cells)
    nc = cells.shape[0]
    valid = np.empty(nc, np.uint8)
    sign = np.empty(nc, np.int8)
    ninv = C.c_int64(0)
    fn = lib().orc_trial_validate_2d if dim == 2 else lib().orc_trial_validate_3d
    fn(_p(x), _p(s), C.c_int64(x.shape[0]), _p(cells), C.c_int64(nc), _p(valid), _p(sign), C.byref(ninv))
    return valid, sign, ninv.value


def legality(xy, shift, edges):
    xy = _f64(xy)
    s = None if shift is None else _f64(shift)
    edges = _i32(edges)
    ne = edges.shape[0]
    flag = np.empty(ne, np.uint8)
    nill = C.c_int64(0)
    lib().orc_legality(_p(xy), _p(s), _p(edges), C.c_int64(ne), _p(flag), C.byref(nill))
    return flag, nill.value


def point_segment_distance(vp, v0, v1):
    vp, v0, v1 = _f64(vp), _f64(v0), _f64(v1)
    return lib().orc_point_segment_distance(_p(vp), _p(v0), _p(v1))


def distance_2d(xy, segs):
    xy, segs = _f64(xy), _i32(segs)
    d = np.empty(xy.shape[0], np.float64)
    lib().orc_distance_2d(_p(xy), C.c_int64(xy.shape[0]), _p(segs), C.c_int64(segs.shape[0]), _p(d))
    return d


def distance_2d_subset(xy, vidx, segs):
    xy, segs, vidx = _f64(xy), _i32(segs), _i32(vidx)
    d = np.empty(vidx.shape[0], np.float64)
    lib().orc_distance_2d_subset(_p(xy), _p(vidx), C.c_int64(vidx.shape[0]), _p(segs), C.c_int64(segs.shape[0]), _p(d))
    return d


def distance_2d_culled(xy, segs, vidx=None):
    """Labelled CULLED variant (grid buckets + ring search): the brute-force values, bit for bit, at O(Nv) cost."""
    xy, segs = _f64(xy), _i32(segs)
    vidx = _i32(np.arange(xy.shape[0]) if vidx is None else vidx)
    out = np.empty(vidx.shape[0], np.float64)
    lib().orc_distance_2d_culled_subset(_p(xy), C.c_int64(xy.shape[0]), _p(vidx), C.c_int64(vidx.shape[0]), _p(segs), C.c_int64(segs.shape[0]), _p(out))
    return out


def distance_3d(xyz, tris):
    xyz, tris = _f64(xyz), _i32(tris)
    d = np.empty(xyz.shape[0], np.float64)
    lib().orc_distance_3d(_p(xyz), C.c_int64(xyz.shape[0]), _p(tris), C.c_int64(tris.shape[0]), _p(d))
    return d


def distance_3d_subset(xyz, vidx, tris):
    """brute-force 3-D estimate (distance.hh:157-173) for the listed vertices only"""
    xyz, tris = _f64(xyz), _i32(tris)
    vidx = np.asarray(vidx, np.int64)
    sub = np.ascontiguousarray(np.concatenate([xyz[vidx], xyz]))      # the queried points first, facet indices shifted behind them
    d = np.empty(vidx.shape[0], np.float64)
    lib().orc_distance_3d(_p(sub), C.c_int64(vidx.shape[0]), _p(_i32(tris + vidx.shape[0])), C.c_int64(tris.shape[0]), _p(d))
    return d


def distance_max(dist):
    dist = _f64(dist)
    return lib().orc_distance_max(_p(dist), C.c_int64(dist.size))


def indicator_init_2d(xy, edges, segs, prm):
    xy, edges, segs = _f64(xy), _i32(edges), _i32(segs)
    lib().orc_indicator_init_2d(_p(xy), _p(edges), C.c_int64(edges.shape[0]), _p(segs), C.c_int64(segs.shape[0]),
                                C.byref(prm))
    return prm


def mark(x, dist, cells, perm, prm, maxDist, dim=2):
    x, dist, cells = _f64(x), _f64(dist), _i32(cells)
    perm = np.ascontiguousarray(perm, np.uint8)
    nc = cells.shape[0]
    m = np.empty(nc, np.int8)
    le = np.empty(nc, np.uint8)
    counts = (C.c_int64 * 2)()
    fn = lib().orc_mark_2d if dim == 2 else lib().orc_mark_3d
    fn(_p(x), _p(dist), _p(cells), _p(perm), C.c_int64(nc), C.byref(prm), C.c_double(maxDist), _p(m), _p(le), counts)
    return m, le, (counts[0], counts[1])


def mark_interface_2d(xy, segs, prm, maxDist):
    xy, segs = _f64(xy), _i32(segs)
    m = np.empty(segs.shape[0], np.int8)
    lib().orc_mark_interface_2d(_p(xy), _p(segs), C.c_int64(segs.shape[0]), C.byref(prm), C.c_double(maxDist), _p(m))
    return m


def coarsening_order(level, length, constrained):
    """libstdc++ insertion-sort order of the coarsening comparator (longestedgerefinement.hh:92-112)."""
    level = np.ascontiguousarray(level, np.int64)
    length = _f64(length)
    constrained = np.ascontiguousarray(constrained, np.uint8)
    n = level.shape[0]
    order = (C.c_int * n)()
    lib().orc_coarsening_order(C.c_int(n), _p(level), _p(length), _p(constrained), order)
    return list(order)


def coarsening_choice_2d(corners, level, constrained):
    corners = _f64(corners)
    level = np.ascontiguousarray(level, np.int64)
    constrained = np.ascontiguousarray(constrained, np.uint8)
    lib().orc_coarsening_choice_2d.restype = C.c_int
    return int(lib().orc_coarsening_choice_2d(_p(corners), _p(level), _p(constrained)))


def adapt_candidates_2d(xy, cells, perm, cell_nb, mark, longest_edge, segs, imark=None, vertex_level=None):
    """Both candidate loops of MMesh::adapt() (mmesh.hh:826-908).  Returns dict of int32 arrays."""
    xy, cells, cell_nb, segs = _f64(xy), _i32(cells), _i32(cell_nb), _i32(segs).reshape(-1, 2)
    perm = np.ascontiguousarray(perm, np.uint8)
    mark = np.ascontiguousarray(mark, np.int8)
    longest_edge = np.ascontiguousarray(longest_edge, np.uint8)
    nv, nc, ns = xy.shape[0], cells.shape[0], segs.shape[0]
    imark_a = None if imark is None else np.ascontiguousarray(imark, np.int8)
    lev = None if vertex_level is None else _i32(vertex_level)
    ins = np.empty(nc, np.int32); remc = np.empty(nc, np.int32); remv = np.empty(nc, np.int32)
    inss = np.empty(max(ns, 1), np.int32); rems = np.empty(max(ns, 1), np.int32); remsv = np.empty(max(ns, 1), np.int32)
    n = [C.c_int64() for _ in range(4)]
    lib().orc_adapt_candidates_2d(_p(xy), C.c_int64(nv), _p(cells), _p(perm), C.c_int64(nc), _p(cell_nb), _p(mark),
                                  _p(longest_edge), _p(segs), C.c_int64(ns),
                                  None if imark_a is None else _p(imark_a), None if lev is None else _p(lev),
                                  _p(ins), C.byref(n[0]), _p(remc), _p(remv), C.byref(n[1]),
                                  _p(inss), C.byref(n[2]), _p(rems), _p(remsv), C.byref(n[3]))
    return dict(insert_cells=ins[:n[0].value].copy(), remove_cells=remc[:n[1].value].copy(),
                remove_vertices=remv[:n[1].value].copy(), insert_segs=inss[:n[2].value].copy(),
                remove_segs=rems[:n[3].value].copy(), remove_seg_vertices=remsv[:n[3].value].copy())


def refine_candidates_3d(cells, perm, mark, longest_edge, tris):
    cells, tris = _i32(cells), _i32(tris).reshape(-1, 3)
    perm = np.ascontiguousarray(perm, np.uint8)
    mark = np.ascontiguousarray(mark, np.int8)
    longest_edge = np.ascontiguousarray(longest_edge, np.uint8)
    nc = cells.shape[0]
    ins = np.empty(nc, np.int32)
    ev = np.empty((nc, 2), np.int32)
    n = C.c_int64()
    lib().orc_refine_candidates_3d(_p(cells), _p(perm), C.c_int64(nc), _p(mark), _p(longest_edge), _p(tris), C.c_int64(tris.shape[0]),
                                   _p(ins), _p(ev), C.byref(n))
    return ins[:n.value].copy(), ev[:n.value].copy()


def cell_geometry_2d(xy, cell, perm):
    xy, cell = _f64(xy), _i32(cell)
    ctr = np.empty(2)
    vol = C.c_double()
    ec = np.empty(6)
    cc = np.empty(2)
    lib().orc_cell_geometry_2d(_p(xy), _p(cell), C.c_uint8(int(perm)), _p(ctr), C.byref(vol), _p(ec), _p(cc))
    return ctr, vol.value, ec.reshape(3, 2), cc


def cell_geometry_3d(xyz, cell, perm):
    xyz, cell = _f64(xyz), _i32(cell)
    ctr = np.empty(3)
    vol = C.c_double()
    cc = np.empty(3)
    lib().orc_cell_geometry_3d(_p(xyz), _p(cell), C.c_uint8(int(perm)), _p(ctr), C.byref(vol), _p(cc))
    return ctr, vol.value, cc


def sutherland_hodgman(subj, clip, eps=1e-14):
    subj, clip = _f64(subj), _f64(clip)
    out = np.empty(24)
    n = lib().orc_sutherland_hodgman(_p(subj), _p(clip), C.c_double(eps), _p(out))
    return out[:2 * n].reshape(n, 2).copy()


def sutherland_hodgman_batch(subj, clip, eps=1e-14):
    subj, clip = _f64(subj).reshape(-1, 6), _f64(clip).reshape(-1, 6)
    n = subj.shape[0]
    out = np.zeros((n, 18))
    cnt = np.zeros(n, np.int32)
    lib().orc_sutherland_hodgman_batch(_p(subj), _p(clip), C.c_int64(n), C.c_double(eps), _p(out), _p(cnt))
    return out, cnt


def clip_edge(poly, q1, q2, eps=1e-14):
    """One clip edge of polygoncutting.hh:31-93 on a general polygon (<= 6 points)."""
    poly, q1, q2 = _f64(poly), _f64(q1), _f64(q2)
    out = np.zeros(24)
    lib().orc_clip_edge.restype = C.c_int
    n = lib().orc_clip_edge(_p(poly), C.c_int(poly.shape[0]), _p(q1), _p(q2), C.c_double(eps), _p(out))
    return out[:2 * max(n, 0)].reshape(-1, 2).copy(), n


def polygon_area(pts):
    pts = _f64(pts)
    return lib().orc_polygon_area(_p(pts), C.c_int(pts.shape[0]))


def intersection_volume(old, new):
    old, new = _f64(old), _f64(new)
    return lib().orc_intersection_volume(_p(old), _p(new))


def cutset(old, new, prm):
    old, new = _f64(old), _f64(new)
    out = np.empty(72)
    n = lib().orc_cutset(_p(old), _p(new), C.byref(prm), _p(out))
    return out[:6 * n].reshape(n, 3, 2).copy()


def project(xy, cells, perm, new_off, new_cell, old_xy, old_idx, dofs_old, n_dofs_new, ncomp, prm):
    xy, cells = _f64(xy), _i32(cells)
    perm = np.ascontiguousarray(perm, np.uint8)
    new_off = np.ascontiguousarray(new_off, np.int64)
    new_cell, old_idx = _i32(new_cell), _i32(old_idx)
    old_xy, dofs_old = _f64(old_xy), _f64(dofs_old)
    dofs_new = np.zeros(n_dofs_new * ncomp, np.float64)
    mass, cov, npc = C.c_double(), C.c_double(), C.c_int64()
    lib().orc_project(_p(xy), _p(cells), _p(perm), _p(new_off), _p(new_cell), C.c_int64(new_cell.shape[0]),
                      _p(old_xy), _p(old_idx), _p(dofs_old), _p(dofs_new), C.c_int(ncomp), C.byref(prm),
                      C.byref(mass), C.byref(cov), C.byref(npc))
    return (dofs_new.reshape(n_dofs_new, ncomp) if ncomp > 1 else dofs_new), mass.value, cov.value, npc.value


def project_interface_p0(new_len, new_off, old_len, old_idx, dofs_old):
    new_len, old_len, dofs_old = _f64(new_len), _f64(old_len), _f64(dofs_old)
    new_off = np.ascontiguousarray(new_off, np.int64)
    old_idx = _i32(old_idx)
    out = np.empty(new_len.shape[0])
    lib().orc_project_interface_p0(_p(new_len), _p(new_off), C.c_int64(new_len.shape[0]), _p(old_len), _p(old_idx),
                                   _p(dofs_old), _p(out))
    return out


def component_numbers(zone_off, zone_cells, nc, component_count=1, cell_number=None):
    """adapt_()'s component numbering (mmesh.hh:1194-1218, 1919-1941).  Returns (zone_number, cell_number, component_count)."""
    zone_off = np.ascontiguousarray(zone_off, np.int64)
    zone_cells = _i32(zone_cells)
    nz = zone_off.shape[0] - 1
    zn = np.zeros(max(nz, 1), np.int32)
    cn = np.zeros(nc, np.int32) if cell_number is None else _i32(cell_number).copy()
    lib().orc_component_numbers.restype = C.c_int64
    cc = lib().orc_component_numbers(C.c_int64(nz), _p(zone_off), _p(zone_cells), C.c_int64(component_count), _p(zn), _p(cn))
    return zn[:nz], cn, int(cc)


def ensure_interface_movement_2d(x, cells, perm, vflags, facets, iverts, ishifts, level=None, removed=()):
    """mmesh.hh:920-1088 restated in C (oracle/mmesh_oracle.c).  Returns ("ok", change, removed, inserted) with inserted rows
    (edge a, edge b, x, y, v0, insertionLevel, crossing facet, component-is-interface-edge), or ("error", code)."""
    x, cells, facets, iverts, ishifts = _f64(x), _i32(cells), _i32(facets).reshape(-1, 2), _i32(iverts), _f64(ishifts).reshape(-1, 2)
    perm = np.ascontiguousarray(perm, np.uint8); vflags = np.ascontiguousarray(vflags, np.uint8)
    nv, nc, ns, niv = x.shape[0], cells.shape[0], facets.shape[0], iverts.shape[0]
    mask = np.zeros(nv, np.uint8); mask[list(removed)] = 1
    rem = np.zeros(nv + 1, np.int32); nrem = C.c_int64(); nins = C.c_int64()
    ie = np.zeros((nc + 1, 2), np.int32); ip = np.zeros((nc + 1, 2)); iv0 = np.zeros(nc + 1, np.int32); il = np.zeros(nc + 1, np.int64)
    ic = np.zeros(nc + 1, np.int32); icomp = np.zeros(nc + 1, np.uint8)
    lev = None if level is None else _i32(level)
    rc = lib().orc_ensure_interface_movement_2d(_p(x), C.c_int64(nv), _p(cells), _p(perm), C.c_int64(nc), _p(vflags), _p(facets), C.c_int64(ns),
                                                _p(iverts), C.c_int64(niv), _p(ishifts), None if lev is None else _p(lev), _p(mask), _p(rem),
                                                C.byref(nrem), _p(ie), _p(ip), _p(iv0), _p(il), _p(ic), _p(icomp), C.byref(nins))
    if rc < 0:
        return ("error", rc)
    n = nins.value
    ins = [(int(ie[i, 0]), int(ie[i, 1]), float(ip[i, 0]), float(ip[i, 1]), int(iv0[i]), int(il[i]), int(ic[i]), int(icomp[i])) for i in range(n)]
    return ("ok", bool(rc), rem[:nrem.value].tolist(), ins)


def project_interface_p1(new_xy, new_off, old_xy, old_idx, dofs_old):
    """interface/grid.hh:388-415 with DG-P1 data (2 nodal DOFs per segment)."""
    new_xy, old_xy, dofs_old = _f64(new_xy).reshape(-1, 4), _f64(old_xy).reshape(-1, 4), _f64(dofs_old).reshape(-1, 2)
    new_off = np.ascontiguousarray(new_off, np.int64); old_idx = _i32(old_idx)
    out = np.zeros((new_xy.shape[0], 2))
    lib().orc_project_interface_p1(_p(new_xy), _p(new_off), C.c_int64(new_xy.shape[0]), _p(old_xy), _p(old_idx), _p(dofs_old), _p(out))
    return out


def indicator_init_3d(xyz, cells, tris, prm):
    xyz, cells, tris = _f64(xyz), _i32(cells), _i32(tris).reshape(-1, 3)
    lib().orc_indicator_init_3d(_p(xyz), _p(cells), C.c_int64(cells.shape[0]), _p(tris), C.c_int64(tris.shape[0]), C.byref(prm))
    return prm


def mark_interface_3d(xyz, tris, prm, maxDist):
    xyz, tris = _f64(xyz), _i32(tris).reshape(-1, 3)
    m = np.empty(tris.shape[0], np.int8)
    lib().orc_mark_interface_3d(_p(xyz), _p(tris), C.c_int64(tris.shape[0]), C.byref(prm), C.c_double(maxDist), _p(m))
    return m


def restore_delaunay_2d(xy, cells, nb, segs=None, max_flips=-1):
    """Sequential Lawson flips (CGAL/Delaunay_triangulation_2.h:936-1022, Triangulation_data_structure_2.h:885-915) with the
    interface edges as constraints.  Returns (cells, nb, n_flips)."""
    xy, cells, nb = _f64(xy), _i32(cells).copy(), _i32(nb).copy()
    segs = np.zeros((0, 2), np.int32) if segs is None else _i32(segs).reshape(-1, 2)
    lib().orc_restore_delaunay_2d.restype = C.c_int64
    n = lib().orc_restore_delaunay_2d(_p(xy), _p(cells), _p(nb), C.c_int64(cells.shape[0]), _p(segs), C.c_int64(segs.shape[0]), C.c_int64(max_flips))
    if n < 0:
        raise RuntimeError("orc_restore_delaunay_2d: flip limit exceeded or inconsistent neighbour table")
    return cells, nb, int(n)
