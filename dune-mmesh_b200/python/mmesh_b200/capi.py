"""ctypes binding of include/mmesh_b200.h (the C-ABI drop-in boundary).

There is NO CPU fallback: if the CUDA library is missing or no B200 is visible every call raises.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # dune-mmesh_b200/
_SO = os.environ.get("MMB_LIB") or os.path.join(_PKG, "lib", "libmmesh_b200.so")   # MMB_LIB: instrumented debug builds

MMB_OK, MMB_ERR_INVALID_ARG, MMB_ERR_CUDA, MMB_ERR_NCCL, MMB_ERR_INVALID_CELL_3D, MMB_ERR_CAPACITY, MMB_ERR_STATE, MMB_MOVE_WITHHELD = range(8)
_STATUS = {1: "invalid argument", 2: "CUDA error", 3: "NCCL error", 4: "invalid 3-D cell", 5: "capacity", 6: "state"}

EXPORTS = [
    "mmb_create", "mmb_destroy", "mmb_last_error", "mmb_stream", "mmb_synchronize", "mmb_upload_mesh", "mmb_patch_mesh",
    "mmb_upload_coords", "mmb_download_coords", "mmb_upload_shifts", "mmb_scale_shifts", "mmb_set_params", "mmb_get_params",
    "mmb_indicator_init", "mmb_move_vertices", "mmb_trial_move_validate", "mmb_invalid_cells", "mmb_legality",
    "mmb_distance_update", "mmb_distance_set", "mmb_mark_elements", "mmb_mark_interface", "mmb_marked_cells",
    "mmb_upload_pairs", "mmb_project", "mmb_intersection_volumes", "mmb_adapt_step", "mmb_adapt_step_host", "mmb_trace_skeleton_p0",
    "mmb_step_host_begin", "mmb_step_host_marks", "mmb_step_host_finish", "mmb_upload_neighbours", "mmb_refinement_candidates", "mmb_upload_vertex_levels", "mmb_coarsening_candidates", "mmb_interface_candidates", "mmb_project_interface_p0", "mmb_project_interface_p1", "mmb_refinement_points", "mmb_upload_dofs", "mmb_download_fields", "mmb_set_owned_cells", "mmb_halo_setup", "mmb_halo_pack_shifts", "mmb_halo_unpack_shifts", "mmb_device_buffer",
    "mmb_component_numbers", "mmb_host_register", "mmb_host_unregister", "mmb_comm_unique_id", "mmb_comm_init", "mmb_comm_destroy", "mmb_halo_counts", "mmb_adapt_step_sharded", "mmb_step_phase_distance", "mmb_step_phase_rest", "mmb_step_phase", "mmb_timing_enable", "mmb_timing_reset", "mmb_timing_get", "mmb_launch_count", "mmb_predicate_batch",
    "mmb_restore_delaunay", "mmb_download_topology", "mmb_flip_log",
]


class MMeshError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("mmesh_b200: %s (%s)" % (msg, _STATUS.get(status, status)))
        self.status = status


class GridError(MMeshError):
    """Raised where the reference does DUNE_THROW(GridError, ...) (mmesh.hh:1104-1107)."""


class Params(C.Structure):
    _fields_ = [("minH", C.c_double), ("maxH", C.c_double), ("factor", C.c_double), ("distProportion", C.c_double),
                ("edgeRatio", C.c_double), ("radiusRatio", C.c_double), ("clip_eps", C.c_double),
                ("drop_area", C.c_double), ("clip_translate", C.c_int32), ("move_policy", C.c_int32)]


class StepResult(C.Structure):
    _fields_ = [("n_invalid", C.c_int64), ("n_orient_nonpos", C.c_int64), ("n_illegal", C.c_int64),
                ("n_refine", C.c_int64), ("n_coarsen", C.c_int64), ("n_exact_orient", C.c_int64),
                ("n_exact_incircle", C.c_int64), ("n_pieces", C.c_int64), ("max_dist", C.c_double),
                ("mass_new", C.c_double), ("covered_area", C.c_double), ("moved", C.c_int64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class FlipResult(C.Structure):
    _fields_ = [("n_flips", C.c_int64), ("n_illegal_left", C.c_int64), ("n_constrained_illegal", C.c_int64),
                ("rounds", C.c_int32), ("converged", C.c_int32)]


def build(force=False):
    """Compile the sm_100a library in-tree (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_PKG, "csrc", f) for f in ("mmesh_b200.cu", "kernels.cuh", "predicates.cuh", "clip.cuh", "geom.cuh", "project.cuh", "flip.cuh")]
    srcs.append(os.path.join(os.path.dirname(_PKG), "include", "mmesh_b200.h"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _PKG, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise MMeshError(MMB_ERR_CUDA, "CUDA library %s is not built (run __graft_entry__.build())" % _SO)
        L = C.CDLL(_SO)
        for name in EXPORTS:
            getattr(L, name).restype = C.c_int
        L.mmb_last_error.restype = C.c_char_p
        L.mmb_stream.restype = C.c_void_p
        L.mmb_launch_count.restype = C.c_int64
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class Context:
    """One mmb_context (one GPU).  Thin, explicit wrapper: numpy host arrays in, numpy host arrays out."""

    def __init__(self, dim=2, device=0, stream=None):
        self._h = C.c_void_p()
        self.dim = dim
        rc = lib().mmb_create(C.c_int(device), C.c_int(dim), C.c_void_p(stream), C.byref(self._h))
        if rc != MMB_OK:
            self._h = None
            raise MMeshError(rc, "mmb_create failed: no usable sm_100a device/kernel image (there is no CPU fallback)")
        self.nv = self.nc = self.ne = self.ns = 0
        self.npairs = 0

    def close(self):
        if self._h:
            lib().mmb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc == MMB_OK or rc == MMB_MOVE_WITHHELD:      # withheld move: not an error, StepResult.moved == 0 says so
            return
        msg = lib().mmb_last_error(self._h).decode()
        raise (GridError if rc == MMB_ERR_INVALID_CELL_3D else MMeshError)(rc, msg)

    # ---- snapshot ----
    def upload_mesh(self, x, cells, perm, edges=None, facets=None):
        x, cells = _f64(x), _i32(cells)
        perm = np.ascontiguousarray(perm, np.uint8)
        edges = _i32(edges) if edges is not None and len(edges) else None
        facets = _i32(facets) if facets is not None and len(facets) else None
        self.nv, self.nc = x.shape[0], cells.shape[0]
        self.ne = 0 if edges is None else edges.shape[0]
        self.ns = 0 if facets is None else facets.shape[0]
        self._ck(lib().mmb_upload_mesh(self._h, C.c_int64(self.nv), _p(x), C.c_int64(self.nc), _p(cells), _p(perm),
                                       C.c_int64(self.ne), _p(edges), C.c_int64(self.ns), _p(facets)))

    def patch_mesh(self, nv, vert_idx, vert_x, nc, cell_idx, cell_v, cell_perm, ne=0, edge_idx=None, edge_q=None, facets=None, ns=None):
        """incremental snapshot update (mmb_patch_mesh): only the listed vertices / cells / edges are transferred"""
        vi, vx = _i32(vert_idx), _f64(vert_x)
        ci, cv = _i32(cell_idx), _i32(cell_v)
        cp = np.ascontiguousarray(cell_perm, np.uint8)
        ei = _i32(edge_idx) if edge_idx is not None else None
        eq = _i32(edge_q) if edge_q is not None else None
        fc = _i32(facets) if facets is not None else None
        ns_new = self.ns if ns is None and fc is None else (fc.shape[0] if fc is not None else ns)
        self._ck(lib().mmb_patch_mesh(self._h, C.c_int64(nv), C.c_int64(vi.shape[0]), _p(vi) if vi.size else None, _p(vx) if vi.size else None,
                                      C.c_int64(nc), C.c_int64(ci.shape[0]), _p(ci) if ci.size else None, _p(cv) if ci.size else None, _p(cp) if ci.size else None,
                                      C.c_int64(ne), C.c_int64(0 if ei is None else ei.shape[0]), None if ei is None or not ei.size else _p(ei),
                                      None if eq is None or not eq.size else _p(eq), C.c_int64(ns_new), None if fc is None else _p(fc)))
        self.nv, self.nc, self.ne, self.ns = nv, nc, ne, ns_new

    def restore_delaunay(self, vertex_id=None, max_rounds=1000):
        """Delaunay restoration by edge flips on the resident snapshot (mmb_restore_delaunay).  Returns a dict with n_flips,
        n_illegal_left, n_constrained_illegal, rounds, converged."""
        r = FlipResult()
        vid = None if vertex_id is None else np.ascontiguousarray(vertex_id, np.int64)
        self._ck(lib().mmb_restore_delaunay(self._h, _p(vid), C.c_int32(max_rounds), C.byref(r)))
        return {k: int(getattr(r, k)) for k, _ in FlipResult._fields_}

    def flip_log(self):
        """flips of the last restore_delaunay: (edge ids 3 * cell + slot in round order, round offsets [rounds + 1], total flips)"""
        tot, nr = C.c_int64(), C.c_int32()
        self._ck(lib().mmb_flip_log(self._h, None, C.c_int64(0), C.byref(tot), None, C.c_int32(0), C.byref(nr)))
        ids = np.zeros(max(tot.value, 1), np.uint32); off = np.zeros(nr.value + 1, np.int64)
        self._ck(lib().mmb_flip_log(self._h, _p(ids), C.c_int64(tot.value), C.byref(tot), _p(off), C.c_int32(nr.value + 1), C.byref(nr)))
        return ids[:tot.value], off, int(tot.value)

    def download_topology(self, neighbours=True, edges=True):
        """resident cells [nc][dim+1], perm [nc], neighbours [nc][3] and edges [ne][4] (2-D)"""
        k = self.dim + 1
        cells = np.empty((self.nc, k), np.int32); perm = np.empty(self.nc, np.uint8)
        nb = np.empty((self.nc, 3), np.int32) if neighbours and self.dim == 2 else None
        e = np.empty((self.ne, 4), np.int32) if edges and self.dim == 2 and self.ne else None
        self._ck(lib().mmb_download_topology(self._h, _p(cells), _p(perm), _p(nb), _p(e)))
        return cells, perm, nb, e

    def upload_coords(self, x):
        x = _f64(x)
        self._ck(lib().mmb_upload_coords(self._h, _p(x)))

    def download_coords(self):
        x = np.empty((self.nv, self.dim), np.float64)
        self._ck(lib().mmb_download_coords(self._h, _p(x)))
        return x

    def upload_shifts(self, s):
        s = _f64(s)
        self._ck(lib().mmb_upload_shifts(self._h, _p(s)))
        self._ck(lib().mmb_synchronize(self._h))

    def scale_shifts(self, a):
        self._ck(lib().mmb_scale_shifts(self._h, C.c_double(a)))

    def set_params(self, prm):
        self._ck(lib().mmb_set_params(self._h, C.byref(prm)))

    def get_params(self):
        prm = Params()
        self._ck(lib().mmb_get_params(self._h, C.byref(prm)))
        return prm

    def synchronize(self):
        self._ck(lib().mmb_synchronize(self._h))

    @property
    def stream(self):
        return lib().mmb_stream(self._h)

    # ---- hot path ----
    def indicator_init(self):
        self._ck(lib().mmb_indicator_init(self._h))
        return self.get_params()

    def move_vertices(self, shifts=None):
        s = _f64(shifts)
        self._ck(lib().mmb_move_vertices(self._h, _p(s)))

    def trial_move_validate(self, shifts=None, download=True):
        s = _f64(shifts)
        valid = np.empty(self.nc, np.uint8) if download else None
        sign = np.empty(self.nc, np.int8) if download else None
        ninv = C.c_int64(-1)
        rc = lib().mmb_trial_move_validate(self._h, _p(s), _p(valid), _p(sign), C.byref(ninv))
        if rc == MMB_ERR_INVALID_CELL_3D:
            err = GridError(rc, lib().mmb_last_error(self._h).decode())
            err.valid_fp, err.orient_sign, err.n_invalid = valid, sign, ninv.value
            raise err
        self._ck(rc)
        return valid, sign, ninv.value

    def invalid_cells(self):
        n = C.c_int64()
        self._ck(lib().mmb_invalid_cells(self._h, None, C.c_int64(0), C.byref(n)))
        out = np.empty(n.value, np.int32)
        if n.value:
            self._ck(lib().mmb_invalid_cells(self._h, _p(out), C.c_int64(n.value), C.byref(n)))
        return out

    def legality(self, download=True):
        flags = np.empty(self.ne, np.uint8) if download else None
        n = C.c_int64(-1)
        self._ck(lib().mmb_legality(self._h, _p(flags), C.byref(n)))
        return flags, n.value

    def distance_update(self, download=True):
        d = np.empty(self.nv, np.float64) if download else None
        mx = C.c_double()
        self._ck(lib().mmb_distance_update(self._h, _p(d), C.byref(mx)))
        return d, mx.value

    def distance_set(self, d):
        d = _f64(d)
        self._ck(lib().mmb_distance_set(self._h, _p(d)))

    def mark_elements(self, run_distance=True, download=True):
        m = np.empty(self.nc, np.int8) if download else None
        le = np.empty(self.nc, np.uint8) if download else None
        counts = (C.c_int64 * 2)()
        self._ck(lib().mmb_mark_elements(self._h, C.c_int(1 if run_distance else 0), _p(m), _p(le), counts))
        return m, le, (counts[0], counts[1])

    def mark_interface(self):
        m = np.empty(self.ns, np.int8)
        self._ck(lib().mmb_mark_interface(self._h, _p(m)))
        return m

    def marked_cells(self, which):
        n = C.c_int64()
        self._ck(lib().mmb_marked_cells(self._h, C.c_int(which), None, C.c_int64(0), C.byref(n)))
        out = np.empty(n.value, np.int32)
        if n.value:
            self._ck(lib().mmb_marked_cells(self._h, C.c_int(which), _p(out), C.c_int64(n.value), C.byref(n)))
        return out

    def upload_pairs(self, new_off, new_cell, old_xy, old_idx):
        new_off = np.ascontiguousarray(new_off, np.int64)
        new_cell, old_idx, old_xy = _i32(new_cell), _i32(old_idx), _f64(old_xy)
        self.npairs = old_idx.shape[0]
        self._ck(lib().mmb_upload_pairs(self._h, C.c_int64(new_cell.shape[0]), _p(new_off), _p(new_cell),
                                        C.c_int64(self.npairs), _p(old_xy), _p(old_idx)))

    def project(self, ncomp, dofs_old=None, download=True):
        d = _f64(dofs_old)
        n_old = 0 if d is None else d.size // ncomp
        out = np.empty(self.nc * ncomp, np.float64) if download else None
        mass, cov, npc = C.c_double(), C.c_double(), C.c_int64()
        self._ck(lib().mmb_project(self._h, C.c_int(ncomp), C.c_int64(n_old), _p(d), _p(out), C.byref(mass),
                                   C.byref(cov), C.byref(npc)))
        if out is not None and ncomp > 1:
            out = out.reshape(self.nc, ncomp)
        return out, mass.value, cov.value, npc.value

    def intersection_volumes(self):
        v = np.empty(self.npairs, np.float64)
        self._ck(lib().mmb_intersection_volumes(self._h, _p(v)))
        return v

    def adapt_step(self, shifts=None, ncomp=0, want_result=True):
        s = _f64(shifts)
        res = StepResult() if want_result else None
        rc = lib().mmb_adapt_step(self._h, _p(s), C.c_int(ncomp), C.byref(res) if want_result else None)
        self._ck(rc)
        return res

    def adapt_step_host(self, shifts, ncomp, dofs_old):
        """The per-step drop-in call with host buffers for every field (mmb_adapt_step_host)."""
        s, d = _f64(shifts), _f64(dofs_old)
        out = dict(marks=np.empty(self.nc, np.int8), longest=np.empty(self.nc, np.uint8), valid_fp=np.empty(self.nc, np.uint8),
                   orient_sign=np.empty(self.nc, np.int8), edge_flags=np.empty(self.ne, np.uint8),
                   dofs_new=np.zeros(self.nc * max(ncomp, 1), np.float64))
        res = StepResult()
        self._ck(lib().mmb_adapt_step_host(self._h, _p(s), C.c_int(ncomp), C.c_int64(0 if d is None else d.size // max(ncomp, 1)), _p(d),
                                           _p(out["dofs_new"]) if ncomp else None, _p(out["marks"]), _p(out["longest"]),
                                           _p(out["valid_fp"]), _p(out["orient_sign"]), _p(out["edge_flags"]), C.byref(res)))
        if ncomp > 1:
            out["dofs_new"] = out["dofs_new"].reshape(self.nc, ncomp)
        out["result"] = res
        return out

    def project_interface_p0(self, new_len, new_off, old_len, old_idx, dofs_old):
        new_len, old_len, dofs_old = _f64(new_len), _f64(old_len), _f64(dofs_old)
        new_off = np.ascontiguousarray(new_off, np.int64)
        old_idx = _i32(old_idx)
        out = np.empty(new_len.shape[0])
        self._ck(lib().mmb_project_interface_p0(self._h, C.c_int64(new_len.shape[0]), _p(new_len), _p(new_off), C.c_int64(old_idx.shape[0]),
                                                _p(old_len), _p(old_idx), C.c_int64(dofs_old.shape[0]), _p(dofs_old), _p(out)))
        return out

    def component_numbers(self, zone_off, zone_cells, component_count=1):
        """adapt_()'s component numbering (mmesh.hh:1194-1218, 1919-1941): (zone_number, cell_number, component_count)."""
        zone_off = np.ascontiguousarray(zone_off, np.int64)
        zone_cells = _i32(zone_cells)
        nz = zone_off.shape[0] - 1
        zn = np.zeros(max(nz, 1), np.int32)
        cn = np.zeros(self.nc, np.int32)
        cc = C.c_int64()
        self._ck(lib().mmb_component_numbers(self._h, C.c_int64(nz), _p(zone_off), _p(zone_cells), C.c_int64(component_count), _p(zn), _p(cn),
                                             C.byref(cc)))
        return zn[:nz], cn, cc.value

    def project_interface_p1(self, new_xy, new_off, old_xy, old_idx, dofs_old):
        new_xy, old_xy, dofs_old = _f64(new_xy).reshape(-1, 4), _f64(old_xy).reshape(-1, 4), _f64(dofs_old).reshape(-1, 2)
        new_off = np.ascontiguousarray(new_off, np.int64)
        old_idx = _i32(old_idx)
        out = np.empty((new_xy.shape[0], 2))
        self._ck(lib().mmb_project_interface_p1(self._h, C.c_int64(new_xy.shape[0]), _p(new_xy), _p(new_off), C.c_int64(old_idx.shape[0]),
                                                _p(old_xy), _p(old_idx), C.c_int64(dofs_old.shape[0]), _p(dofs_old), _p(out)))
        return out

    def upload_neighbours(self, nb):
        nb = _i32(nb)
        self._ck(lib().mmb_upload_neighbours(self._h, _p(nb)))

    def refinement_candidates(self):
        n = C.c_int64()
        self._ck(lib().mmb_refinement_candidates(self._h, None, C.c_int64(0), C.byref(n)))
        out = np.empty(n.value, np.int32)
        if n.value:
            self._ck(lib().mmb_refinement_candidates(self._h, _p(out), C.c_int64(n.value), C.byref(n)))
        return out

    def upload_vertex_levels(self, levels):
        levels = _i32(levels)
        self._ck(lib().mmb_upload_vertex_levels(self._h, _p(levels)))

    def coarsening_candidates(self):
        """(cells, vertices) of remove_ after the bulk loop of adapt() (mmesh.hh:853-865)."""
        n = C.c_int64()
        self._ck(lib().mmb_coarsening_candidates(self._h, None, None, C.c_int64(0), C.byref(n)))
        cells = np.empty(n.value, np.int32)
        verts = np.empty(n.value, np.int32)
        if n.value:
            self._ck(lib().mmb_coarsening_candidates(self._h, _p(cells), _p(verts), C.c_int64(n.value), C.byref(n)))
        return cells, verts

    def interface_candidates(self):
        """(insert_segs, remove_segs, remove_vertices) of the interface loop of adapt() (mmesh.hh:869-908)."""
        ni, nr = C.c_int64(), C.c_int64()
        self._ck(lib().mmb_interface_candidates(self._h, None, C.c_int64(0), C.byref(ni), None, None, C.c_int64(0), C.byref(nr)))
        ins = np.empty(ni.value, np.int32)
        rs = np.empty(nr.value, np.int32)
        rv = np.empty(nr.value, np.int32)
        if ni.value or nr.value:
            self._ck(lib().mmb_interface_candidates(self._h, _p(ins) if ni.value else None, C.c_int64(ni.value), C.byref(ni),
                                                    _p(rs) if nr.value else None, _p(rv) if nr.value else None,
                                                    C.c_int64(nr.value), C.byref(nr)))
        return ins, rs, rv

    def refinement_points(self, cells):
        cells = _i32(cells)
        pts = np.empty((cells.shape[0], self.dim))
        ev = np.empty((cells.shape[0], 2), np.int32)
        self._ck(lib().mmb_refinement_points(self._h, C.c_int64(cells.shape[0]), _p(cells), _p(pts), _p(ev)))
        return pts, ev

    def trace_skeleton_p0(self, inside, outside, u_bulk, u_gamma):
        inside, outside, u_bulk, u_gamma = _i32(inside), _i32(outside), _f64(u_bulk), _f64(u_gamma)
        nf = inside.shape[0]
        a, b, c = np.empty(nf), np.empty(nf), np.empty(nf)
        self._ck(lib().mmb_trace_skeleton_p0(self._h, C.c_int64(nf), _p(inside), _p(outside), _p(u_bulk),
                                             C.c_int64(u_bulk.shape[0]), _p(u_gamma), _p(a), _p(b), _p(c)))
        return a, b, c

    def predicate_batch(self, which, pts):
        k = {0: 6, 1: 8, 2: 12}[which]
        pts = _f64(pts).reshape(-1, k)
        out = np.empty(pts.shape[0], np.int8)
        nex = C.c_int64()
        self._ck(lib().mmb_predicate_batch(self._h, C.c_int(which), C.c_int64(pts.shape[0]), _p(pts), _p(out), C.byref(nex)))
        return out, nex.value

    # ---- multi-GPU plumbing ----
    def set_owned_cells(self, lo, hi):
        self._ck(lib().mmb_set_owned_cells(self._h, C.c_int64(lo), C.c_int64(hi)))

    def halo_setup(self, send_idx, recv_idx):
        send_idx, recv_idx = _i32(send_idx), _i32(recv_idx)
        self._ck(lib().mmb_halo_setup(self._h, C.c_int64(send_idx.shape[0]), _p(send_idx), C.c_int64(recv_idx.shape[0]), _p(recv_idx)))

    def halo_pack_shifts(self):
        self._ck(lib().mmb_halo_pack_shifts(self._h))

    def halo_unpack_shifts(self):
        self._ck(lib().mmb_halo_unpack_shifts(self._h))

    def device_buffer(self, which):
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._ck(lib().mmb_device_buffer(self._h, C.c_int(which), C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    # ---- collectives inside the library (NCCL) ----
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        rc = lib().mmb_comm_unique_id(buf)
        if rc != MMB_OK:
            raise MMeshError(rc, "mmb_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def comm_init(self, world, rank, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(lib().mmb_comm_init(self._h, C.c_int(world), C.c_int(rank), buf))

    def halo_counts(self, send_counts, recv_counts):
        sc, rc = np.ascontiguousarray(send_counts, np.int64), np.ascontiguousarray(recv_counts, np.int64)
        self._ck(lib().mmb_halo_counts(self._h, _p(sc), _p(rc)))

    def adapt_step_sharded(self, shifts=None, ncomp=0, want_result=True):
        s = _f64(shifts)
        res = StepResult() if want_result else None
        self._ck(lib().mmb_adapt_step_sharded(self._h, _p(s), C.c_int(ncomp), C.byref(res) if want_result else None))
        return res

    def download_fields(self, ncomp=0):
        """resident per-cell / per-edge fields of the last step (synchronises)"""
        out = dict(marks=np.empty(self.nc, np.int8), longest=np.empty(self.nc, np.uint8), valid_fp=np.empty(self.nc, np.uint8),
                   orient_sign=np.empty(self.nc, np.int8), edge_flags=np.empty(self.ne, np.uint8),
                   dofs_new=np.empty(self.nc * ncomp, np.float64) if ncomp else None)
        self._ck(lib().mmb_download_fields(self._h, C.c_int(ncomp), _p(out["dofs_new"]), _p(out["marks"]), _p(out["longest"]), _p(out["valid_fp"]),
                                           _p(out["orient_sign"]), _p(out["edge_flags"]) if self.ne else None))
        self._ck(lib().mmb_synchronize(self._h))
        if ncomp > 1:
            out["dofs_new"] = out["dofs_new"].reshape(self.nc, ncomp)
        return out

    def step_phase_distance(self):
        self._ck(lib().mmb_step_phase_distance(self._h))

    def step_phase_rest(self, ncomp):
        self._ck(lib().mmb_step_phase_rest(self._h, C.c_int(ncomp)))

    def step_phase(self, which, ncomp):
        self._ck(lib().mmb_step_phase(self._h, C.c_int(which), C.c_int(ncomp)))

    # ---- instrumentation ----
    def timing_enable(self, on=True):
        self._ck(lib().mmb_timing_enable(self._h, C.c_int(1 if on else 0)))

    def timing_reset(self):
        self._ck(lib().mmb_timing_reset(self._h))

    def timing_get(self):
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        cnt = (C.c_int64 * cap)()
        n = lib().mmb_timing_get(self._h, C.c_int(cap), names, ms, cnt)
        return {names[i].decode(): (ms[i], cnt[i]) for i in range(min(n, cap))}

    @property
    def launch_count(self):
        return lib().mmb_launch_count(self._h)
