"""Python mirror of the adaptation members the reference exports through pybind11 (dune/python/mmesh/grid.hh:82-137,
python/dune/mmesh/_move.py): same names, argument meaning and error behaviour, on top of the C-ABI.

Only the numeric part of the path lives here; topology edits (adapt) need CGAL and are the host's business -- after an
edit the host hands over a new snapshot (`MMesh(snapshot)` / `setSnapshot`).
"""
import numpy as np

from . import capi
from .capi import GridError  # noqa: F401  (re-exported: DUNE_THROW(GridError, ...))


class Distance:
    """Dune::Distance<Grid> (dune/mmesh/remeshing/distance.hh:42-129)."""

    def __init__(self, grid):
        self._g, self._d, self._max = grid, None, 0.0

    def update(self):
        self._d, self._max = self._g._ctx.distance_update()

    def initialized(self):
        return self._d is not None

    def __getitem__(self, i):
        return float(self._d[i])

    def __call__(self, element_index):
        """distance.hh:92-101: average of the element's vertex distances (id-ordered corners)."""
        m = self._g._m
        c, p = m.cells[element_index], int(m.perm[element_index])
        k = c.shape[0]
        d = 0.0
        for i in range(k):
            d += self._d[c[(p >> (2 * i)) & 3]]
        return d / k

    def maximum(self):
        return self._max

    def size(self):
        return self._d.shape[0]

    def array(self):
        return self._d


class RatioIndicator:
    """Dune::RatioIndicator<Grid> (ratioindicator.hh:62-182): members are read/written like the reference's references."""

    def __init__(self, grid):
        object.__setattr__(self, "_g", grid)
        object.__setattr__(self, "_distance", Distance(grid))

    def init(self):
        self._g._ctx.indicator_init()

    def update(self):
        self._distance.update()

    def distance(self):
        if not self._distance.initialized():
            self._distance.update()
        return self._distance

    def __getattr__(self, name):
        if name in ("minH", "maxH", "factor", "distProportion", "edgeRatio", "radiusRatio"):
            return getattr(self._g._ctx.get_params(), name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in ("minH", "maxH", "factor", "distProportion"):
            p = self._g._ctx.get_params()
            setattr(p, name, float(value))
            self._g._ctx.set_params(p)
        else:
            raise AttributeError("RatioIndicator.%s is not writable (const in the reference)" % name)


class MMesh:
    """Dune::MMesh adaptation API (mmesh.hh:725-1474), numeric part."""

    def __init__(self, snapshot, device=0):
        self._ctx = capi.Context(dim=snapshot.dim, device=device)
        self.setSnapshot(snapshot)
        self._indicator = RatioIndicator(self)
        self._marks = np.zeros(snapshot.nc, np.int8)
        self._imarks = None
        self._longest = None
        self.removed = []

    def setSnapshot(self, m):
        self._m = m
        x = m.xy if m.dim == 2 else m.xyz
        facets = m.segs if m.dim == 2 else m.tris
        self._ctx.upload_mesh(x, m.cells, m.perm, getattr(m, "edges", None), facets)
        if m.dim == 2 and getattr(m, "nb", None) is not None:
            self._ctx.upload_neighbours(m.nb)                      # face->neighbor(k), needed by adaptCandidates
        if m.dim == 2 and getattr(m, "insertion_level", None) is not None:
            self._ctx.upload_vertex_levels(m.insertion_level)

    # ---- movement ----
    def moveVertices(self, shifts):
        """mmesh.hh:1421-1430"""
        s = np.ascontiguousarray(shifts, np.float64)
        assert s.shape == (self._m.nv, self._m.dim)            # the reference asserts the size (mmesh.hh:1423)
        self._ctx.move_vertices(s)

    def ensureVertexMovement(self, shifts):
        """mmesh.hh:1094-1134: True when vertices were scheduled for removal (self.removed); GridError in 3-D."""
        s = np.ascontiguousarray(shifts, np.float64)
        assert s.shape == (self._m.nv, self._m.dim)
        valid, _, ninv = self._ctx.trial_move_validate(s)
        change = False
        m = self._m
        for c in np.nonzero(valid == 0)[0]:
            for j in range(m.dim + 1):
                v = int(m.cells[c][(int(m.perm[c]) >> (2 * j)) & 3])
                if m.vflags[v] & 1 or m.vflags[v] & 2:            # isInterface / boundaryFlag == 1
                    continue
                if v not in self.removed:
                    self.removed.append(v)
                    change = True
        return change

    def moveInterface(self, shifts):
        """mmesh.hh:1407-1416: shifts indexed by the interface leaf vertex index."""
        self._ctx.move_vertices(self._scatter_interface(shifts))

    def ensureInterfaceMovement(self, shifts):
        """Numeric part of mmesh.hh:920-1088; returns the cells that would invert."""
        _, _, ninv = self._ctx.trial_move_validate(np.zeros((self._m.nv, self._m.dim)))
        if ninv:
            raise capi.GridError(capi.MMB_ERR_INVALID_CELL_3D, "A cell has a negative volume! Maybe the interface has been moved too far?")
        self._ctx.trial_move_validate(self._scatter_interface(shifts))
        return self._ctx.invalid_cells()

    def _scatter_interface(self, shifts):
        iv = np.nonzero(self._m.vflags & 1)[0]
        s = np.ascontiguousarray(shifts, np.float64)
        assert s.shape == (iv.shape[0], self._m.dim)
        full = np.zeros((self._m.nv, self._m.dim))
        full[iv] = s
        return full

    # ---- marking ----
    def markElements(self):
        """mmesh.hh:736-751"""
        self._marks, self._longest, counts = self._ctx.mark_elements(run_distance=True)
        self._indicator._distance._d = None                        # refreshed lazily
        change = counts[0] + counts[1] > 0
        if self._m.dim == 2 and self._ctx.ns:
            self._imarks = self._ctx.mark_interface()
            change = change or bool(np.any(self._imarks != 0))
        return bool(change)

    def adaptCandidates(self):
        """insert_ / remove_ of adapt() before adapt_() (mmesh.hh:826-908), 2-D: dict of index lists + insertion points.
        Vertices already in self.removed (ensureVertexMovement) are filtered like removed_.insert(...).second."""
        assert self._m.dim == 2
        ins = self._ctx.refinement_candidates()
        pts, ev = self._ctx.refinement_points(ins) if len(ins) else (np.zeros((0, 2)), np.zeros((0, 2), np.int32))
        rc, rv = self._ctx.coarsening_candidates()
        keep = np.array([v not in self.removed for v in rv.tolist()], bool)
        out = dict(insert_cells=ins, insert_points=pts, insert_edge_vertices=ev,
                   remove_cells=rc[keep], remove_vertices=rv[keep])
        if self._ctx.ns:
            si, rs, rsv = self._ctx.interface_candidates()
            keep = np.array([v not in self.removed for v in rsv.tolist()], bool)
            out.update(insert_segments=si, remove_segments=rs[keep], remove_segment_vertices=rsv[keep])
        return out

    def getMark(self, element_index):
        return int(self._marks[element_index])

    def mark(self, ref_count, element_index):
        self._marks[element_index] = ref_count
        return True

    def preAdapt(self):
        return bool(np.any(self._marks < 0)) or len(self.removed) > 0

    def postAdapt(self):
        self._marks[:] = 0
        self.removed = []

    def restoreDelaunay(self, max_rounds=1000):
        """Delaunay restoration of the resident 2-D snapshot by edge flips on the device (mmb_restore_delaunay; the flips of
        CGAL/Delaunay_triangulation_2.h:936-1022, interface edges kept).  The snapshot object receives the new cells, corner
        permutations, neighbours and edges.  Returns the flip statistics."""
        m = self._m
        res = self._ctx.restore_delaunay(vertex_id=getattr(m, "vid", None), max_rounds=max_rounds)
        if res["n_flips"]:
            m.cells, m.perm, m.nb, edges = self._ctx.download_topology()
            if edges is not None:
                m.edges = edges
            self._marks[:] = 0
        return res

    def indicator(self):
        return self._indicator

    def distance(self):
        return self._indicator.distance()

    def coordinates(self):
        return self._ctx.download_coords()
