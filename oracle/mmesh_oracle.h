/*
 * mmesh_oracle.h -- CPU ORACLE for the Dune-MMesh moving-mesh hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * path (dune-mmesh_b200/, include/mmesh_b200.h) never links, imports or calls it.
 *
 * It is a plain-C restatement of the reference algorithms, op for op, with the reference
 * file:line each function follows (paths relative to the reference checkout).  The reference
 * itself cannot be compiled here (no Boost / GMP / DUNE), so parity is
 * pinned against (i) the golden values the reference's own tests assert (tests/golden/),
 * (ii) python fractions.Fraction for every exact sign, (iii) the reference tests' tolerance
 * properties (distance to y=0.5, clip area vs exact, conservation).  Bit-level op order of
 * un-vendored dune-geometry/dune-common/dune-fem arithmetic is "parity unpinned" (restated
 * from their published sources).
 *
 * All arrays are caller-owned.  Layout conventions are the ones of include/mmesh_b200.h:
 *   xy      [nv][2]  interleaved doubles (same memory layout as std::vector<FieldVector<double,2>>)
 *   cells   [nc][3]  int32 vertex indices in CGAL TDS (ccw) order, face->vertex(0..2)
 *   perm    [nc]     uint8, bits (2k+1:2k) = cgalIndex[k] : id-sorted corner k is TDS vertex cgalIndex[k]
 *   edges   [ne][4]  int32 (a,b,c,d): edge (a,b), c = apex with (a,b,c) ccw, d = opposite apex or -1
 *   segs    [ns][2]  int32 interface segment corners in DUNE facet order (lower persistent id first)
 */
#ifndef MMESH_ORACLE_H
#define MMESH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
  double minH, maxH, factor, distProportion; /* ratioindicator.hh:191-199 */
  double edgeRatio, radiusRatio;             /* 4, 30 (ratioindicator.hh:45-57) */
  double clip_eps;                           /* 1e-14 polygoncutting.hh:63 */
  double drop_area;                          /* 1e-8  cutsettriangulation.hh:60 */
  int32_t clip_translate;                    /* 0 = reference arithmetic; 1 = translate to new-cell corner 0 first */
  int32_t pad_;
} orc_params;

/* ---- exact predicates (true sign; CGAL Epick result by construction) ------------------- */
int orc_orient2d(const double* p, const double* q, const double* r);
int orc_orient3d(const double* p, const double* q, const double* r, const double* s);
int orc_incircle(const double* p, const double* q, const double* r, const double* t);
/* static-filter stage only: returns -1/0/+1, or 2 when the CGAL semi-static filter cannot decide */
int orc_orient2d_filter(const double* p, const double* q, const double* r);
int orc_orient3d_filter(const double* p, const double* q, const double* r, const double* s);
int orc_incircle_filter(const double* p, const double* q, const double* r, const double* t);
/* batch: pts [n][k][dim] */
void orc_orient2d_batch(const double* pts, int64_t n, int8_t* sign);
void orc_orient3d_batch(const double* pts, int64_t n, int8_t* sign);
void orc_incircle_batch(const double* pts, int64_t n, int8_t* sign);

/* ---- move / validity --------------------------------------------------------------------- */
void orc_move(double* x, const double* shift, int64_t n_doubles);                 /* mmesh.hh:1421-1430 */
/* trial move: validity evaluated on x (+) s, coordinates not modified */
void orc_trial_validate_2d(const double* xy, const double* shift /*may be NULL*/, int64_t nv,
                           const int32_t* cells, int64_t nc,
                           uint8_t* valid_fp, int8_t* orient_sign, int64_t* n_invalid);
void orc_trial_validate_3d(const double* xyz, const double* shift, int64_t nv,
                           const int32_t* cells /*[nc][4]*/, int64_t nc,
                           uint8_t* valid_fp, int8_t* orient_sign, int64_t* n_invalid);
double orc_signed_area(const double* p, const double* q, const double* r);        /* Compute_area_2 */
double orc_signed_volume(const double* p0, const double* p1, const double* p2, const double* p3);

/* ---- per-edge Delaunay legality ---------------------------------------------------------- */
/* flag: 1 legal, 0 illegal (incircle(a,b,c,d) > 0), 2 boundary (d < 0) */
void orc_legality(const double* xy, const double* shift /*may be NULL*/, const int32_t* edges, int64_t ne,
                  uint8_t* flag, int64_t* n_illegal);

/* ---- distance field ---------------------------------------------------------------------- */
double orc_point_segment_distance(const double* vp, const double* v0, const double* v1); /* distance.hh:143-156 */
void orc_distance_2d(const double* xy, int64_t nv, const int32_t* segs, int64_t ns, double* dist); /* brute force */
/* brute force restricted to a vertex subset (bounded CPU-baseline samples / full-size spot checks) */
void orc_distance_2d_subset(const double* xy, const int32_t* vidx, int64_t nsub,
                            const int32_t* segs, int64_t ns, double* dist_sub);
void orc_distance_3d(const double* xyz, int64_t nv, const int32_t* tris /*[ns][3]*/, int64_t ns, double* dist);
double orc_distance_max(const double* dist, int64_t nv);                           /* distance.hh:118-123 */

/* ---- RatioIndicator ---------------------------------------------------------------------- */
void orc_indicator_init_2d(const double* xy, const int32_t* edges, int64_t ne,
                           const int32_t* segs, int64_t ns, orc_params* prm);     /* ratioindicator.hh:62-91 */
void orc_mark_2d(const double* xy, const double* dist, const int32_t* cells, const uint8_t* perm, int64_t nc,
                 const orc_params* prm, double maxDist,
                 int8_t* mark, uint8_t* longest_edge, int64_t counts[2] /*refine, coarsen*/);
void orc_mark_3d(const double* xyz, const double* dist, const int32_t* cells, const uint8_t* perm, int64_t nc,
                 const orc_params* prm, double maxDist,
                 int8_t* mark, uint8_t* longest_edge, int64_t counts[2]);
/* interface elements: mark by length only */
void orc_mark_interface_2d(const double* xy, const int32_t* segs, int64_t ns, const orc_params* prm,
                           double maxDist, int8_t* mark);
/* geometry helpers pinned by the reference's golden tests */
void orc_cell_geometry_2d(const double* xy, const int32_t* cell, uint8_t perm,
                          double* center2, double* volume, double* edge_centers6, double* circumcenter2);
void orc_cell_geometry_3d(const double* xyz, const int32_t* cell, uint8_t perm,
                          double* center3, double* volume, double* circumcenter3);

/* ---- clipping / projection --------------------------------------------------------------- */
/* Sutherland-Hodgman (polygoncutting.hh:17-96): subject/clip triangles, out <= 7 points; returns count */
int orc_sutherland_hodgman(const double* subj6, const double* clip6, double eps, double* out14);
double orc_polygon_area(const double* pts, int n);                                 /* polygoncutting.hh:121-137 */
double orc_intersection_volume(const double* old6, const double* new6);            /* cachingentity.hh:154-167 */
/* CutSetTriangulation (cutsettriangulation.hh:31-62): returns number of kept fan triangles (<=5),
   tris [k][6] */
int orc_cutset(const double* old6_cached, const double* new6_tds, const orc_params* prm, double* tris30);

/* pairs CSR: new_off[nnew+1], new_cell[nnew] (index into cells/perm/xy), old_xy[P][6] cached corners
   (id order), old_idx[P] index into dofs_old.  ncomp = 1 (P0) or 3 (DG-P1, nodal on id-order corners). */
void orc_project(const double* xy, const int32_t* cells, const uint8_t* perm,
                 const int64_t* new_off, const int32_t* new_cell, int64_t nnew,
                 const double* old_xy, const int32_t* old_idx,
                 const double* dofs_old, double* dofs_new, int ncomp,
                 const orc_params* prm, double* mass_new, double* covered_area, int64_t* n_pieces);

/* interface (1-D) transfer, interface/grid.hh:388-415 with P0 data */
void orc_project_interface_p0(const double* new_len, const int64_t* new_off, int64_t nnew,
                              const double* old_len, const int32_t* old_idx,
                              const double* dofs_old, double* dofs_new);

/* ---- MMesh::adapt() candidate loops (mmesh.hh:826-908) -------------------------------------------------------
   LongestEdgeRefinement::coarsening (longestedgerefinement.hh:64-112) on the id-ordered corners of one triangle:
   returns the element-local corner index std::sort puts first.  The comparator is not a strict weak order, so the
   result depends on the sort: this follows libstdc++'s insertion sort (n <= 16), the reference's toolchain;
   tests/native/sort_pin.cc pins it against the real std::sort. */
void orc_coarsening_order(int n, const int64_t* level, const double* len, const uint8_t* constrained, int* order);
int orc_coarsening_choice_2d(const double corners6[6], const int64_t level[3], const uint8_t constrained[3]);
/* Both loops of adapt() up to (not including) adapt_():
   bulk: mark +1 -> insert_ entry unless the longest edge is an interface edge or its id is already in inserted_;
         mark -1 -> remove_ entry unless the chosen vertex is already in removed_ (both sets start empty);
   interface: mark +1 -> insert_ (segment midpoint), mark -1 -> first vertex with exactly two incident interface
         elements (isRemoveable), deduplicated against removed_ (which the bulk loop filled first).
   cell_nb[c][k] = face->neighbor(k) (-1 = infinite face); vertex_level = VertexInfo::insertionLevel.
   Output arrays must hold nc (bulk) / ns (interface) entries; imark may be NULL (interface loop skipped). */
void orc_adapt_candidates_2d(const double* xy, int64_t nv, const int32_t* cells, const uint8_t* perm, int64_t nc,
                             const int32_t* cell_nb, const int8_t* mark, const uint8_t* longest_edge,
                             const int32_t* segs, int64_t ns, const int8_t* imark, const int32_t* vertex_level,
                             int32_t* ins_cells, int64_t* n_ins, int32_t* rem_cells, int32_t* rem_vertices, int64_t* n_rem,
                             int32_t* ins_segs, int64_t* n_ins_segs, int32_t* rem_segs, int32_t* rem_seg_vertices,
                             int64_t* n_rem_segs);

/* 3-D bulk loop of adapt() (mmesh.hh:828-851): cells whose longest edge enters insert_ (+ its two vertices, id order) */
void orc_refine_candidates_3d(const int32_t* cells, const uint8_t* perm, int64_t nc, const int8_t* mark,
                              const uint8_t* longest_edge, const int32_t* tris, int64_t ns,
                              int32_t* ins_cells, int32_t* ins_edge_vertices, int64_t* n_ins);

/* one clip edge q1 -> q2 (polygoncutting.hh:31-93) on a general polygon of n <= 6 points; out holds <= 9 points */
int orc_clip_edge(const double* poly, int n, const double* q1, const double* q2, double eps, double* out);
/* many triangle pairs at once (point lists padded to 9 points) */
void orc_sutherland_hodgman_batch(const double* subj, const double* clip, int64_t n, double eps, double* out18, int32_t* count);

/* OpenMP threads used by the batch loops (the reference is single-threaded per rank; 1 by default) */
void orc_set_threads(int n);
int orc_num_threads(void);

/* labelled CULLED variant of Distance::update for a vertex subset (grid buckets + ring search; same values as the brute force) */
void orc_distance_2d_culled_subset(const double* xy, int64_t nv_all, const int32_t* vidx, int64_t nsub, const int32_t* segs,
                                   int64_t ns, double* dist_sub);

/* RatioIndicator::init in 3-D (ratioindicator.hh:62-91 is dimension generic): interface edges = the three edges of every
   interface triangle, bulk edges = the six edges of every tetrahedron (min / max do not depend on multiplicity or order) */
void orc_indicator_init_3d(const double* xyz, const int32_t* cells, int64_t nc, const int32_t* tris, int64_t ns, orc_params* prm);
/* interface elements of a 3-D grid through RatioIndicator::operator() (mmesh.hh:745-748): distance 0 => l = 0; the coarsening
   criteria cannot fire because dim == dimensionworld == 3 (ratioindicator.hh:151-153) => +1 iff an edge is longer than maxH */
void orc_mark_interface_3d(const double* xyz, const int32_t* tris, int64_t ns, const orc_params* prm, double maxDist, int8_t* mark);

/* MMeshInterfaceGrid::adapt(handle) (interface/grid.hh:388-415) with DG-P1 data on the interface segments (2 nodal DOFs per
   element on its id-ordered corners).  new_xy / old_xy: the two corners (x0, y0, x1, y1) of every new element / old child.
   father = the longer one (volume() = two_norm of c1 - c0):
     old longer  -> prolongLocal(father = old, son = new, true): the father's function evaluated at the son's corners through
                    geometryInFather (interface/entity.hh:491-507: father.geometry().local(corner)), overwriting;
     else        -> restrictLocal(father = new, son = old, initialize): local L2 projection of the son's function on the
                    father's P1 space over the son, M_f^-1 int_son u_son phi_i, set (initialize) or added.
   dune-fem / dune-geometry are un-vendored: AffineGeometry<1,2>::local(x) is taken as ((x - c0) . d) / (d . d), d = c1 - c0,
   and the transfer arithmetic is this definition (parity unpinned; pins: linears reproduced, mass conserved). */
void orc_project_interface_p1(const double* new_xy, const int64_t* new_off, int64_t nnew, const double* old_xy, const int32_t* old_idx,
                              const double* dofs_old, double* dofs_new);

/* ---- MMesh::ensureInterfaceMovement, 2-D (dune/mmesh/grid/mmesh.hh:920-1088) ------------------------ */
/* Statement-by-statement restatement on index arrays.  x = current vertex coordinates, cells TDS order + perm (sub-entity
   order = id order), vflags bit0 isInterface / bit1 boundaryFlag == 1, facets = interface elements (vertex indices, facet
   order), iverts[i] = bulk vertex of interface vertex i, ishifts[niv][2], level = insertionLevel or NULL, removed_mask[nv] =
   removed_ on entry (updated).  Outputs: removed[<= nv] (appended to remove_, in order), ins_* per insertion point
   (edge vertices, point, v0, insertionLevel, crossing facet, connectedcomponent == interface edge).
   Returns 0 / 1 = `change`, or the GridError: -1 "A cell has a negative volume" (:926-930), -2 "more than one interface
   edge" (:970-974), -3 "without any interface edge" (:1001-1005), -4 "two interfaces cross" (:1023-1026). */
int orc_ensure_interface_movement_2d(const double* x, int64_t nv, const int32_t* cells, const uint8_t* perm, int64_t nc,
                                     const uint8_t* vflags, const int32_t* facets, int64_t ns, const int32_t* iverts, int64_t niv,
                                     const double* ishifts, const int32_t* level, uint8_t* removed_mask, int32_t* removed,
                                     int64_t* n_removed, int32_t* ins_edge, double* ins_point, int32_t* ins_v0, int64_t* ins_level,
                                     int32_t* ins_crossing, uint8_t* ins_comp_is_edge, int64_t* n_inserted);

/* ---- component numbers of adapt_() ------------------------------------------------------------- */
/* dune/mmesh/grid/mmesh.hh:1194-1218 (the `while (markAgain)` loop over insert_ then remove_) with getComponentNumber_
   (:1919-1941) and the marking of the zone's elements (:1862-1872, :1889-1903).  zones = CSR over cell indices in the order the
   reference visits them; cell_number must hold ElementInfo::componentNumber on entry (0 after postAdapt) and holds it on
   exit; zone_number[k] = the componentNumber the LAST pass returned for zone k; returns componentCount_ afterwards. */
int64_t orc_component_numbers(int64_t nz, const int64_t* zone_off, const int32_t* zone_cells, int64_t component_count,
                              int32_t* zone_number, int32_t* cell_number);

/* ---- Delaunay restoration by edge flips (sequential Lawson) --------------------------------------- */
/* What CGAL does after an insertion, applied to every edge: an edge whose opposite vertex is strictly inside the circumcircle
   (side_of_oriented_circle == ON_POSITIVE_SIDE, CGAL/Delaunay_triangulation_2.h:969,999) is flipped
   (Triangulation_data_structure_2::flip, CGAL/Triangulation_data_structure_2.h:885-915) and its four outer edges are
   re-examined (propagating_flip, :984-1005); interface edges (segs) are constraints and never flip.  cells [nc][3] (TDS order,
   ccw) and nb [nc][3] (neighbour opposite vertex k, -1 = none) are updated in place.  Returns the number of flips, or -1 when
   max_flips was exceeded. */
int64_t orc_restore_delaunay_2d(const double* xy, int32_t* cells, int32_t* nb, int64_t nc, const int32_t* segs, int64_t ns,
                                int64_t max_flips);

#ifdef __cplusplus
}
#endif
#endif
