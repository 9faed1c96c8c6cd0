/*
 * mmesh_b200.h -- C-ABI of the B200-native moving-mesh hot path (drop-in boundary).
 *
 * The reference (samuelburbulla/dune-mmesh) has no FFI for this path: its boundary is the C++ member API
 * of Dune::MMesh / RatioIndicator / Distance (dune/mmesh/grid/mmesh.hh, dune/mmesh/remeshing/*.hh),
 * re-exported 1:1 to Python by dune/python/mmesh/grid.hh:82-137.  Each entry point below names the
 * reference member it replaces.  Plain C types only; no exceptions cross this boundary; every function
 * returns an mmb_status.  All `*_host` pointers are caller-owned HOST memory; a NULL output pointer means
 * "leave the result resident on the device", a NULL shifts pointer means "use the resident shifts"
 * (mmb_upload_shifts).  One context per GPU; calls on one context must be serialised by the caller
 * (the reference is single-threaded and not re-entrant either, mmesh.hh:608,1787-1789).
 *
 * Snapshot conventions (what a host adapter extracts from a live Dune::MMesh):
 *   x       [nv][dim] doubles, leaf vertex index order (indexsets.hh:242-245); same memory layout as
 *                     std::vector<FieldVector<double,dim>>, so reference `shifts` vectors pass through as-is
 *   cells   [nc][dim+1] int32, face->vertex(k)->info().index in CGAL TDS order (Triangulation_2.h:1152-1154)
 *   perm    [nc] uint8, bits (2k+1:2k) = ElementInfo::cgalIndex[k] (cgal/defaults.hh:29, common.hh:55-65)
 *   edges   [ne][4] int32 (a,b,c,d): edge (a,b); c = apex with (a,b,c) ccw; d = opposite apex, -1 on the hull
 *   facets  [ns][dim] int32 interface facets, corners in DUNE facet order (lower persistent id first)
 */
#ifndef MMESH_B200_H
#define MMESH_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmb_context mmb_context;

typedef enum mmb_status {
  MMB_OK = 0,
  MMB_ERR_INVALID_ARG = 1,
  MMB_ERR_CUDA = 2,
  MMB_ERR_NCCL = 3,
  MMB_ERR_INVALID_CELL_3D = 4, /* the DUNE_THROW(GridError) of mmesh.hh:1104-1107 */
  MMB_ERR_CAPACITY = 5,        /* an index space exceeded, or an exact predicate whose input lies outside the exponent range the
                                  expansion arithmetic can hold (ratio largest / smallest nonzero magnitude > 2^212 for in-circle,
                                  2^300 orient3d, 2^480 orient2d, or non-finite coordinates): the outputs are filled, the affected
                                  signs are 0 and must not be trusted.  Any absolute scale is fine (inputs are rescaled exactly). */
  MMB_ERR_STATE = 6,           /* e.g. mark before distance update, project before upload */
  MMB_MOVE_WITHHELD = 7        /* not an error: the step ran, but ensureVertexMovement found cells that would invert
                                  (n_invalid > 0), so moveVertices was NOT applied -- the caller removes the vertices
                                  (mmesh.hh:1109-1124 -> adapt()) and moves afterwards, like the reference's user loop.
                                  All outputs are filled.  Only with mmb_params.move_policy == 0. */
} mmb_status;

/* Mutable members of RatioIndicator (ratioindicator.hh:165-182,191-199) + the projection constants that are
   hard-coded in the reference (polygoncutting.hh:63, cutsettriangulation.hh:60). */
typedef struct mmb_params {
  double minH, maxH, factor, distProportion;
  double edgeRatio, radiusRatio;  /* 4, 30 */
  double clip_eps;                /* 1e-14 */
  double drop_area;               /* 1e-8 */
  int32_t clip_translate;         /* 0 = reference arithmetic (default); 1 = clip relative to the new cell's vertex 0 */
  int32_t move_policy;            /* fused step only.  0 (default) = reference order of the user loop: when the trial move
                                     inverts a cell, ensureVertexMovement returns true and the vertices are removed BEFORE
                                     moveVertices (doc/examples/moving.py:231-252), so the step withholds the move (decided on
                                     the device, no host round trip) and returns MMB_MOVE_WITHHELD; 1 = always move (stress
                                     runs: the resident mesh may then hold inverted cells). */
} mmb_params;

/* Scalars produced by one adapt step (all reductions are deterministic: fixed-shape trees, no float atomics). */
typedef struct mmb_step_result {
  int64_t n_invalid;      /* cells with signedVolume_ <= 0 after the trial move (mmesh.hh:1102) */
  int64_t n_orient_nonpos;/* cells whose exact orientation sign is <= 0 */
  int64_t n_illegal;      /* edges failing the in-circle test */
  int64_t n_refine, n_coarsen; /* refineMarked_/coarsenMarked_ (mmesh.hh:725-731) */
  int64_t n_exact_orient, n_exact_incircle; /* predicates that needed the exact fallback (diagnostic) */
  int64_t n_pieces;       /* cut triangles kept by the projection */
  double  max_dist;       /* Distance::maximum() (distance.hh:118-123) */
  double  mass_new;       /* sum_K |K| mean(u_K) over projected cells */
  double  covered_area;   /* sum of kept cut-triangle areas */
  int64_t moved;          /* 1 = moveVertices was applied, 0 = withheld (see MMB_MOVE_WITHHELD) */
} mmb_step_result;

typedef struct mmb_flip_result {
  int64_t n_flips;               /* edge flips applied */
  int64_t n_illegal_left;        /* illegal unconstrained edges at the last sweep (0 when converged) */
  int64_t n_constrained_illegal; /* interface edges failing the in-circle test: constraints, never flipped */
  int32_t rounds;                /* mark / select / apply rounds executed */
  int32_t converged;             /* 1 = no illegal unconstrained edge is left */
} mmb_flip_result;

/* ---- lifetime ---------------------------------------------------------------------------------------- */
/* dim = 2 or 3.  stream = a cudaStream_t to launch on (as void*), or NULL for an internal stream. */
int mmb_create(int device, int dim, void* stream, mmb_context** out);
int mmb_destroy(mmb_context* ctx);
const char* mmb_last_error(const mmb_context* ctx);
void* mmb_stream(const mmb_context* ctx);
int mmb_synchronize(mmb_context* ctx);

/* Page-lock caller-owned host memory (e.g. the storage of the std::vector<FieldVector<double,dim>> shifts, or of a dune-fem
   DOF vector) so that the asynchronous copies of mmb_adapt_step_host overlap with the kernels; pageable buffers work too,
   but the driver then stages every copy.  Thin wrappers over cudaHostRegister / cudaHostUnregister; ctx may be NULL. */
int mmb_host_register(mmb_context* ctx, void* ptr, int64_t bytes);
int mmb_host_unregister(mmb_context* ctx, void* ptr);

/* ---- snapshot upload / download (outside the timed path, like the reference's CGAL TDS edits) ------------ */
int mmb_upload_mesh(mmb_context* ctx, int64_t nv, const double* x_host, int64_t nc, const int32_t* cells_host,
                    const uint8_t* perm_host, int64_t ne, const int32_t* edges_host, int64_t ns,
                    const int32_t* facets_host);
/* Incremental snapshot update after a small TDS edit -- what adapt_() changes (mmesh.hh:1338-1355: a few inserted / removed
   vertices, the cells of their conflict zones, the index sets): only the listed entries cross the host link, the resident tables
   grow in place, and the per-vertex distance hints of untouched vertices survive when the interface is unchanged (the next
   Distance::update starts warm).  New sizes nv/nc/ne/ns; vert_idx / vert_x: changed or new vertices; cell_idx / cell_v [n][dim+1]
   / cell_perm: changed or new cells; edge_idx / edge_q [n][4]: changed or new edge quads (2-D); every index >= the old size must
   be listed.  facets = the full new interface list, or NULL when the interface is unchanged (ns_new must then equal the old ns).
   Afterwards: shifts, neighbours, vertex levels, pairs and halo lists must be supplied again, like after mmb_upload_mesh. */
int mmb_patch_mesh(mmb_context* ctx, int64_t nv_new, int64_t n_vert, const int32_t* vert_idx_host, const double* vert_x_host,
                   int64_t nc_new, int64_t n_cell, const int32_t* cell_idx_host, const int32_t* cell_v_host, const uint8_t* cell_perm_host,
                   int64_t ne_new, int64_t n_edge, const int32_t* edge_idx_host, const int32_t* edge_q_host,
                   int64_t ns_new, const int32_t* facets_host);
/* Delaunay restoration on the resident 2-D snapshot by edge flips: what CGAL does edge by edge after an
   insertion -- restore_Delaunay / propagating_flip, CGAL/Delaunay_triangulation_2.h:936-1022, flip =
   Triangulation_data_structure_2::flip, CGAL/Triangulation_data_structure_2.h:885-915 -- as data-parallel rounds over all cells:
   an edge flips when the opposite vertex is strictly inside the circumcircle (exact sign; ON_POSITIVE_SIDE, :969); interface edges
   are constraints (dune/mmesh/cgal/triangulationwrapper.hh:65-125) and stay.  Needs mmb_upload_neighbours; every cell must be
   positively oriented (MMB_ERR_STATE otherwise: run ensureVertexMovement first).  vertex_id = persistent id per vertex for the
   corner permutation of the new cells (NULL: id = index).  At most max_rounds rounds (each flips a maximal set of edges that do not
   touch one another).  Cells, neighbours, corner permutations, the edge list and the tile windows are updated in place
   (mmb_download_topology reads them back); marks, flags, candidates and projection pairs of the old topology are void.
   The result is the constrained Delaunay triangulation: unique for points in general position, so it equals CGAL's. */
int mmb_restore_delaunay(mmb_context* ctx, const int64_t* vertex_id_host, int32_t max_rounds, mmb_flip_result* out);
/* The flips of the last mmb_restore_delaunay, round after round: edge_id = 3 * cell + slot with cell and slot as they were when the
   flip was applied, i.e. the arguments f, i of Triangulation_data_structure_2::flip(f, i); round_off [rounds + 1] delimits the rounds.
   The flips of one round touch disjoint cells, so a host TDS that replays them round by round (any order inside a round) ends up
   slot for slot in the state mmb_download_topology returns.  At most nc flips are logged (n_total reports all of them). */
int mmb_flip_log(mmb_context* ctx, uint32_t* edge_id_host, int64_t cap, int64_t* n_total, int64_t* round_off_host, int32_t round_cap,
                 int32_t* n_rounds);
/* resident topology -> host: cells [nc][dim+1] (TDS order), perm [nc], neighbours [nc][3] (2-D, if uploaded), edges [ne][4]
   (2-D); NULL skips an array */
int mmb_download_topology(mmb_context* ctx, int32_t* cells_host, uint8_t* perm_host, int32_t* cell_nb_host, int32_t* edges_host);
int mmb_upload_coords(mmb_context* ctx, const double* x_host);      /* vh->point() = ... for every vertex */
int mmb_download_coords(mmb_context* ctx, double* x_host);
int mmb_upload_shifts(mmb_context* ctx, const double* shifts_host); /* make `shifts` resident */
/* resident shifts *= a (the `s *= -1.0` of mmesh.hh:1130 when a == -1; exact for +-1) */
int mmb_scale_shifts(mmb_context* ctx, double a);
int mmb_set_params(mmb_context* ctx, const mmb_params* p);
int mmb_get_params(const mmb_context* ctx, mmb_params* p);

/* ---- the hot path ---------------------------------------------------------------------------------------- */
/* RatioIndicator::init (ratioindicator.hh:62-91): minH/maxH/factor from interface + bulk edge lengths. */
int mmb_indicator_init(mmb_context* ctx);

/* MMesh::moveVertices (mmesh.hh:1421-1430): x <- x (+) shifts. */
int mmb_move_vertices(mmb_context* ctx, const double* shifts_host);

/* Numeric part of MMesh::ensureVertexMovement (mmesh.hh:1094-1134): evaluates every cell on x (+) shifts without
   modifying x (restore = exact copy).  valid_fp[c] = !(signedVolume_ <= 0.0) bit-exact to the reference test;
   orient_sign[c] = CGAL::orientation (Epick) in {-1,0,+1}.  dim==3 with an invalid cell returns
   MMB_ERR_INVALID_CELL_3D after filling the outputs (mmesh.hh:1104-1107). */
int mmb_trial_move_validate(mmb_context* ctx, const double* shifts_host, uint8_t* valid_fp_host,
                            int8_t* orient_sign_host, int64_t* n_invalid);
/* indices of the invalid cells in increasing order (<= cap entries written; returns the total count) */
int mmb_invalid_cells(mmb_context* ctx, int32_t* cells_host, int64_t cap, int64_t* n_total);

/* Per-edge Delaunay legality on the current coordinates: flag 1 legal, 0 illegal, 2 hull edge.
   Predicate: CGAL Side_of_oriented_circle_2 (Delaunay_triangulation_2.h:663-681). */
int mmb_legality(mmb_context* ctx, uint8_t* flags_host, int64_t* n_illegal);

/* Distance::update + maximum (distance.hh:42-59,118-123), exact same per-pair arithmetic, BVH-culled. */
int mmb_distance_update(mmb_context* ctx, double* dist_host, double* max_out);
/* Distance::set for all vertices (distance.hh:84-90): overwrite the resident field. */
int mmb_distance_set(mmb_context* ctx, const double* dist_host);

/* RatioIndicator::update + MMesh::markElements bulk loop (ratioindicator.hh:94-160, mmesh.hh:736-743) and the
   argmax of LongestEdgeRefinement::refinement (longestedgerefinement.hh:41-57).  counts = {refine, coarsen}.
   run_distance != 0 performs indicator_.update() (distance update) first, like the reference. */
int mmb_mark_elements(mmb_context* ctx, int run_distance, int8_t* marks_host, uint8_t* longest_edge_host,
                      int64_t counts[2]);
/* interface elements (mmesh.hh:745-748): marks per facet */
int mmb_mark_interface(mmb_context* ctx, int8_t* marks_host);
/* ordered compaction of marked cells (index order, as MMesh::adapt() visits them, mmesh.hh:832-864) */
int mmb_marked_cells(mmb_context* ctx, int which /* +1 refine, -1 coarsen */, int32_t* cells_host, int64_t cap,
                     int64_t* n_total);

/* Conservative projection old -> new cells (numeric part of MMesh::adapt(handle), mmesh.hh:1138-1165 with
   CutSetTriangulation cutsettriangulation.hh:31-62 and PolygonCutting polygoncutting.hh:17-137).
   Pair CSR: new_off[nnew+1]; new_cell[nnew] indexes the uploaded mesh; old_xy[P][6] = CachingEntity::vertex_
   (cachingentity.hh:83-89); old_idx[P] indexes dofs_old.  ncomp = 1 (P0/FV) or 3 (DG-P1, nodal, id-order corners).
   dofs_new has ncomp entries for every cell of the uploaded mesh; only projected cells are written. */
int mmb_upload_pairs(mmb_context* ctx, int64_t nnew, const int64_t* new_off_host, const int32_t* new_cell_host,
                     int64_t npairs, const double* old_xy_host, const int32_t* old_idx_host);
int mmb_project(mmb_context* ctx, int ncomp, int64_t n_old, const double* dofs_old_host, double* dofs_new_host,
                double* mass_new, double* covered_area, int64_t* n_pieces);
/* MMeshCachingEntity::intersectionVolume (cachingentity.hh:154-167) for every uploaded pair */
int mmb_intersection_volumes(mmb_context* ctx, double* vol_host);

/* MMeshInterfaceGrid::adapt(handle) (interface/grid.hh:388-415) for FV (P0) interface data: per new interface element
   the old children with their lengths; the longer of (old, new) is the father. */
int mmb_project_interface_p0(mmb_context* ctx, int64_t nnew, const double* new_len_host, const int64_t* new_off_host,
                             int64_t npairs, const double* old_len_host, const int32_t* old_idx_host, int64_t n_old,
                             const double* dofs_old_host, double* dofs_new_host);
/* The same for DG-P1 interface data (two nodal DOFs per element on its id-ordered corners): new_xy / old_xy hold the two
   corners (x0, y0, x1, y1) of every new element / old child; old longer => the father's function evaluated at the son's
   corners through geometryInFather (interface/entity.hh:491-507), else the local L2 projection of the son onto the
   father's P1 space.  dofs_old [n_old][2], dofs_new [nnew][2]. */
int mmb_project_interface_p1(mmb_context* ctx, int64_t nnew, const double* new_xy_host, const int64_t* new_off_host,
                             int64_t npairs, const double* old_xy_host, const int32_t* old_idx_host, int64_t n_old,
                             const double* dofs_old_host, double* dofs_new_host);
/* LongestEdgeRefinement::refinement (longestedgerefinement.hh:41-57) for the listed cells (after mmb_mark_elements):
   centre of the longest edge ([n][dim]) and, optionally, its two vertex indices (id order).  2-D and 3-D. */
int mmb_refinement_points(mmb_context* ctx, int64_t n, const int32_t* cells_host, double* points_host, int32_t* edge_vertices_host);

/* Refinement candidates of MMesh::adapt() (mmesh.hh:828-851): the cells (index order) whose longest edge enters insert_,
   i.e. marked +1, edge not an interface edge, edge id not yet inserted by an earlier cell.  Needs the face neighbours
   (cell_nb[c][k] = index of face->neighbor(k), -1 for the infinite face).  Feed the result to mmb_refinement_points.
   3-D meshes need no neighbours: an edge has many cells, so inserted_ is a device hash table keyed by the edge's vertex
   pair holding the lowest claiming cell; isInterface(edge) = edge of an interface triangle (mmesh.hh:624-645). */
int mmb_upload_neighbours(mmb_context* ctx, const int32_t* cell_nb_host);
int mmb_refinement_candidates(mmb_context* ctx, int32_t* cells_host, int64_t cap, int64_t* n_total);

/* Coarsening candidates of MMesh::adapt() (mmesh.hh:853-865): for every cell with mark -1, in index order, the vertex
   LongestEdgeRefinement::coarsening picks (longestedgerefinement.hh:64-112: per-corner sum of element edge lengths,
   insertionLevel, interface/boundary constraint; std::sort of libstdc++ reproduced), kept only when
   removed_.insert(id) succeeds, i.e. for the first cell choosing that vertex.  removed_ is taken as empty on entry (the
   host filters against what ensureVertexMovement already put there).  Output: (cell, vertex) pairs in cell order.
   Boundary vertices (atBoundary, :114-120) are derived from the hull edges of the uploaded edge quads, interface
   vertices from the facets; mmb_upload_vertex_levels supplies VertexInfo::insertionLevel (default 0, mmesh.hh:132).
   Single-GPU snapshots only: a shard's edge list covers its owned edges, not the halo's. */
int mmb_upload_vertex_levels(mmb_context* ctx, const int32_t* insertion_level_host);
int mmb_coarsening_candidates(mmb_context* ctx, int32_t* cells_host, int32_t* vertices_host, int64_t cap, int64_t* n_total);
/* Interface loop of MMesh::adapt() (mmesh.hh:869-908) after mmb_mark_interface: insert_ = interface elements with
   mark +1 (refinement point = element centre); remove_ = for mark -1 the first element vertex with exactly two
   incident interface elements (isRemoveable, longestedgerefinement.hh:160-168), dropped when the bulk loop or an
   earlier interface element already removed it.  Facet rows are the interface elements' vertex order. */
int mmb_interface_candidates(mmb_context* ctx, int32_t* insert_segs_host, int64_t cap_insert, int64_t* n_insert,
                             int32_t* remove_segs_host, int32_t* remove_vertices_host, int64_t cap_remove, int64_t* n_remove);

/* The component numbering of MMesh::adapt_() before its TDS edits (mmesh.hh:1194-1218): getComponentNumber_ (:1919-1941)
   applied to every zone -- the conflict zone of each insertion point, then the star of each vertex to remove, in the order of
   insert_ / remove_ -- and repeated `while (markAgain)`.  zones = CSR over cell indices of the uploaded mesh; component_count =
   componentCount_ on entry (connectedComponents_.size() + 1, i.e. 1 after postAdapt; all ElementInfo::componentNumber are
   taken as 0 on entry).  zone_number[k] is what insertComponentIds / removeComponentIds hold (+ 1); cell_number[c] = the
   cell's componentNumber afterwards (0 = in no zone); *component_count_out = componentCount_ afterwards.  The order-dependent
   loop is evaluated in closed form on the device (first zone per cell, scan over the fresh zones, min-label propagation). */
int mmb_component_numbers(mmb_context* ctx, int64_t nz, const int64_t* zone_off_host, const int32_t* zone_cells_host,
                          int64_t component_count, int32_t* zone_number_host, int32_t* cell_number_host,
                          int64_t* component_count_out);

/* One whole adapt step on resident data, in the reference's user-loop order (doc/examples/moving.py:231-252):
   markElements (distance + marks) -> ensureVertexMovement (trial validity) -> adapt(handle) projection of the
   uploaded pairs -> moveVertices -> in-circle legality of the moved mesh (input of the next host TDS pass).
   ncomp = 0 skips the projection.  res may be NULL (no host synchronisation at all). */
int mmb_adapt_step(mmb_context* ctx, const double* shifts_host, int ncomp, mmb_step_result* res);

/* The same step with HOST buffers for every per-step field (the drop-in call a host adapter makes each step):
   shifts and the old DOFs go in, marks / longest-edge / valid_fp / orient_sign / edge flags / new DOFs and the scalars
   come out.  Any output pointer may be NULL.  dofs_old_host == NULL keeps the resident DOFs. */
int mmb_adapt_step_host(mmb_context* ctx, const double* shifts_host, int ncomp, int64_t n_old, const double* dofs_old_host,
                        double* dofs_new_host, int8_t* marks_host, uint8_t* longest_edge_host, uint8_t* valid_fp_host,
                        int8_t* orient_sign_host, uint8_t* edge_flags_host, mmb_step_result* res);
/* The same pipelined step in three calls, for a sharded caller that must place its collectives in between:
     mmb_step_host_begin   async H2D of shifts and old DOFs (own copy stream), Distance::update
       -> all-reduce MAX of the distance maximum (mmb_device_buffer(2))
     mmb_step_host_marks   marks (+ async D2H); afterwards the launching stream is ordered behind the shift upload
       -> halo exchange of the shifts (mmb_halo_pack_shifts / all-to-all / mmb_halo_unpack_shifts)
     mmb_step_host_finish  trial validity, chunked projection with the new DOFs streaming back, moveVertices, legality,
                           D2H of the flags; reduce_vector != 0 also fills mmb_device_buffer(3) for the SUM all-reduce
                           and returns without synchronising (the caller's read of that vector does). */
int mmb_step_host_begin(mmb_context* ctx, const double* shifts_host, int ncomp, int64_t n_old, const double* dofs_old_host,
                        int will_download_dofs);
int mmb_step_host_marks(mmb_context* ctx, int8_t* marks_host, uint8_t* longest_edge_host);
int mmb_step_host_finish(mmb_context* ctx, double* dofs_new_host, uint8_t* valid_fp_host, int8_t* orient_sign_host,
                         uint8_t* edge_flags_host, int reduce_vector, mmb_step_result* res);

/* Asynchronous field transfers on the context's stream (pinned host buffers; complete after mmb_synchronize).
   Used by the sharded step, where collectives sit between the phases. */
int mmb_upload_dofs(mmb_context* ctx, int ncomp, int64_t n_old, const double* dofs_old_host);
int mmb_download_fields(mmb_context* ctx, int ncomp, double* dofs_new_host, int8_t* marks_host, uint8_t* longest_edge_host,
                        uint8_t* valid_fp_host, int8_t* orient_sign_host, uint8_t* edge_flags_host);

/* ---- multi-GPU plumbing (one context per rank; the collectives themselves are issued by the host, e.g.
   torch.distributed / NCCL, on the buffers exposed by mmb_device_buffer) ------------------------------------------
   The mesh is sharded by contiguous Hilbert ranges of cells (the reference's own scheme is contiguous ranges of the
   iteration order, dune/mmesh/misc/partitionhelper.hh:175-200) with a one-ring halo.  Halo cells are evaluated
   redundantly but only cells in [lo, hi) enter the counters. */
int mmb_set_owned_cells(mmb_context* ctx, int64_t lo, int64_t hi);
/* vertex halo: send_idx = local indices of owned vertices other ranks need (concatenated per peer), recv_idx = local
   indices of the halo vertices in the order their values arrive */
int mmb_halo_setup(mmb_context* ctx, int64_t n_send, const int32_t* send_idx_host, int64_t n_recv, const int32_t* recv_idx_host);
int mmb_halo_pack_shifts(mmb_context* ctx);      /* resident shifts[send_idx] -> send buffer */
int mmb_halo_unpack_shifts(mmb_context* ctx);    /* recv buffer -> resident shifts[recv_idx] */
/* which: 0 halo send buffer, 1 halo recv buffer (double2 per vertex), 2 the max-distance word (uint64, all-reduce MAX),
   3 the additive step scalars (double[16], all-reduce SUM; layout of mmb_step_result without max_dist) */
int mmb_device_buffer(mmb_context* ctx, int which, void** device_ptr, int64_t* bytes);
int mmb_step_phase_distance(mmb_context* ctx);            /* counters = 0; Distance::update on the local vertices */
int mmb_step_phase_rest(mmb_context* ctx, int ncomp);     /* mark -> validate -> project -> move -> legality */
/* the same split so that collectives overlap compute: which = 1 validity + projection (no dependence on the global max
   distance: run it while the MAX all-reduce is in flight), 2 marks, 3 move + legality + reduce vector, 0 = all */
int mmb_step_phase(mmb_context* ctx, int which, int ncomp);

/* ---- the sharded step with its collectives INSIDE the library (NCCL over NVLink / NVSwitch) ----------------------------------
   Replaces the reference's raw MPI point-to-point layer for this path (dune/mmesh/misc/communication.hh:83-127,
   partitionhelper.hh:175-236).  One context per rank / GPU.  libnccl.so.2 is loaded at run time (the library does not link
   against it): failures of the load or of any NCCL call return MMB_ERR_NCCL with the NCCL message in mmb_last_error.
     mmb_comm_unique_id   ncclGetUniqueId on one rank; the caller distributes the 128 bytes (MPI, a file, torch.distributed)
     mmb_comm_init        ncclCommInitRank on the context's device; from then on the context is one shard of `world`
     mmb_halo_counts      per-peer sizes of the halo lists given to mmb_halo_setup (send_counts[p] vertices go to rank p,
                          recv_counts[p] arrive from it, in the order of send_idx / recv_idx)
     mmb_adapt_step_sharded  the whole adapt step of mmb_adapt_step on this shard, in the reference's order, with
         halo exchange of the shifts (grouped ncclSend/ncclRecv) overlapped with Distance::update,
         all-reduce MAX of Distance::maximum() overlapped with trial validity + projection,
         all-reduce SUM of n_invalid (the device-side gate of moveVertices, identical on every rank),
         one all-reduce SUM of the step's additive scalars (counts < 2^53 and the mass as doubles);
       res (may be NULL) receives the GLOBAL scalars on every rank.  shifts_host = this shard's local vertices (owned entries
       are used, halo entries are overwritten by the exchange) or NULL for the resident shifts.  The kernels and collectives
       of a step are captured once into a CUDA graph and replayed (MMB_GRAPH=0 disables that); no host synchronisation
       happens when res == NULL. */
int mmb_comm_unique_id(uint8_t id_out[128]);
int mmb_comm_init(mmb_context* ctx, int world, int rank, const uint8_t id[128]);
int mmb_comm_destroy(mmb_context* ctx);
int mmb_halo_counts(mmb_context* ctx, const int64_t* send_counts, const int64_t* recv_counts);
int mmb_adapt_step_sharded(mmb_context* ctx, const double* shifts_host, int ncomp, mmb_step_result* res);

/* ---- 3-D extras -------------------------------------------------------------------------------------- */
/* P0 interface-coupled transfer (dune/python/mmesh/pyskeletontrace.hh:42-64,133-160):
   trace_in[f] = u[inside(f)], trace_out[f] = u[outside(f)], skeleton[f] = u_gamma[f]. */
int mmb_trace_skeleton_p0(mmb_context* ctx, int64_t nf, const int32_t* inside_host, const int32_t* outside_host,
                          const double* u_bulk_host, int64_t n_bulk, const double* u_gamma_host,
                          double* trace_in_host, double* trace_out_host, double* skeleton_host);

/* ---- instrumentation --------------------------------------------------------------------------------- */
/* Per-kernel CUDA-event timing of the launches issued since the last mmb_timing_reset (on ctx's stream). */
int mmb_timing_enable(mmb_context* ctx, int on);
int mmb_timing_reset(mmb_context* ctx);
/* returns the number of distinct kernels; fills up to cap entries (name pointers stay valid for ctx's life) */
int mmb_timing_get(mmb_context* ctx, int cap, const char** names, double* total_ms, int64_t* launches);
int64_t mmb_launch_count(const mmb_context* ctx); /* kernels launched by this context so far */

/* predicate batches on raw points (config 4 stress; pts [n][k][2|3]); which: 0 orient2d, 1 incircle, 2 orient3d */
int mmb_predicate_batch(mmb_context* ctx, int which, int64_t n, const double* pts_host, int8_t* sign_host,
                        int64_t* n_exact);

#ifdef __cplusplus
}
#endif
#endif /* MMESH_B200_H */
