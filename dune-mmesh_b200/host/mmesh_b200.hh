// mmesh_b200.hh -- C++ host side of the B200 hot path with the reference's member names.
//
// The reference's boundary for this path is the member API of Dune::MMesh<HostGrid,dim>, RatioIndicator and
// Distance (dune/mmesh/grid/mmesh.hh, dune/mmesh/remeshing/{ratioindicator,distance}.hh).  DUNE/CGAL/Boost are not
// available in this build environment, so instead of deriving from Dune::MMesh this header provides a
// self-contained class that keeps the same member names, argument meaning and error behaviour on top of the
// C-ABI (include/mmesh_b200.h).  A host adapter that owns a live Dune::MMesh fills `Snapshot` from the CGAL
// triangulation after every TDS edit (INTEGRATION.md shows the loops) and forwards the calls below.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mmesh_b200.h"
#include "snapshot.hh"
#include "pairs.hh"

namespace Dune {
namespace B200 {

// insert_ / remove_ of MMesh::adapt() before adapt_() (mmesh.hh:826-908), as index lists
struct AdaptCandidates {
  std::vector<int32_t> insertCells;                       // cells whose longest edge is bisected (element order)
  std::vector<std::array<double, 2>> insertPoints;        // RefinementInsertionPoint::point (edge centre)
  std::vector<std::array<int32_t, 2>> insertEdgeVertices; // ip.v0, ip.v1
  std::vector<int32_t> removeCells, removeVertices;       // bulk loop: (element, vertex to remove)
  std::vector<int32_t> insertSegments;                    // interface elements to bisect
  std::vector<int32_t> removeSegments, removeSegmentVertices;
};

// ---- ensureInterfaceMovement, the branch for cells that invert under the interface movement (mmesh.hh:935-1078) ----
// Sequential and rare, therefore host logic: the device reports the inverted cells of the trial movement
// (mmb_trial_move_validate / mmb_invalid_cells), this function walks them in element order exactly like the reference:
// non-interface corners are scheduled for removal; a cell whose corners are all interface vertices is resolved through
// its single interface edge (intersection with the interface element ending in the third vertex -> insertion point, third
// vertex moved back), by removing a removable interface vertex, or it is an error (GridError, same messages).
struct InterfaceInsertion {
  std::array<int32_t, 2> edge;        // ip.edge: the interface edge of the inverted cell (vertex indices, id order)
  std::array<double, 2> point;        // ip.point = lineIntersectionPoint(interface edge, crossing interface element)
  int32_t v0;                         // ip.v0 = the third vertex
  int64_t insertionLevel;             // thirdVertex.insertionLevel() + 1
  int32_t crossingSegment;            // facet index of the crossing interface element
  bool componentIsInterfaceEdge;      // ip.connectedcomponent: the longer of (interface edge, crossing element)
};
struct InterfaceMovementOutcome {
  bool change = false;
  std::vector<int32_t> removed;       // vertices appended to remove_ (in order)
  std::vector<InterfaceInsertion> inserted;
};
// xCurrent = the coordinates the mesh has NOW (the snapshot's x is the state at the last TDS edit: after moveVertices /
// moveInterface / step() the two differ); nullptr = the snapshot's own.
inline InterfaceMovementOutcome resolveInterfaceMovement2d(const Snapshot<2>& s, std::vector<std::array<double, 2>> ishifts,
                                                           const std::vector<int32_t>& invalidCellsOfTrialMove,
                                                           const std::vector<int32_t>& alreadyRemoved = {},
                                                           const std::vector<std::array<double, 2>>* xCurrent = nullptr) {
  using P = std::array<double, 2>;
  InterfaceMovementOutcome out;
  const std::vector<P>& x0 = xCurrent ? *xCurrent : s.x;
  const std::size_t nv = s.x.size(), nc = s.cells.size();
  if (ishifts.size() != s.interfaceVertices.size()) throw InvalidStateException("shifts.size() != number of interface vertices");
  if (x0.size() != nv) throw InvalidStateException("current coordinates do not match the snapshot");
  std::vector<P> xt(x0.begin(), x0.end());                                       // moveInterface(shifts), mmesh.hh:933
  std::vector<int32_t> ivIndex(nv, -1);
  for (std::size_t i = 0; i < ishifts.size(); ++i) {
    const std::size_t v = (std::size_t)s.interfaceVertices[i];
    ivIndex[v] = (int32_t)i;
    xt[v] = P{ x0[v][0] + ishifts[i][0], x0[v][1] + ishifts[i][1] };
  }
  auto isIf = [&](int32_t v) { return !s.vflags.empty() && (s.vflags[(std::size_t)v] & 1); };
  auto bflag = [&](int32_t v) { return !s.vflags.empty() && (s.vflags[(std::size_t)v] & 2); };
  // incident interface elements per vertex (facet indices, facet order) and the set of interface edges
  std::vector<std::vector<int32_t>> incident(nv);
  for (std::size_t f = 0; f < s.facets.size(); ++f) { incident[(std::size_t)s.facets[f][0]].push_back((int32_t)f); incident[(std::size_t)s.facets[f][1]].push_back((int32_t)f); }
  auto interfaceFacetOf = [&](int32_t a, int32_t b) -> int32_t {                 // isInterface(edge), mmesh.hh:607-621
    if (!isIf(a) || !isIf(b)) return -1;
    for (int32_t f : incident[(std::size_t)a]) { const auto& q = s.facets[(std::size_t)f]; if ((q[0] == a && q[1] == b) || (q[0] == b && q[1] == a)) return f; }
    return -1;
  };
  std::vector<uint8_t> removedSet(nv, 0);                                        // removed_ (mmesh.hh:1806) as a bitmap
  for (int32_t r : alreadyRemoved) removedSet[(std::size_t)r] = 1;
  auto removeOnce = [&](int32_t v) {                                             // removed_.insert(id).second
    if (removedSet[(std::size_t)v]) return;
    removedSet[(std::size_t)v] = 1; out.removed.push_back(v); out.change = true;
  };
  auto corner1 = [](const P& c0, const P& c1) { return P{ c0[0] + 1.0 * (c1[0] - c0[0]), c0[1] + 1.0 * (c1[1] - c0[1]) }; };   // AffineGeometry::corner(1)
  auto length = [](const P& a, const P& b) { const double jx = b[0] - a[0], jy = b[1] - a[1]; double q = 0.0; q += jx * jx; q += jy * jy; return std::sqrt(q); };
  auto invalidNow = [&](std::size_t c) {                                         // signedVolume_ <= 0, TDS order (mmesh.hh:1169-1173)
    const P &p = xt[(std::size_t)s.cells[c][0]], &q = xt[(std::size_t)s.cells[c][1]], &r = xt[(std::size_t)s.cells[c][2]];
    const double area = ((q[0] - p[0]) * (r[1] - p[1]) - (r[0] - p[0]) * (q[1] - p[1])) / 2;
    return area <= 0.0;
  };
  std::vector<uint8_t> movedBack(nv, 0), pending(nc, 0);
  std::vector<int32_t> order(invalidCellsOfTrialMove);
  for (int32_t c : order) pending[(std::size_t)c] = 1;
  std::vector<std::vector<int32_t>> cellsOf;                                     // built on the first move-back only
  std::size_t pos = 0;
  std::sort(order.begin(), order.end());
  while (pos < order.size()) {
    const std::size_t c = (std::size_t)order[pos++];
    const auto& tds = s.cells[c];
    if ((movedBack[(std::size_t)tds[0]] | movedBack[(std::size_t)tds[1]] | movedBack[(std::size_t)tds[2]]) && !invalidNow(c)) continue;
    int32_t k[3];
    for (int j = 0; j < 3; ++j) k[j] = tds[(s.perm[c] >> (2 * j)) & 3];          // subEntity order = id order
    bool allInterface = true;
    for (int j = 0; j < 3; ++j) {                                                // mmesh.hh:945-962
      if (isIf(k[j])) continue;
      if (bflag(k[j])) continue;                                                 // boundaryFlag(v) == 1
      removeOnce(k[j]);
      allInterface = false;
    }
    if (!allInterface) continue;
    static const int EV[3][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 } };
    int ie = -1; int32_t ifacet = -1;
    for (int j = 0; j < 3; ++j) {                                                // mmesh.hh:966-980
      const int32_t f = interfaceFacetOf(k[EV[j][0]], k[EV[j][1]]);
      if (f < 0) continue;
      if (ie >= 0)
        throw GridError("Interface could not be moved because a constrained cell with more than one interface edge would have negative volume!");
      ie = j; ifacet = f;
    }
    if (ie < 0) {                                                                // mmesh.hh:983-1006
      bool allNonRemovable = true;
      for (int j = 0; j < 3; ++j)
        if (incident[(std::size_t)k[j]].size() == 2) { allNonRemovable = false; removeOnce(k[j]); }   // isRemoveable
      if (allNonRemovable)
        throw GridError("Interface could not be moved because a constrained cell without any interface edge would have negative volume!");
      continue;
    }
    const int32_t ea = k[EV[ie][0]], eb = k[EV[ie][1]];
    const int32_t third = k[3 - EV[ie][0] - EV[ie][1]];
    int32_t crossing = -1, adj = -1;
    for (int32_t f : incident[(std::size_t)third]) {                             // mmesh.hh:1021-1036
      if (crossing >= 0) throw GridError("Interface could not be moved because two interfaces cross!");
      crossing = f;
      adj = s.facets[(std::size_t)f][0] == third ? s.facets[(std::size_t)f][1] : s.facets[(std::size_t)f][0];
    }
    (void)adj; (void)ifacet;
    if (crossing < 0) throw GridError("Interface could not be moved because two interfaces cross!");   // unreachable: third is an interface vertex
    const P c0 = xt[(std::size_t)ea], c1 = corner1(c0, xt[(std::size_t)eb]);     // iegeo.corner(0/1)
    const P ce0 = xt[(std::size_t)s.facets[(std::size_t)crossing][0]], ce1 = corner1(ce0, xt[(std::size_t)s.facets[(std::size_t)crossing][1]]);
    // PolygonCutting::lineIntersectionPoint(c0, c1, ce0, ce1), polygoncutting.hh:100-114
    P dP{ c0[0] - c1[0], c0[1] - c1[1] }, dQ{ ce0[0] - ce1[0], ce0[1] - ce1[1] };
    const double scale = dP[0] * dQ[1] - dQ[0] * dP[1];
    const double fq = ce0[0] * ce1[1] - ce0[1] * ce1[0], fp = c0[0] * c1[1] - c0[1] * c1[0];
    dP[0] *= fq; dP[1] *= fq; dQ[0] *= fp; dQ[1] *= fp;
    const P x{ (dQ[0] - dP[0]) / scale, (dQ[1] - dP[1]) / scale };
    double dot = 0.0; dot += (c0[0] - x[0]) * (c1[0] - x[0]); dot += (c0[1] - x[1]) * (c1[1] - x[1]);
    if (dot > 0.) continue;                                                      // intersection not inside the interface edge
    InterfaceInsertion ip;
    ip.edge = { ea, eb }; ip.point = x; ip.v0 = third;
    ip.insertionLevel = (s.insertionLevel.empty() ? 0 : (int64_t)s.insertionLevel[(std::size_t)third]) + 1;
    ip.crossingSegment = crossing;
    ip.componentIsInterfaceEdge = length(c0, c1) > length(ce0, ce1);             // volume() of the two 1-d geometries
    out.inserted.push_back(ip); out.change = true;
    // move the third vertex back (mmesh.hh:1072-1076) and re-examine the later cells around it
    const std::size_t idx = (std::size_t)ivIndex[(std::size_t)third];
    xt[(std::size_t)third] = P{ xt[(std::size_t)third][0] - ishifts[idx][0], xt[(std::size_t)third][1] - ishifts[idx][1] };
    ishifts[idx] = P{ 0.0, 0.0 };
    movedBack[(std::size_t)third] = 1;
    if (cellsOf.empty()) { cellsOf.resize(nv); for (std::size_t cc = 0; cc < nc; ++cc) for (int j = 0; j < 3; ++j) cellsOf[(std::size_t)s.cells[cc][j]].push_back((int32_t)cc); }
    bool added = false;
    for (int32_t cc : cellsOf[(std::size_t)third])
      if ((std::size_t)cc > c && !pending[(std::size_t)cc]) { pending[(std::size_t)cc] = 1; order.push_back(cc); added = true; }
    if (added) std::sort(order.begin() + (std::ptrdiff_t)pos, order.end());
  }
  return out;
}

template <int dim> class MMesh;

//! mirrors Dune::Distance<Grid> (distance.hh:42-129)
template <int dim>
class Distance {
 public:
  explicit Distance(MMesh<dim>& g) : g_(&g) {}
  void update();                                             // distance.hh:42-59
  bool initialized() const { return !d_.empty(); }
  double operator[](std::size_t i) const { return d_.at(i); }   // distance.hh:110-114
  double maximum() const { return max_; }                    // distance.hh:118-123
  std::size_t size() const { return d_.size(); }
  void set(const std::vector<double>& d);                    // distance.hh:84-90 for all vertices
 private:
  MMesh<dim>* g_; std::vector<double> d_; double max_ = 0.0;
};

//! mirrors Dune::RatioIndicator<Grid> (ratioindicator.hh:62-182)
template <int dim>
class RatioIndicator {
 public:
  explicit RatioIndicator(MMesh<dim>& g) : g_(&g), distance_(g) {}
  void init();                                               // ratioindicator.hh:62-91
  void update() { distance_.update(); }                      // ratioindicator.hh:94-97 (maxDist_ lives on the device)
  double& maxH() { sync_(); return p_.maxH; }
  double& minH() { sync_(); return p_.minH; }
  double& factor() { sync_(); return p_.factor; }
  double& distProportion() { sync_(); return p_.distProportion; }
  double edgeRatio() const { return 4.0; }
  double radiusRatio() const { return 30.0; }
  const Distance<dim>& distance() { if (!distance_.initialized()) distance_.update(); return distance_; }
  void push();                                               // write the (possibly modified) members to the device
 private:
  void sync_();
  MMesh<dim>* g_; Distance<dim> distance_; mmb_params p_{}; bool have_ = false;
  friend class MMesh<dim>;
};

//! mirrors the adaptation API of Dune::MMesh (mmesh.hh:725-1474) for the numeric part of the path
template <int dim>
class MMesh {
 public:
  using Coordinate = GlobalCoordinate<dim>;

  explicit MMesh(int device = 0, void* stream = nullptr) : indicator_(*this) {
    if (mmb_create(device, dim, stream, &ctx_) != MMB_OK)
      throw InvalidStateException("mmesh_b200: no usable sm_100a device (there is no CPU fallback)");
  }
  ~MMesh() { if (ctx_) mmb_destroy(ctx_); }
  MMesh(const MMesh&) = delete;
  MMesh& operator=(const MMesh&) = delete;

  //! called by the adapter after every CGAL TDS edit (adapt_(), mmesh.hh:1180-1363)
  void setSnapshot(const Snapshot<dim>& s) {
    snap_ = &s;
    check(mmb_upload_mesh(ctx_, (int64_t)s.x.size(), s.x.empty() ? nullptr : s.x[0].data(), (int64_t)s.cells.size(),
                          s.cells.empty() ? nullptr : s.cells[0].data(), s.perm.data(), (int64_t)s.edges.size(),
                          s.edges.empty() ? nullptr : s.edges[0].data(), (int64_t)s.facets.size(),
                          s.facets.empty() ? nullptr : s.facets[0].data()));
    if (dim == 2 && s.neighbours.size() == s.cells.size() && !s.cells.empty())
      check(mmb_upload_neighbours(ctx_, s.neighbours[0].data()));
    if (dim == 2 && s.insertionLevel.size() == s.x.size() && !s.x.empty())
      check(mmb_upload_vertex_levels(ctx_, s.insertionLevel.data()));
    marks_.assign(s.cells.size(), 0);
    refineMarked_ = coarsenMarked_ = 0; interfaceMarked_ = false;
    removed_.clear(); removedMask_.assign(s.x.size(), 0);
  }

  //! Delaunay restoration of the resident snapshot by edge flips on the device (the flips CGAL performs one
  //! by one in restore_Delaunay / propagating_flip, CGAL/Delaunay_triangulation_2.h:936-1022).  `s` must be the snapshot
  //! handed to setSnapshot (with neighbours); its cells, perm, neighbours and edges are replaced by the flipped topology, so
  //! the adapter can replay it into its TDS.  Interface edges are constraints.  Throws GridError if a cell is inverted.
  mmb_flip_result restoreDelaunay(Snapshot<dim>& s, int maxRounds = 1000) {
    static_assert(dim == 2, "edge flips are the 2-D algorithm");
    if (&s != snap_) throw InvalidStateException("restoreDelaunay: pass the snapshot given to setSnapshot");
    if (s.neighbours.size() != s.cells.size()) throw InvalidStateException("restoreDelaunay needs Snapshot::neighbours");
    mmb_flip_result r{};
    const int rc = mmb_restore_delaunay(ctx_, s.vertexId.size() == s.x.size() && !s.x.empty() ? s.vertexId.data() : nullptr, maxRounds, &r);
    if (rc == MMB_ERR_STATE) throw GridError(mmb_last_error(ctx_));
    check(rc);
    if (r.n_flips > 0 && !s.cells.empty())
      check(mmb_download_topology(ctx_, s.cells[0].data(), s.perm.data(), s.neighbours[0].data(), s.edges.empty() ? nullptr : s.edges[0].data()));
    marks_.assign(s.cells.size(), 0);
    refineMarked_ = coarsenMarked_ = 0; interfaceMarked_ = false;
    return r;
  }

  //! The flips of the last restoreDelaunay as (cell, slot) = the arguments of Triangulation_data_structure_2::flip(f, i), round after
  //! round (roundOffsets delimits the rounds; flips of one round touch disjoint cells).  An adapter that prefers replaying to copying
  //! the tables calls tds().flip(faceOfIndex[cell], slot) for every entry, in this order.
  struct FlipLog { std::vector<std::pair<int32_t, int>> flips; std::vector<int64_t> roundOffsets; int64_t total = 0; };
  FlipLog lastFlips() const {
    FlipLog log;
    int32_t rounds = 0;
    check(mmb_flip_log(ctx_, nullptr, 0, &log.total, nullptr, 0, &rounds));
    std::vector<uint32_t> ids((std::size_t)std::max<int64_t>(log.total, 1));
    log.roundOffsets.assign((std::size_t)rounds + 1, 0);
    check(mmb_flip_log(ctx_, ids.data(), log.total, &log.total, log.roundOffsets.data(), rounds + 1, &rounds));
    log.flips.reserve((std::size_t)log.total);
    for (int64_t i = 0; i < log.total; ++i) log.flips.push_back({ (int32_t)(ids[(std::size_t)i] / 3u), (int)(ids[(std::size_t)i] % 3u) });
    return log;
  }

  //! mmesh.hh:1421-1430
  void moveVertices(const std::vector<Coordinate>& shifts) {
    requireSize(shifts);
    check(mmb_move_vertices(ctx_, shifts[0].data()));
  }
  //! mmesh.hh:1094-1134.  Returns true when vertices were scheduled for removal; their indices are in removed().
  bool ensureVertexMovement(std::vector<Coordinate> shifts) {
    requireSize(shifts);
    valid_.resize(snap_->cells.size());
    int64_t ninv = 0;
    const int rc = mmb_trial_move_validate(ctx_, shifts[0].data(), valid_.data(), nullptr, &ninv);
    if (rc == MMB_ERR_INVALID_CELL_3D)
      throw GridError("Vertices could not be moved, because the removal of vertices is not supported in 3D!");
    check(rc);
    bool change = false;
    if (ninv > 0) {                                  // rare: host walks only the invalid cells (mmesh.hh:1109-1124)
      std::vector<int32_t> bad((std::size_t)ninv);
      check(mmb_invalid_cells(ctx_, bad.data(), ninv, &ninv));
      for (int32_t c : bad)
        for (int k = 0; k <= dim; ++k) {
          const int32_t v = snap_->cells[(std::size_t)c][(snap_->perm[(std::size_t)c] >> (2 * k)) & 3];   // subEntity order = id order
          const uint8_t fl = snap_->vflags.empty() ? 0 : snap_->vflags[(std::size_t)v];
          if (fl & 1) continue;                      // interface vertex: handled by ensureInterfaceMovement
          if (fl & 2) continue;                      // boundaryFlag(v) == 1
          if (insertRemoved(v)) change = true;       // removed_.insert(id).second (mmesh.hh:1119)
        }
    }
    return change;
  }
  const std::vector<int32_t>& removed() const { return removed_; }

  //! mmesh.hh:1407-1416: shifts are indexed by the interface grid's leaf vertex index
  void moveInterface(const std::vector<Coordinate>& ishifts) {
    scatterInterfaceShifts(ishifts);
    check(mmb_move_vertices(ctx_, full_[0].data()));
  }
  //! Numeric part of mmesh.hh:920-1088: the current mesh must be valid (:926-930, GridError otherwise); returns the cells
  //! that invert under the interface movement (the element loop over them, :935-1078, is resolveInterfaceMovement).
  std::vector<int32_t> invertedCellsUnderInterfaceMovement(const std::vector<Coordinate>& ishifts) {
    std::vector<Coordinate> zero(snap_->x.size(), Coordinate{});
    int64_t ninv = 0;
    int rc = mmb_trial_move_validate(ctx_, zero[0].data(), nullptr, nullptr, &ninv);
    if (rc == MMB_ERR_INVALID_CELL_3D || (rc == MMB_OK && ninv > 0))
      throw GridError("A cell has a negative volume! Maybe the interface has been moved too far?");   // mmesh.hh:928
    check(rc);
    scatterInterfaceShifts(ishifts);
    rc = mmb_trial_move_validate(ctx_, full_[0].data(), nullptr, nullptr, &ninv);
    if (rc == MMB_ERR_INVALID_CELL_3D)
      throw GridError("Interface could not be moved, because the removal of vertices is not supported in 3D!");   // mmesh.hh:939
    check(rc);
    std::vector<int32_t> bad((std::size_t)ninv);
    if (ninv > 0) check(mmb_invalid_cells(ctx_, bad.data(), ninv, &ninv));
    return bad;
  }

  //! The whole of ensureInterfaceMovement in 2-D (mmesh.hh:920-1088): validity of the current mesh and the inverted cells of the
  //! trial movement on the device, then the reference's element loop over those cells on the host (removals, intersection
  //! insertion points, GridError for unresolvable constrained cells).  Scheduled removals are appended to removed().
  //! The loop works on the coordinates the mesh has NOW: they are fetched from the device, because the snapshot's x is
  //! the state at the last TDS edit and moveVertices / moveInterface / step() have moved the resident mesh since.
  InterfaceMovementOutcome resolveInterfaceMovement(const std::vector<Coordinate>& ishifts) {
    if constexpr (dim == 2) {
      const std::vector<int32_t> bad = invertedCellsUnderInterfaceMovement(ishifts);
      InterfaceMovementOutcome o;
      if (!bad.empty()) {
        std::vector<Coordinate> xNow(snap_->x.size());
        check(mmb_download_coords(ctx_, xNow[0].data()));
        o = resolveInterfaceMovement2d(*snap_, ishifts, bad, removed_, &xNow);
      }
      for (int32_t v : o.removed) insertRemoved(v);
      lastInterfaceOutcome_ = o;
      return o;
    } else {
      invertedCellsUnderInterfaceMovement(ishifts);      // throws GridError on an inverted cell (mmesh.hh:938-942)
      lastInterfaceOutcome_ = InterfaceMovementOutcome{};
      return lastInterfaceOutcome_;
    }
  }
  //! mmesh.hh:920-1088 with the reference's signature: true when vertices were scheduled for removal or insertion
  //! points were created (the details are in interfaceMovementOutcome(), the removals also in removed()).
  bool ensureInterfaceMovement(std::vector<Coordinate> ishifts) { return resolveInterfaceMovement(ishifts).change; }
  const InterfaceMovementOutcome& interfaceMovementOutcome() const { return lastInterfaceOutcome_; }

  //! mmesh.hh:736-751 (bulk loop on the device, interface loop on the device, counters like :725-731)
  bool markElements() {
    indicator_.push();
    int64_t counts[2] = { 0, 0 };
    marks_.resize(snap_->cells.size());
    longest_.resize(snap_->cells.size());
    check(mmb_mark_elements(ctx_, 1, marks_.data(), longest_.data(), counts));
    refineMarked_ = counts[0]; coarsenMarked_ = counts[1];
    bool change = counts[0] + counts[1] > 0;
    if (dim == 2 && !snap_->facets.empty()) {
      imarks_.resize(snap_->facets.size());
      check(mmb_mark_interface(ctx_, imarks_.data()));
      interfaceMarked_ = false;
      for (int8_t m : imarks_) interfaceMarked_ |= (m != 0);
      change |= interfaceMarked_;
    }
    return change;
  }
  int getMark(std::size_t cellIndex) const { return marks_.at(cellIndex); }          // mmesh.hh:757-759
  bool mark(int refCount, std::size_t cellIndex) {                                    // mmesh.hh:725-731
    marks_.at(cellIndex) = (int8_t)refCount;
    if (refCount > 0) ++refineMarked_;
    if (refCount < 0) ++coarsenMarked_;
    return true;
  }
  int longestEdge(std::size_t cellIndex) const { return longest_.at(cellIndex); }    // longestedgerefinement.hh:41-57
  //! the two candidate loops of adapt() (mmesh.hh:826-908) after markElements(); needs Snapshot::neighbours (2-D).
  //! Vertices already in removed() (from ensureVertexMovement) are filtered here like removed_.insert(...).second.
  AdaptCandidates adaptCandidates() {
    static_assert(dim == 2 || dim == 3, "");
    AdaptCandidates a;
    if (dim != 2) throw InvalidStateException("adaptCandidates: 2-D only");
    int64_t n = 0;
    check(mmb_refinement_candidates(ctx_, nullptr, 0, &n));
    a.insertCells.resize((std::size_t)n); a.insertPoints.resize((std::size_t)n); a.insertEdgeVertices.resize((std::size_t)n);
    if (n > 0) {
      check(mmb_refinement_candidates(ctx_, a.insertCells.data(), n, &n));
      check(mmb_refinement_points(ctx_, n, a.insertCells.data(), a.insertPoints[0].data(), a.insertEdgeVertices[0].data()));
    }
    check(mmb_coarsening_candidates(ctx_, nullptr, nullptr, 0, &n));
    std::vector<int32_t> rc((std::size_t)n), rv((std::size_t)n);
    if (n > 0) check(mmb_coarsening_candidates(ctx_, rc.data(), rv.data(), n, &n));
    auto seen = [&](int32_t v) { return removedMask_[(std::size_t)v] != 0; };
    for (std::size_t i = 0; i < rc.size(); ++i)
      if (!seen(rv[i])) { a.removeCells.push_back(rc[i]); a.removeVertices.push_back(rv[i]); }
    if (!snap_->facets.empty()) {
      int64_t ni = 0, nr = 0;
      check(mmb_interface_candidates(ctx_, nullptr, 0, &ni, nullptr, nullptr, 0, &nr));
      a.insertSegments.resize((std::size_t)ni);
      std::vector<int32_t> rs((std::size_t)nr), rsv((std::size_t)nr);
      check(mmb_interface_candidates(ctx_, a.insertSegments.data(), ni, &ni, rs.data(), rsv.data(), nr, &nr));
      for (std::size_t i = 0; i < rs.size(); ++i)
        if (!seen(rsv[i])) { a.removeSegments.push_back(rs[i]); a.removeSegmentVertices.push_back(rsv[i]); }
    }
    return a;
  }
  bool preAdapt() const { return coarsenMarked_ > 0 || !removed_.empty(); }          // mmesh.hh:762-765 (bulk part)
  void postAdapt() {                                                                  // mmesh.hh:1456-1474
    std::fill(marks_.begin(), marks_.end(), 0);
    refineMarked_ = coarsenMarked_ = 0; interfaceMarked_ = false;
    for (int32_t v : removed_) removedMask_[(std::size_t)v] = 0;
    removed_.clear();
  }

  //! MMesh::adapt() up to the TDS edits (mmesh.hh:826-908): the insert_ / remove_ lists of both candidate loops, built on the
  //! device from the marks; true when there is something to insert or remove (the reference then runs adapt_(), :910-913,
  //! whose CGAL edits stay with the host adapter).  The lists are in candidates().
  bool adapt() {
    // no mark since the last postAdapt (every ElementInfo::mark is 0): both candidate loops of the reference find nothing
    candidates_ = (refineMarked_ > 0 || coarsenMarked_ > 0 || interfaceMarked_) ? adaptCandidates() : AdaptCandidates{};
    return !(candidates_.insertCells.empty() && candidates_.removeCells.empty() && candidates_.insertSegments.empty() &&
             candidates_.removeSegments.empty() && removed_.empty());
  }
  const AdaptCandidates& candidates() const { return candidates_; }

  //! MMesh::adapt(handle) (mmesh.hh:1138-1165) as a drop-in: preAdapt, adapt() (candidate lists on the device), the CGAL TDS
  //! edits through `edit` (the host adapter applies candidates() + removed() to its triangulation, like adapt_(), and returns
  //! the snapshot of the edited mesh -- it must outlive this object's use of it, like the one given to setSnapshot), then
  //! the data transfer: the reference calls handle.prolongLocal(father = old child, son = cut triangle) and
  //! handle.restrictLocal(father = new element, son = cut triangle, initialize) per cut triangle; here the same transfer is
  //! one batched device call, so the handle only has to move DOF blocks:
  //!     int  handle.localDofs()                          1 (FV / P0) or 3 (DG-P1, nodal on id-ordered corners)
  //!     void handle.gather(oldCellIndex, double* out)    DOFs of a cell of the OLD snapshot (a child of some component)
  //!     void handle.scatter(newCellIndex, const double*) DOFs of a cell of the NEW snapshot flagged isNew
  //! Both snapshots need Snapshot::vertexId.  Returns true like the reference.
  template <class DataHandle, class TdsEdit>
  bool adapt(DataHandle& handle, TdsEdit&& edit) {
    static_assert(dim == 2, "adapt(handle): projection is 2-D (geometryInFather, entity.hh:573-594)");
    preAdapt();
    adapt();
    const Snapshot<dim>* oldS = snap_;
    const Snapshot<dim>& newS = edit(candidates_, removed_);
    lastPairs_ = buildProjectionPairs(*oldS, oldS->vertexId, newS, newS.vertexId);
    const int ncomp = handle.localDofs();
    if (ncomp != 1 && ncomp != 3) throw InvalidStateException("adapt(handle): localDofs() must be 1 (P0) or 3 (DG-P1)");
    // gather: only the children are read (entity ids of vanished cells, persistentcontainer.hh:33-89 keeps them alive until here)
    std::vector<double> dofsOld(oldS->cells.size() * (std::size_t)ncomp, 0.0);
    std::vector<uint8_t> gathered(oldS->cells.size(), 0);
    for (int32_t o : lastPairs_.oldIndex)
      if (!gathered[(std::size_t)o]) { gathered[(std::size_t)o] = 1; handle.gather(o, &dofsOld[(std::size_t)o * (std::size_t)ncomp]); }
    setSnapshot(newS);
    if (!lastPairs_.newCell.empty()) {
      setProjectionPairs(lastPairs_.newOff, lastPairs_.newCell, lastPairs_.oldCorners, lastPairs_.oldIndex);
      std::vector<double> dofsNew;
      project(ncomp, dofsOld, dofsNew);
      for (int32_t K : lastPairs_.newCell) handle.scatter(K, &dofsNew[(std::size_t)K * (std::size_t)ncomp]);
    }
    postAdapt();
    return true;
  }
  const ProjectionPairs& lastProjectionPairs() const { return lastPairs_; }

  //! numeric part of adapt(handle) (mmesh.hh:1138-1165): pairs come from the connected components built by adapt_()
  void setProjectionPairs(const std::vector<int64_t>& newOff, const std::vector<int32_t>& newCell,
                          const std::vector<std::array<double, 6>>& oldCorners, const std::vector<int32_t>& oldIndex) {
    check(mmb_upload_pairs(ctx_, (int64_t)newCell.size(), newOff.data(), newCell.data(), (int64_t)oldIndex.size(),
                           oldCorners.empty() ? nullptr : oldCorners[0].data(), oldIndex.data()));
  }
  //! ncomp = 1 (FV / P0) or 3 (DG-P1).  dofsNew must hold ncomp entries per cell of the current snapshot.
  double project(int ncomp, const std::vector<double>& dofsOld, std::vector<double>& dofsNew) {
    double mass = 0.0;
    dofsNew.resize(snap_->cells.size() * (std::size_t)ncomp);
    check(mmb_project(ctx_, ncomp, (int64_t)dofsOld.size() / ncomp, dofsOld.data(), dofsNew.data(), &mass, nullptr, nullptr));
    return mass;
  }

  //! One whole time step of the user loop (doc/examples/moving.py:231-252) when no TDS edit intervenes, as a single
  //! pipelined call on host buffers: markElements() -> ensureVertexMovement(shifts) (numeric part) -> adapt(handle)
  //! numeric part on the uploaded pairs (ncomp = 0 skips it) -> moveVertices(shifts) -> edge legality of the moved mesh.
  //! Afterwards getMark/longestEdge/validity()/edgeFlags() hold the step's fields.
  mmb_step_result step(const std::vector<Coordinate>& shifts, int ncomp = 0, const std::vector<double>* dofsOld = nullptr,
                       std::vector<double>* dofsNew = nullptr) {
    requireSize(shifts);
    indicator_.push();
    const std::size_t nc = snap_->cells.size();
    marks_.resize(nc); longest_.resize(nc); valid_.resize(nc); orient_.resize(nc); edgeFlags_.resize(snap_->edges.size());
    if (ncomp && dofsNew) dofsNew->resize(nc * (std::size_t)ncomp);
    mmb_step_result r{};
    const int rc = mmb_adapt_step_host(ctx_, shifts[0].data(), ncomp, (ncomp && dofsOld) ? (int64_t)dofsOld->size() / ncomp : 0,
                                       (ncomp && dofsOld) ? dofsOld->data() : nullptr, (ncomp && dofsNew) ? dofsNew->data() : nullptr,
                                       marks_.data(), longest_.data(), valid_.data(), orient_.data(),
                                       edgeFlags_.empty() ? nullptr : edgeFlags_.data(), &r);
    if (rc == MMB_ERR_INVALID_CELL_3D)
      throw GridError("Vertices could not be moved, because the removal of vertices is not supported in 3D!");
    if (rc != MMB_MOVE_WITHHELD) check(rc);          // withheld: r.moved == 0, the caller removes vertices first (mmesh.hh:1109-1124)
    refineMarked_ = r.n_refine; coarsenMarked_ = r.n_coarsen;
    return r;
  }
  //! step() with move_policy 0 (default) withholds moveVertices when the trial move inverts a cell; 1 = always move
  void setMovePolicy(int policy) { indicator_.sync_(); indicator_.p_.move_policy = policy; indicator_.push(); }
  const std::vector<uint8_t>& validity() const { return valid_; }          // !(signedVolume_ <= 0) per cell (mmesh.hh:1102)
  const std::vector<int8_t>& orientation() const { return orient_; }      // exact CGAL orientation sign per cell
  const std::vector<uint8_t>& edgeFlags() const { return edgeFlags_; }    // 1 legal, 0 illegal, 2 hull edge

  //! per-edge in-circle flags (CGAL Delaunay_triangulation_2::is_valid loop, :663-681)
  std::vector<uint8_t> edgeLegality() {
    std::vector<uint8_t> f(snap_->edges.size());
    int64_t n = 0;
    check(mmb_legality(ctx_, f.data(), &n));
    return f;
  }

  RatioIndicator<dim>& indicator() { return indicator_; }                            // mmesh.hh:1773-1775
  const Distance<dim>& distance() { return indicator_.distance(); }                  // mmesh.hh:1777
  mmb_context* context() { return ctx_; }
  const Snapshot<dim>& snapshot() const { return *snap_; }

  void check(int rc) const {
    if (rc == MMB_OK) return;
    const std::string msg = mmb_last_error(ctx_);
    if (rc == MMB_ERR_INVALID_CELL_3D) throw GridError(msg);
    throw InvalidStateException("mmesh_b200: " + msg);
  }

 private:
  void scatterInterfaceShifts(const std::vector<Coordinate>& ishifts) {
    if (!snap_ || ishifts.size() != snap_->interfaceVertices.size())
      throw InvalidStateException("shifts.size() != number of interface vertices");           // assert at mmesh.hh:921,1409
    full_.assign(snap_->x.size(), Coordinate{});
    for (std::size_t i = 0; i < ishifts.size(); ++i) full_[(std::size_t)snap_->interfaceVertices[i]] = ishifts[i];
  }
  void requireSize(const std::vector<Coordinate>& s) const {                          // asserts at mmesh.hh:1095,1423
    if (!snap_ || s.size() != snap_->x.size()) throw InvalidStateException("shifts.size() != number of vertices");
  }
  mmb_context* ctx_ = nullptr;
  const Snapshot<dim>* snap_ = nullptr;
  RatioIndicator<dim> indicator_;
  std::vector<int8_t> marks_, imarks_;
  std::vector<uint8_t> longest_, valid_, edgeFlags_;
  std::vector<int8_t> orient_;
  bool insertRemoved(int32_t v) {                                                     // removed_.insert(id).second, O(1)
    if (removedMask_.size() < snap_->x.size()) removedMask_.resize(snap_->x.size(), 0);
    if (removedMask_[(std::size_t)v]) return false;
    removedMask_[(std::size_t)v] = 1; removed_.push_back(v);
    return true;
  }
  std::vector<int32_t> removed_;                     // insertion order, like the reference's remove_ vector
  std::vector<uint8_t> removedMask_;                 // removed_ (std::unordered_set<IdType>, mmesh.hh:1806) as a bitmap
  AdaptCandidates candidates_;
  InterfaceMovementOutcome lastInterfaceOutcome_;
  ProjectionPairs lastPairs_;
  std::vector<Coordinate> full_;
  int64_t refineMarked_ = 0, coarsenMarked_ = 0;
  bool interfaceMarked_ = false;
};

template <int dim>
void Distance<dim>::update() {
  d_.resize(g_->snapshot().x.size());
  g_->check(mmb_distance_update(g_->context(), d_.data(), &max_));
}
template <int dim>
void Distance<dim>::set(const std::vector<double>& d) {
  d_ = d;
  g_->check(mmb_distance_set(g_->context(), d_.data()));
  max_ = 0.0;
  for (double v : d_) max_ = v > max_ ? v : max_;
}
template <int dim>
void RatioIndicator<dim>::sync_() {
  if (!have_) { g_->check(mmb_get_params(g_->context(), &p_)); have_ = true; }
}
template <int dim>
void RatioIndicator<dim>::init() {
  g_->check(mmb_indicator_init(g_->context()));
  g_->check(mmb_get_params(g_->context(), &p_));
  have_ = true;
}
template <int dim>
void RatioIndicator<dim>::push() {
  if (have_) g_->check(mmb_set_params(g_->context(), &p_));
}

}  // namespace B200
}  // namespace Dune
