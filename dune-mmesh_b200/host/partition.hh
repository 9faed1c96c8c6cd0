// partition.hh -- Hilbert-range sharding of a 2-D mesh snapshot with a one-ring halo (host logic, no device calls).
//
// Reference: dune/mmesh/misc/partitionhelper.hh:175-236 gives every rank the WHOLE triangulation, marks contiguous ranges of
// the iteration order as interior (:175-200) and adds one face-neighbour ghost layer (:219-236); the exchange is raw MPI
// point-to-point (dune/mmesh/misc/communication.hh:83-127).  Here a rank stores only
//   * its contiguous range of the (Hilbert-sorted) cell order                        -> owned cells,
//   * every cell sharing a vertex with an owned cell                                 -> one-ring halo cells (a superset of
//                                                                                       the reference's ghost layer),
//   * all vertices of those cells plus the (small, replicated) interface vertices    -> local vertices.
// A vertex is owned by the rank owning its lowest-index incident cell.  Local numbering preserves the global order, so
// the Hilbert locality the kernels rely on survives.  Per step only the shifts of non-owned local vertices (halo
// exchange) and a few scalars cross NVLink (mmb_adapt_step_sharded, include/mmesh_b200.h).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <vector>

#include "snapshot.hh"

namespace Dune {
namespace B200 {

struct Shard {
  int rank = 0, world = 1;
  int64_t c0 = 0, c1 = 0;                       // owned global cell range
  std::vector<int32_t> localCells, localVertices;   // global ids, ascending
  Snapshot<2> local;                            // x, cells, perm, edges, facets, vflags in local numbering
  int64_t ownLo = 0, ownHi = 0;                 // owned range inside localCells (mmb_set_owned_cells)
  std::vector<int32_t> vertexOwner;             // owner rank per local vertex
  std::vector<int32_t> sendIdx, recvIdx;        // local vertex ids: per peer concatenated (mmb_halo_setup)
  std::vector<int64_t> sendCounts, recvCounts;  // per peer (mmb_halo_counts)
  // projection pairs of the owned new cells (local numbering; old indices refer to local cells)
  std::vector<int64_t> newOff; std::vector<int32_t> newCell, oldIndex; std::vector<std::array<double, 6>> oldCorners;
};

inline std::vector<int64_t> cellBounds(int64_t nc, int world) {
  std::vector<int64_t> b((std::size_t)world + 1);
  for (int r = 0; r <= world; ++r) b[(std::size_t)r] = ((int64_t)r * nc) / world;      // partitionhelper.hh:175-200: equal contiguous ranges
  return b;
}

namespace detail {
// cells sharing a vertex with [c0, c1) and the vertices of those cells (+ the interface's)
inline void localSets(const Snapshot<2>& g, int64_t c0, int64_t c1, std::vector<uint8_t>& vmask, std::vector<int32_t>& lcells,
                      std::vector<int32_t>& lverts) {
  const std::size_t nv = g.x.size(), nc = g.cells.size();
  vmask.assign(nv, 0);
  for (int64_t c = c0; c < c1; ++c) for (int k = 0; k < 3; ++k) vmask[(std::size_t)g.cells[(std::size_t)c][k]] = 1;
  lcells.clear();
  for (std::size_t c = 0; c < nc; ++c)
    if (vmask[(std::size_t)g.cells[c][0]] | vmask[(std::size_t)g.cells[c][1]] | vmask[(std::size_t)g.cells[c][2]]) lcells.push_back((int32_t)c);
  std::vector<uint8_t> lv(nv, 0);
  for (int32_t c : lcells) for (int k = 0; k < 3; ++k) lv[(std::size_t)g.cells[(std::size_t)c][k]] = 1;
  for (const auto& f : g.facets) { lv[(std::size_t)f[0]] = 1; lv[(std::size_t)f[1]] = 1; }              // interface is replicated
  lverts.clear();
  for (std::size_t v = 0; v < nv; ++v) if (lv[v]) lverts.push_back((int32_t)v);
}
}  // namespace detail

//! edgeLeftCell[e] = index of the cell to the left of edge e = (a, b) (the one holding apex c); the edge list must be ordered
//! by that cell (the layout the kernels' Hilbert tiles expect), so the owned edges are a contiguous range.
//! Pair CSR (optional, nullptr = none): rows belong to the new cells newCell[i] (ascending); every row of an owned cell goes
//! to this rank; its old children must be local cells (they are: children share a vertex with the new cell).
inline Shard makeShard(const Snapshot<2>& g, const std::vector<int32_t>& edgeLeftCell, int world, int rank,
                       const std::vector<int64_t>* newOff = nullptr, const std::vector<int32_t>* newCell = nullptr,
                       const std::vector<std::array<double, 6>>* oldCorners = nullptr, const std::vector<int32_t>* oldIndex = nullptr) {
  if (world < 1 || rank < 0 || rank >= world) throw InvalidStateException("makeShard: bad rank / world");
  if (edgeLeftCell.size() != g.edges.size()) throw InvalidStateException("makeShard: edgeLeftCell.size() != number of edges");
  const std::size_t nv = g.x.size(), nc = g.cells.size();
  const std::vector<int64_t> bounds = cellBounds((int64_t)nc, world);
  Shard sh;
  sh.rank = rank; sh.world = world; sh.c0 = bounds[(std::size_t)rank]; sh.c1 = bounds[(std::size_t)rank + 1];
  // vertex owner = rank of the lowest-index incident cell
  std::vector<int64_t> ownerCell(nv, (int64_t)nc);
  for (std::size_t c = nc; c-- > 0;) for (int k = 0; k < 3; ++k) ownerCell[(std::size_t)g.cells[c][k]] = (int64_t)c;
  auto rankOfCell = [&](int64_t c) { return (int32_t)(std::upper_bound(bounds.begin(), bounds.end(), c) - bounds.begin() - 1); };
  std::vector<int32_t> vowner(nv);
  for (std::size_t v = 0; v < nv; ++v) vowner[v] = rankOfCell(ownerCell[v]);
  std::vector<uint8_t> vmask;
  detail::localSets(g, sh.c0, sh.c1, vmask, sh.localCells, sh.localVertices);
  std::vector<int32_t> g2lv(nv, -1), g2lc(nc, -1);
  for (std::size_t i = 0; i < sh.localVertices.size(); ++i) g2lv[(std::size_t)sh.localVertices[i]] = (int32_t)i;
  for (std::size_t i = 0; i < sh.localCells.size(); ++i) g2lc[(std::size_t)sh.localCells[i]] = (int32_t)i;
  // local snapshot
  Snapshot<2>& L = sh.local;
  L.x.reserve(sh.localVertices.size());
  for (int32_t v : sh.localVertices) { L.x.push_back(g.x[(std::size_t)v]); if (!g.vflags.empty()) L.vflags.push_back(g.vflags[(std::size_t)v]); }
  for (int32_t c : sh.localCells) {
    const auto& t = g.cells[(std::size_t)c];
    L.cells.push_back({ g2lv[(std::size_t)t[0]], g2lv[(std::size_t)t[1]], g2lv[(std::size_t)t[2]] });
    L.perm.push_back(g.perm[(std::size_t)c]);
  }
  sh.ownLo = std::lower_bound(sh.localCells.begin(), sh.localCells.end(), (int32_t)sh.c0) - sh.localCells.begin();
  sh.ownHi = std::lower_bound(sh.localCells.begin(), sh.localCells.end(), (int32_t)sh.c1) - sh.localCells.begin();
  // owned edges: left cell owned -> contiguous range of the edge array
  {
    const std::size_t e0 = std::lower_bound(edgeLeftCell.begin(), edgeLeftCell.end(), (int32_t)sh.c0) - edgeLeftCell.begin();
    const std::size_t e1 = std::lower_bound(edgeLeftCell.begin(), edgeLeftCell.end(), (int32_t)sh.c1) - edgeLeftCell.begin();
    for (std::size_t e = e0; e < e1; ++e) {
      std::array<int32_t, 4> q;
      for (int k = 0; k < 4; ++k) q[(std::size_t)k] = g.edges[e][(std::size_t)k] < 0 ? -1 : g2lv[(std::size_t)g.edges[e][(std::size_t)k]];
      L.edges.push_back(q);
    }
  }
  for (const auto& f : g.facets) L.facets.push_back({ g2lv[(std::size_t)f[0]], g2lv[(std::size_t)f[1]] });
  sh.vertexOwner.reserve(sh.localVertices.size());
  for (int32_t v : sh.localVertices) sh.vertexOwner.push_back(vowner[(std::size_t)v]);
  // halo lists.  recv: my non-owned local vertices grouped by owner rank, ascending global id inside a group
  sh.recvCounts.assign((std::size_t)world, 0); sh.sendCounts.assign((std::size_t)world, 0);
  {
    std::vector<std::vector<int32_t>> by((std::size_t)world);
    for (std::size_t i = 0; i < sh.localVertices.size(); ++i)
      if (sh.vertexOwner[i] != rank) by[(std::size_t)sh.vertexOwner[i]].push_back((int32_t)i);
    for (int p = 0; p < world; ++p) { sh.recvCounts[(std::size_t)p] = (int64_t)by[(std::size_t)p].size(); sh.recvIdx.insert(sh.recvIdx.end(), by[(std::size_t)p].begin(), by[(std::size_t)p].end()); }
  }
  // send: for every peer p, the vertices I own that p holds as non-owned, in p's receive order (ascending global id)
  {
    std::vector<int32_t> pc, pv; std::vector<uint8_t> pm;
    for (int p = 0; p < world; ++p) {
      if (p == rank) continue;
      detail::localSets(g, bounds[(std::size_t)p], bounds[(std::size_t)p + 1], pm, pc, pv);
      for (int32_t v : pv) if (vowner[(std::size_t)v] == rank) { sh.sendIdx.push_back(g2lv[(std::size_t)v]); ++sh.sendCounts[(std::size_t)p]; }
    }
  }
  // projection pairs of the owned new cells: rows selected by the cell they belong to (the CSR may list only isNew cells)
  if (newOff && newCell && oldCorners && oldIndex) {
    const std::size_t r0 = std::lower_bound(newCell->begin(), newCell->end(), (int32_t)sh.c0) - newCell->begin();
    const std::size_t r1 = std::lower_bound(newCell->begin(), newCell->end(), (int32_t)sh.c1) - newCell->begin();
    const int64_t p0 = (*newOff)[r0], p1 = (*newOff)[r1];
    for (std::size_t r = r0; r <= r1; ++r) sh.newOff.push_back((*newOff)[r] - p0);
    for (std::size_t r = r0; r < r1; ++r) sh.newCell.push_back(g2lc[(std::size_t)(*newCell)[r]]);
    for (int64_t p = p0; p < p1; ++p) {
      const int32_t o = g2lc[(std::size_t)(*oldIndex)[(std::size_t)p]];
      if (o < 0) throw InvalidStateException("makeShard: an old child of an owned new cell lies outside the one-ring halo");
      sh.oldIndex.push_back(o); sh.oldCorners.push_back((*oldCorners)[(std::size_t)p]);
    }
  }
  return sh;
}

}  // namespace B200
}  // namespace Dune
