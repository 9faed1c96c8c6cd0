// pairs.hh -- the (new cell, old children) list of MMesh::adapt(handle) from an old and a new snapshot.
//
// Reference: adapt_() marks the cells of every conflict zone `mightVanish` and gives overlapping zones one component number
// (dune/mmesh/grid/mmesh.hh:1194-1218,1919-1941), buildConnectedComponents_ (:1506-1537) collects the vanishing cells of a
// component as caching entities (MMeshConnectedComponent, dune/mmesh/grid/connectedcomponent.hh:52-127: seed + recursive
// insertNeighbors_ over the intersections), the cells created by the TDS edit are flagged isNew and mapped to that component
// (:1875-1884,1905-1916), and adapt(handle) walks `element.isNew()` x `connectedComponent().children()` (:1143-1161).
//
// Here the same list is derived from what a host adapter can hand over without CGAL: the snapshot before the TDS edit and the
// snapshot after it, with persistent vertex ids (VertexInfo::id, cgal/defaults.hh:41).
//   vanished cell  = old cell whose id triple does not occur in the new snapshot          (the reference's mightVanish cells
//                    that really vanished; cells of a conflict zone that survive the edit unchanged keep their DOFs)
//   isNew cell     = new cell whose id triple does not occur in the old snapshot
//   component      = vanished cells connected through shared faces -- or the caller's componentNumber per old cell, when the
//                    adapter exports ElementInfo::componentNumber (cgal/defaults.hh:33)
//   new -> component through a shared DIRECTED edge (u -> v) with a vanished cell: both lie to the left of that edge, i.e. on
//                    the same side, so they belong to the same hole; new cells without such an edge (interior of a larger
//                    hole) inherit it from the new cells they touch
//   children order = connectedcomponent.hh:52-127: lowest vanished cell first, then depth first over the faces in the
//                    order of the intersection iterator (host neighbour dim - i, intersectioniterator.hh:60-62)
//   old corners    = MMeshCachingEntity::vertex_ = geo.corner(i), i.e. id-sorted corner order (cachingentity.hh:83-89)
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "snapshot.hh"

namespace Dune {
namespace B200 {

struct ProjectionPairs {
  std::vector<int64_t> newOff;                         // CSR over the isNew cells (ascending cell index)
  std::vector<int32_t> newCell;                        // index of the isNew cell in the new snapshot
  std::vector<std::array<double, 6>> oldCorners;       // CachingEntity::vertex_ of every child
  std::vector<int32_t> oldIndex;                       // index of the child in the old snapshot (DOF lookup)
  std::vector<int32_t> componentOfOld, componentOfNew; // dense component id, -1 = unchanged cell
  std::size_t components = 0;
};

namespace detail {
struct Key3 { int64_t a, b, c; bool operator==(const Key3& o) const { return a == o.a && b == o.b && c == o.c; } };
struct Key2 { int64_t a, b; bool operator==(const Key2& o) const { return a == o.a && b == o.b; } };
inline uint64_t mix(uint64_t z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
struct Hash3 { std::size_t operator()(const Key3& k) const { return (std::size_t)mix(mix(mix((uint64_t)k.a) ^ (uint64_t)k.b) ^ (uint64_t)k.c); } };
struct Hash2 { std::size_t operator()(const Key2& k) const { return (std::size_t)mix(mix((uint64_t)k.a) ^ (uint64_t)k.b); } };
struct UnionFind {
  std::vector<int32_t> p;
  explicit UnionFind(std::size_t n) : p(n) { for (std::size_t i = 0; i < n; ++i) p[i] = (int32_t)i; }
  int32_t find(int32_t x) { while (p[(std::size_t)x] != x) { p[(std::size_t)x] = p[(std::size_t)p[(std::size_t)x]]; x = p[(std::size_t)x]; } return x; }
  void unite(int32_t a, int32_t b) { a = find(a); b = find(b); if (a != b) p[(std::size_t)std::max(a, b)] = std::min(a, b); }   // root = lowest index
};
}  // namespace detail

//! vertexId: persistent id per vertex index of the snapshot (empty = the index itself)
inline ProjectionPairs buildProjectionPairs(const Snapshot<2>& oldS, const std::vector<int64_t>& oldVertexId,
                                            const Snapshot<2>& newS, const std::vector<int64_t>& newVertexId,
                                            const std::vector<int32_t>* componentNumberOld = nullptr) {
  using namespace detail;
  const std::size_t nco = oldS.cells.size(), ncn = newS.cells.size();
  auto oid = [&](int32_t v) { return oldVertexId.empty() ? (int64_t)v : oldVertexId[(std::size_t)v]; };
  auto nid = [&](int32_t v) { return newVertexId.empty() ? (int64_t)v : newVertexId[(std::size_t)v]; };
  auto key3 = [](int64_t a, int64_t b, int64_t c) { if (a > b) std::swap(a, b); if (b > c) std::swap(b, c); if (a > b) std::swap(a, b); return Key3{ a, b, c }; };
  if (componentNumberOld && componentNumberOld->size() != nco) throw InvalidStateException("componentNumber.size() != number of old cells");

  ProjectionPairs out;
  out.componentOfOld.assign(nco, -1); out.componentOfNew.assign(ncn, -1);
  // 1. unchanged cells: same id triple on both sides
  std::unordered_map<Key3, int32_t, Hash3> oldByKey; oldByKey.reserve(nco * 2);
  for (std::size_t c = 0; c < nco; ++c) oldByKey.emplace(key3(oid(oldS.cells[c][0]), oid(oldS.cells[c][1]), oid(oldS.cells[c][2])), (int32_t)c);
  std::vector<uint8_t> vanished(nco, 1), isNew(ncn, 0);
  for (std::size_t c = 0; c < ncn; ++c) {
    auto it = oldByKey.find(key3(nid(newS.cells[c][0]), nid(newS.cells[c][1]), nid(newS.cells[c][2])));
    if (it == oldByKey.end()) isNew[c] = 1; else vanished[(std::size_t)it->second] = 0;
  }
  // 2. directed edges (u -> v), TDS order is counter-clockwise: the cell lies to the left of each of its three edges
  std::unordered_map<Key2, int32_t, Hash2> vanEdge, newEdge;
  for (std::size_t c = 0; c < nco; ++c) if (vanished[c])
    for (int k = 0; k < 3; ++k) vanEdge.emplace(Key2{ oid(oldS.cells[c][k]), oid(oldS.cells[c][(k + 1) % 3]) }, (int32_t)c);
  for (std::size_t c = 0; c < ncn; ++c) if (isNew[c])
    for (int k = 0; k < 3; ++k) newEdge.emplace(Key2{ nid(newS.cells[c][k]), nid(newS.cells[c][(k + 1) % 3]) }, (int32_t)c);
  // face neighbour k of a vanished cell (opposite TDS vertex k) if it vanished too, else -1
  auto vanNeighbour = [&](std::size_t c, int k) -> int32_t {
    const int64_t u = oid(oldS.cells[c][(k + 1) % 3]), v = oid(oldS.cells[c][(k + 2) % 3]);
    auto it = vanEdge.find(Key2{ v, u });
    return it == vanEdge.end() ? -1 : it->second;
  };
  // 3. components of the vanished cells
  UnionFind uf(nco);
  if (componentNumberOld) {
    std::unordered_map<int32_t, int32_t> first;
    for (std::size_t c = 0; c < nco; ++c) if (vanished[c]) {
      const int32_t num = (*componentNumberOld)[c];
      if (num <= 0) throw InvalidStateException("a vanished cell carries no component number");
      auto it = first.emplace(num, (int32_t)c);
      if (!it.second) uf.unite(it.first->second, (int32_t)c);
    }
  } else {
    for (std::size_t c = 0; c < nco; ++c) if (vanished[c])
      for (int k = 0; k < 3; ++k) { const int32_t n = vanNeighbour(c, k); if (n >= 0) uf.unite((int32_t)c, n); }
  }
  std::vector<int32_t> compOfRoot(nco, -1);
  for (std::size_t c = 0; c < nco; ++c) if (vanished[c]) {
    const int32_t r = uf.find((int32_t)c);
    if (compOfRoot[(std::size_t)r] < 0) compOfRoot[(std::size_t)r] = (int32_t)out.components++;
    out.componentOfOld[c] = compOfRoot[(std::size_t)r];
  }
  // 4. children lists in the order of MMeshConnectedComponent: seed, then insertNeighbors_ depth first, faces dim .. 0
  std::vector<std::vector<int32_t>> children(out.components);
  {
    std::vector<uint8_t> placed(nco, 0);
    std::vector<std::pair<int32_t, int>> stack;
    for (std::size_t seed = 0; seed < nco; ++seed) {
      if (!vanished[seed] || placed[seed]) continue;
      const int32_t comp = out.componentOfOld[seed];
      placed[seed] = 1; children[(std::size_t)comp].push_back((int32_t)seed);
      stack.assign(1, { (int32_t)seed, 2 });
      while (!stack.empty()) {
        auto& top = stack.back();
        if (top.second < 0) { stack.pop_back(); continue; }
        const int k = top.second--;
        const int32_t n = vanNeighbour((std::size_t)top.first, k);
        if (n >= 0 && !placed[(std::size_t)n] && out.componentOfOld[(std::size_t)n] == comp) {
          placed[(std::size_t)n] = 1; children[(std::size_t)comp].push_back(n);
          stack.push_back({ n, 2 });
        }
      }
    }
  }
  // 5. new cells -> component
  UnionFind ufn(ncn);
  for (std::size_t c = 0; c < ncn; ++c) if (isNew[c])
    for (int k = 0; k < 3; ++k) {
      const int64_t u = nid(newS.cells[c][k]), v = nid(newS.cells[c][(k + 1) % 3]);
      auto it = vanEdge.find(Key2{ u, v });
      if (it != vanEdge.end()) { out.componentOfNew[c] = out.componentOfOld[(std::size_t)it->second]; continue; }
      auto jt = newEdge.find(Key2{ v, u });                 // an edge inside the hole: the cell across is new as well
      if (jt != newEdge.end()) ufn.unite((int32_t)c, jt->second);
    }
  {
    std::vector<int32_t> compOfSet(ncn, -1);
    for (std::size_t c = 0; c < ncn; ++c) if (isNew[c] && out.componentOfNew[c] >= 0) compOfSet[(std::size_t)ufn.find((int32_t)c)] = out.componentOfNew[c];
    for (std::size_t c = 0; c < ncn; ++c) if (isNew[c] && out.componentOfNew[c] < 0) {
      out.componentOfNew[c] = compOfSet[(std::size_t)ufn.find((int32_t)c)];
      if (out.componentOfNew[c] < 0) throw InvalidStateException("a new cell touches no vanished cell: the snapshots do not describe one adapt step");
    }
  }
  // 6. the CSR, new cells in element order (mmesh.hh:1143), children in component order
  out.newOff.push_back(0);
  for (std::size_t c = 0; c < ncn; ++c) {
    if (!isNew[c]) continue;
    out.newCell.push_back((int32_t)c);
    for (int32_t o : children[(std::size_t)out.componentOfNew[c]]) {
      const auto& cell = oldS.cells[(std::size_t)o];
      const unsigned pm = oldS.perm[(std::size_t)o];
      std::array<double, 6> xy;
      for (int k = 0; k < 3; ++k) {                         // id-sorted corners = geo.corner(k) of the old entity
        const auto& p = oldS.x[(std::size_t)cell[(pm >> (2 * k)) & 3]];
        xy[2 * (std::size_t)k] = p[0]; xy[2 * (std::size_t)k + 1] = p[1];
      }
      out.oldCorners.push_back(xy); out.oldIndex.push_back(o);
    }
    out.newOff.push_back((int64_t)out.oldIndex.size());
  }
  return out;
}

}  // namespace B200
}  // namespace Dune
