// C wrapper around the PRODUCT host function Dune::B200::makeShard (dune-mmesh_b200/host/partition.hh) for the CPU test
// tests/test_host_partition.py, which compares it with the Python partitioner used by the multi-GPU bench driver.
#include <cstring>
#include "../../dune-mmesh_b200/host/partition.hh"
using namespace Dune::B200;
static Shard g_sh;
extern "C" {
// sizes out: nlv, nlc, nle, own_lo, own_hi, nsend, nrecv, nrows, npairs
int hpart_make(long nv, const double* x, long nc, const int32_t* cells, const uint8_t* perm, long ne, const int32_t* edges,
               const int32_t* edge_cell, long ns, const int32_t* segs, int world, int rank, long nrows, const int64_t* new_off,
               const int32_t* new_cell, const double* old_xy, const int32_t* old_idx, long* sizes) {
  Snapshot<2> g;
  g.x.resize(nv); g.cells.resize(nc); g.perm.assign(perm, perm + nc); g.edges.resize(ne); g.facets.resize(ns);
  for (long i = 0; i < nv; ++i) g.x[i] = { x[2 * i], x[2 * i + 1] };
  for (long i = 0; i < nc; ++i) g.cells[i] = { cells[3 * i], cells[3 * i + 1], cells[3 * i + 2] };
  for (long i = 0; i < ne; ++i) g.edges[i] = { edges[4 * i], edges[4 * i + 1], edges[4 * i + 2], edges[4 * i + 3] };
  for (long i = 0; i < ns; ++i) g.facets[i] = { segs[2 * i], segs[2 * i + 1] };
  std::vector<int32_t> ec(edge_cell, edge_cell + ne);
  try {
    if (new_off) {
      std::vector<int64_t> no(new_off, new_off + nrows + 1); std::vector<int32_t> ncl(new_cell, new_cell + nrows);
      const long np = (long)no.back();
      std::vector<std::array<double, 6>> oxy(np); std::memcpy(oxy.data(), old_xy, np * 48);
      std::vector<int32_t> oi(old_idx, old_idx + np);
      g_sh = makeShard(g, ec, world, rank, &no, &ncl, &oxy, &oi);
    } else
      g_sh = makeShard(g, ec, world, rank);
  } catch (const std::exception&) { return 1; }
  const long s[9] = { (long)g_sh.localVertices.size(), (long)g_sh.localCells.size(), (long)g_sh.local.edges.size(), (long)g_sh.ownLo, (long)g_sh.ownHi,
                      (long)g_sh.sendIdx.size(), (long)g_sh.recvIdx.size(), (long)g_sh.newCell.size(), (long)g_sh.oldIndex.size() };
  std::memcpy(sizes, s, sizeof s);
  return 0;
}
void hpart_get(int32_t* lverts, int32_t* lcells, double* x, int32_t* cells, uint8_t* perm, int32_t* edges, int32_t* segs, int32_t* vowner,
               int32_t* send_idx, int32_t* recv_idx, int64_t* send_counts, int64_t* recv_counts, int64_t* new_off, int32_t* new_cell,
               double* old_xy, int32_t* old_idx) {
  auto cp = [](void* d, const void* s, std::size_t n) { if (n) std::memcpy(d, s, n); };
  cp(lverts, g_sh.localVertices.data(), g_sh.localVertices.size() * 4); cp(lcells, g_sh.localCells.data(), g_sh.localCells.size() * 4);
  cp(x, g_sh.local.x.data(), g_sh.local.x.size() * 16); cp(cells, g_sh.local.cells.data(), g_sh.local.cells.size() * 12);
  cp(perm, g_sh.local.perm.data(), g_sh.local.perm.size()); cp(edges, g_sh.local.edges.data(), g_sh.local.edges.size() * 16);
  cp(segs, g_sh.local.facets.data(), g_sh.local.facets.size() * 8); cp(vowner, g_sh.vertexOwner.data(), g_sh.vertexOwner.size() * 4);
  cp(send_idx, g_sh.sendIdx.data(), g_sh.sendIdx.size() * 4); cp(recv_idx, g_sh.recvIdx.data(), g_sh.recvIdx.size() * 4);
  cp(send_counts, g_sh.sendCounts.data(), g_sh.sendCounts.size() * 8); cp(recv_counts, g_sh.recvCounts.data(), g_sh.recvCounts.size() * 8);
  cp(new_off, g_sh.newOff.data(), g_sh.newOff.size() * 8); cp(new_cell, g_sh.newCell.data(), g_sh.newCell.size() * 4);
  cp(old_xy, g_sh.oldCorners.data(), g_sh.oldCorners.size() * 48); cp(old_idx, g_sh.oldIndex.data(), g_sh.oldIndex.size() * 4);
}
}
