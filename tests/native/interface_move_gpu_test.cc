// Drives Dune::B200::MMesh<2>::ensureInterfaceMovement (mmesh.hh:920-1088) on the GPU path AFTER the mesh has been moved:
// validity of the current mesh and the inverted cells of the trial movement come from the device, the element loop runs on
// the coordinates the device holds now (not the snapshot's).  Input: the text snapshot of interface_move_test plus one
// vertex shift per vertex (applied with moveVertices first).  Output: like interface_move_test.
#include <cstdio>
#include <fstream>
#include "mmesh_b200.hh"
using namespace Dune::B200;
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1]);
  std::size_t nv, nc, ns, niv;
  in >> nv >> nc >> ns >> niv;
  Snapshot<2> s;
  s.x.resize(nv); s.cells.resize(nc); s.perm.resize(nc); s.facets.resize(ns); s.vflags.resize(nv); s.insertionLevel.resize(nv);
  s.interfaceVertices.resize(niv);
  for (auto& p : s.x) in >> p[0] >> p[1];
  for (auto& c : s.cells) in >> c[0] >> c[1] >> c[2];
  for (auto& p : s.perm) { int t; in >> t; p = (uint8_t)t; }
  for (auto& f : s.facets) in >> f[0] >> f[1];
  for (auto& f : s.vflags) { int t; in >> t; f = (uint8_t)t; }
  for (auto& l : s.insertionLevel) in >> l;
  for (auto& v : s.interfaceVertices) in >> v;
  std::vector<GlobalCoordinate<2>> ish(niv), vsh(nv);
  for (auto& p : ish) in >> p[0] >> p[1];
  for (auto& p : vsh) in >> p[0] >> p[1];
  try {
    MMesh<2> grid(0);
    grid.setSnapshot(s);
    grid.moveVertices(vsh);                       // the snapshot's x is stale from here on
    const bool change = grid.ensureInterfaceMovement(ish);
    const InterfaceMovementOutcome& o = grid.interfaceMovementOutcome();
    std::printf("change %d\nremoved", (int)change);
    for (int32_t v : grid.removed()) std::printf(" %d", v);
    std::printf("\ninserted %zu\n", o.inserted.size());
    for (const auto& ip : o.inserted)
      std::printf("ip %d %d %.17g %.17g %d %lld %d %d\n", ip.edge[0], ip.edge[1], ip.point[0], ip.point[1], ip.v0,
                  (long long)ip.insertionLevel, ip.crossingSegment, (int)ip.componentIsInterfaceEdge);
    // the trial movement must have left the resident coordinates untouched
    std::vector<GlobalCoordinate<2>> now(nv);
    grid.check(mmb_download_coords(grid.context(), now[0].data()));
    bool same = true;
    for (std::size_t v = 0; v < nv; ++v) same = same && now[v][0] == s.x[v][0] + vsh[v][0] && now[v][1] == s.x[v][1] + vsh[v][1];
    std::printf("coords_kept %d\n", (int)same);
  } catch (const GridError& e) {
    std::printf("GridError %s\n", e.what());
  } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
  return 0;
}
