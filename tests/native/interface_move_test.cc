// Drives the host-side resolution of inverted cells under an interface movement (resolveInterfaceMovement2d in
// dune-mmesh_b200/host/mmesh_b200.hh, mmesh.hh:935-1078) on a snapshot read from a text file; no GPU involved: the list of
// inverted cells that the device would report is part of the input.
#include <cstdio>
#include <fstream>
#include "mmesh_b200.hh"
using namespace Dune::B200;
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1]);
  std::size_t nv, nc, ns, niv, ninv;
  in >> nv >> nc >> ns >> niv >> ninv;
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
  std::vector<std::array<double, 2>> ish(niv);
  for (auto& p : ish) in >> p[0] >> p[1];
  std::vector<int32_t> invalid(ninv);
  for (auto& c : invalid) in >> c;
  try {
    const InterfaceMovementOutcome o = resolveInterfaceMovement2d(s, ish, invalid);
    std::printf("change %d\nremoved", (int)o.change);
    for (int32_t v : o.removed) std::printf(" %d", v);
    std::printf("\ninserted %zu\n", o.inserted.size());
    for (const auto& ip : o.inserted)
      std::printf("ip %d %d %.17g %.17g %d %lld %d %d\n", ip.edge[0], ip.edge[1], ip.point[0], ip.point[1], ip.v0,
                  (long long)ip.insertionLevel, ip.crossingSegment, (int)ip.componentIsInterfaceEdge);
  } catch (const GridError& e) {
    std::printf("GridError %s\n", e.what());
  }
  return 0;
}
