// Drives the C++ host class (dune-mmesh_b200/host/mmesh_b200.hh) the way a user of the reference drives Dune::MMesh:
//   markElements(); ensureVertexMovement(shifts); moveVertices(shifts); postAdapt();
// Input: a text snapshot written by tests/test_gpu_host_api.py; output: one line per result array.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include "mmesh_b200.hh"

using namespace Dune::B200;

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: host_api_test snapshot.txt\n"); return 2; }
  std::ifstream in(argv[1]);
  Snapshot<2> s;
  std::size_t nv, nc, ne, ns;
  in >> nv >> nc >> ne >> ns;
  s.x.resize(nv); s.cells.resize(nc); s.perm.resize(nc); s.edges.resize(ne); s.facets.resize(ns); s.vflags.resize(nv);
  std::vector<GlobalCoordinate<2>> shifts(nv);
  for (auto& p : s.x) in >> p[0] >> p[1];
  for (auto& p : shifts) in >> p[0] >> p[1];
  for (auto& f : s.vflags) { int t; in >> t; f = (uint8_t)t; }
  for (auto& c : s.cells) in >> c[0] >> c[1] >> c[2];
  for (auto& p : s.perm) { int t; in >> t; p = (uint8_t)t; }
  for (auto& e : s.edges) in >> e[0] >> e[1] >> e[2] >> e[3];
  for (auto& f : s.facets) in >> f[0] >> f[1];
  s.neighbours.resize(nc);
  for (auto& n : s.neighbours) in >> n[0] >> n[1] >> n[2];
  try {
    MMesh<2> grid(0);
    grid.setSnapshot(s);
    grid.indicator().init();
    std::printf("params %.17g %.17g %.17g\n", grid.indicator().minH(), grid.indicator().maxH(), grid.indicator().factor());
    grid.indicator().maxH() *= 0.45;            // users tune the members directly (test-femadaptation.cc:219-220)
    grid.indicator().minH() *= 3.0;
    const bool marked = grid.markElements();
    std::printf("marked %d\nmarks", (int)marked);
    for (std::size_t c = 0; c < nc; ++c) std::printf(" %d", grid.getMark(c));
    std::printf("\nlongest");
    for (std::size_t c = 0; c < nc; ++c) std::printf(" %d", grid.longestEdge(c));
    std::printf("\nmaxdist %.17g\n", grid.distance().maximum());
    {                                             // the candidate loops of adapt() (mmesh.hh:826-908)
      const AdaptCandidates a = grid.adaptCandidates();
      std::printf("insertCells");
      for (int32_t c : a.insertCells) std::printf(" %d", c);
      std::printf("\ninsertPoints");
      for (auto& p : a.insertPoints) std::printf(" %.17g %.17g", p[0], p[1]);
      std::printf("\nremoveCells");
      for (int32_t c : a.removeCells) std::printf(" %d", c);
      std::printf("\nremoveVertices");
      for (int32_t c : a.removeVertices) std::printf(" %d", c);
      std::printf("\ninsertSegments");
      for (int32_t c : a.insertSegments) std::printf(" %d", c);
      std::printf("\nremoveSegmentVertices");
      for (int32_t c : a.removeSegmentVertices) std::printf(" %d", c);
      std::printf("\n");
    }
    const bool change = grid.ensureVertexMovement(shifts);
    std::printf("change %d\nremoved", (int)change);
    for (int32_t v : grid.removed()) std::printf(" %d", v);
    std::printf("\n");
    grid.moveVertices(shifts);
    const auto legal = grid.edgeLegality();
    std::printf("legal");
    for (uint8_t f : legal) std::printf(" %d", (int)f);
    std::printf("\n");
    grid.postAdapt();
    std::printf("preAdapt %d\n", (int)grid.preAdapt());
    // interface movement (mmesh.hh:920-1088, :1407-1416): the moved mesh above may contain inverted cells, so the
    // pre-check of ensureInterfaceMovement must throw GridError exactly when the oracle sees an invalid cell
    for (std::size_t v = 0; v < nv; ++v) if (s.vflags[v] & 1) s.interfaceVertices.push_back((int32_t)v);
    std::vector<GlobalCoordinate<2>> ish(s.interfaceVertices.size(), GlobalCoordinate<2>{ 0.0, 1e-3 });
    try { auto bad = grid.invertedCellsUnderInterfaceMovement(ish); std::printf("ifmove ok %zu\n", bad.size()); }
    catch (const GridError&) { std::printf("ifmove GridError\n"); }
    // wrong-sized shifts are an error, as the assert at mmesh.hh:1095
    shifts.pop_back();
    try { grid.moveVertices(shifts); std::printf("sizecheck missing\n"); }
    catch (const InvalidStateException&) { std::printf("sizecheck ok\n"); }
    {   // the same step as ONE pipelined call on a fresh grid: identical fields
      MMesh<2> g2(0);
      s.interfaceVertices.clear();
      g2.setSnapshot(s);
      g2.indicator().init();
      g2.indicator().maxH() *= 0.45;
      g2.indicator().minH() *= 3.0;
      shifts.resize(nv);
      std::ifstream in2(argv[1]);
      { std::size_t a, b, c, d; in2 >> a >> b >> c >> d; double t; for (std::size_t i = 0; i < 2 * nv; ++i) in2 >> t; for (auto& p : shifts) in2 >> p[0] >> p[1]; }
      const mmb_step_result r = g2.step(shifts);
      std::printf("step_counts %lld %lld %lld %lld\nstep_moved %lld\nstep_marks", (long long)r.n_refine, (long long)r.n_coarsen, (long long)r.n_invalid, (long long)r.n_illegal, (long long)r.moved);
      for (std::size_t c = 0; c < nc; ++c) std::printf(" %d", g2.getMark(c));
      std::printf("\nstep_valid");
      for (uint8_t f : g2.validity()) std::printf(" %d", (int)f);
      std::printf("\nstep_legal");
      for (uint8_t f : g2.edgeFlags()) std::printf(" %d", (int)f);
      std::printf("\n");
      // the trial move inverts cells, so the default policy withheld moveVertices: a second step with move_policy 1 applies it
      g2.setMovePolicy(1);
      const mmb_step_result r1 = g2.step(shifts);
      std::printf("step1_moved %lld %lld\nstep1_legal", (long long)r1.moved, (long long)r1.n_invalid);
      for (uint8_t f : g2.edgeFlags()) std::printf(" %d", (int)f);
      std::printf("\n");
    }
    {   // Delaunay restoration by device flips: the snapshot object receives the flipped topology
      std::ifstream in3(argv[1]);
      { std::size_t a, b, c, d; in3 >> a >> b >> c >> d; for (auto& p : s.x) in3 >> p[0] >> p[1]; }
      MMesh<2> g3(0);
      g3.setSnapshot(s);
      const mmb_flip_result fr = g3.restoreDelaunay(s);
      const auto flog = g3.lastFlips();
      std::printf("fliplog %lld %zu %zu\n", (long long)flog.total, flog.flips.size(), flog.roundOffsets.size());
      std::printf("flip %lld %d %d %lld\nflipcells", (long long)fr.n_flips, (int)fr.rounds, (int)fr.converged, (long long)fr.n_constrained_illegal);
      for (auto& c : s.cells) std::printf(" %d %d %d", c[0], c[1], c[2]);
      std::printf("\nflipedges");
      for (auto& e : s.edges) std::printf(" %d %d %d %d", e[0], e[1], e[2], e[3]);
      const auto legal3 = g3.edgeLegality();
      std::size_t nill = 0; for (uint8_t f : legal3) nill += (f == 0);
      std::printf("\nflipillegal %zu\n", nill);
      // an inverted cell makes the flip pass refuse (GridError), like the reference refuses to move into an invalid mesh
      Snapshot<2> bad = s;
      const auto& c0 = bad.cells[nc / 2];
      for (int k = 0; k < 2; ++k) bad.x[(std::size_t)c0[0]][k] = 2.0 * bad.x[(std::size_t)c0[1]][k] - bad.x[(std::size_t)c0[0]][k] + 2.0 * (bad.x[(std::size_t)c0[2]][k] - bad.x[(std::size_t)c0[1]][k]);
      g3.setSnapshot(bad);
      try { g3.restoreDelaunay(bad); std::printf("flipinverted missing\n"); }
      catch (const GridError&) { std::printf("flipinverted GridError\n"); }
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
