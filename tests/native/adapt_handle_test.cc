// Drives Dune::B200::MMesh<2>::adapt(handle, edit) the way dune-fem's AdaptationManager drives Dune::MMesh::adapt(handle)
// (mmesh.hh:1138-1165): the grid holds the old snapshot, `edit` plays the host adapter's CGAL TDS edits and returns the new
// snapshot, the handle moves DOF blocks.  Input: a binary file written by tests/test_gpu_host_api.py; output: the DOFs of
// every cell of the new snapshot (cells that are not isNew keep the value the handle carried over), one line.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mmesh_b200.hh"
using namespace Dune::B200;

template <class T> static std::vector<T> rd(std::FILE* f, std::size_t n) { std::vector<T> v(n); if (n && std::fread(v.data(), sizeof(T), n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); } return v; }
static Snapshot<2> read_snapshot(std::FILE* f) {
  const auto h = rd<int64_t>(f, 4);            // nv, nc, ne, ns
  Snapshot<2> s;
  const auto x = rd<double>(f, 2 * h[0]); const auto id = rd<int64_t>(f, h[0]);
  const auto c = rd<int32_t>(f, 3 * h[1]); const auto p = rd<uint8_t>(f, h[1]);
  const auto e = rd<int32_t>(f, 4 * h[2]); const auto sg = rd<int32_t>(f, 2 * h[3]);
  s.x.resize(h[0]); s.cells.resize(h[1]); s.edges.resize(h[2]); s.facets.resize(h[3]); s.perm = p; s.vertexId = id;
  for (int64_t i = 0; i < h[0]; ++i) s.x[i] = { x[2 * i], x[2 * i + 1] };
  for (int64_t i = 0; i < h[1]; ++i) s.cells[i] = { c[3 * i], c[3 * i + 1], c[3 * i + 2] };
  for (int64_t i = 0; i < h[2]; ++i) s.edges[i] = { e[4 * i], e[4 * i + 1], e[4 * i + 2], e[4 * i + 3] };
  for (int64_t i = 0; i < h[3]; ++i) s.facets[i] = { sg[2 * i], sg[2 * i + 1] };
  return s;
}
// what a dune-fem shim looks like from here: DOF blocks keyed by the entity (persistentcontainer.hh:33-89)
struct VectorDataHandle {
  int ncomp; const std::vector<double>& oldDofs; std::vector<double>& newDofs; std::size_t gathers = 0, scatters = 0;
  int localDofs() const { return ncomp; }
  void gather(int32_t oldCell, double* out) { for (int k = 0; k < ncomp; ++k) out[k] = oldDofs[(std::size_t)oldCell * ncomp + k]; ++gathers; }
  void scatter(int32_t newCell, const double* in) { for (int k = 0; k < ncomp; ++k) newDofs[(std::size_t)newCell * ncomp + k] = in[k]; ++scatters; }
};

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  const Snapshot<2> oldS = read_snapshot(f), newS = read_snapshot(f);
  const auto hd = rd<int64_t>(f, 1);
  const int ncomp = (int)hd[0];
  const auto oldDofs = rd<double>(f, oldS.cells.size() * ncomp);
  std::vector<double> newDofs = rd<double>(f, newS.cells.size() * ncomp);      // values carried over for unchanged cells
  std::fclose(f);
  try {
    MMesh<2> grid(0);
    grid.setSnapshot(oldS);
    VectorDataHandle handle{ ncomp, oldDofs, newDofs };
    std::size_t edits = 0;
    const bool ok = grid.adapt(handle, [&](const AdaptCandidates&, const std::vector<int32_t>&) -> const Snapshot<2>& { ++edits; return newS; });
    const ProjectionPairs& pr = grid.lastProjectionPairs();
    std::printf("adapt %d edits %zu gathers %zu scatters %zu nnew %zu npairs %zu components %zu\ndofs", (int)ok, edits, handle.gathers, handle.scatters,
                pr.newCell.size(), pr.oldIndex.size(), pr.components);
    for (double v : newDofs) std::printf(" %.17g", v);
    std::printf("\n");
    // the grid now works on the new snapshot: the next step's calls see its sizes
    std::vector<GlobalCoordinate<2>> zero(newS.x.size(), GlobalCoordinate<2>{ 0.0, 0.0 });
    const bool change = grid.ensureVertexMovement(zero);
    std::printf("after change %d\n", (int)change);
  } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
  return 0;
}
