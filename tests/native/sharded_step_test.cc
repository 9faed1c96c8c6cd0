// A C++ caller of the sharded step: partitions the global snapshot with Dune::B200::makeShard (host/partition.hh), creates one
// context on GPU `rank`, joins the library's NCCL communicator (the unique id travels through a file, as it would through
// MPI_Bcast in the reference's setting, communication.hh:83-127) and runs mmb_adapt_step_sharded.  Launched once per rank by
// tests/test_gpu_multi.py, which compares the printed GLOBAL scalars and the owned marks with a single-GPU run.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "mmesh_b200.hh"
#include "partition.hh"
using namespace Dune::B200;

template <class T> static std::vector<T> rd(std::FILE* f, std::size_t n) { std::vector<T> v(n); if (n && std::fread(v.data(), sizeof(T), n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); } return v; }
#define CHECK(call) do { const int rc_ = (call); if (rc_ != MMB_OK && rc_ != MMB_MOVE_WITHHELD) { std::fprintf(stderr, "rank %d: %s -> %d (%s)\n", rank, #call, rc_, mmb_last_error(ctx)); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: sharded_step_test snapshot.bin rank world idfile\n"); return 2; }
  const int rank = std::atoi(argv[2]), world = std::atoi(argv[3]);
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  const auto h = rd<int64_t>(f, 6);            // nv, nc, ne, ns, npairs, move_policy
  Snapshot<2> g;
  const auto x = rd<double>(f, 2 * h[0]); const auto c = rd<int32_t>(f, 3 * h[1]); g.perm = rd<uint8_t>(f, h[1]);
  const auto e = rd<int32_t>(f, 4 * h[2]); const auto ec = rd<int32_t>(f, h[2]); const auto sg = rd<int32_t>(f, 2 * h[3]);
  const auto shifts = rd<double>(f, 2 * h[0]);
  const auto newOff = rd<int64_t>(f, h[1] + 1); const auto newCell = rd<int32_t>(f, h[1]);
  const auto oxy = rd<double>(f, 6 * h[4]); const auto oldIdx = rd<int32_t>(f, h[4]);
  const auto dofs = rd<double>(f, 3 * h[1]);
  const auto prmv = rd<double>(f, 3);          // minH, maxH, factor of the GLOBAL mesh
  std::fclose(f);
  g.x.resize(h[0]); g.cells.resize(h[1]); g.edges.resize(h[2]); g.facets.resize(h[3]);
  for (int64_t i = 0; i < h[0]; ++i) g.x[i] = { x[2 * i], x[2 * i + 1] };
  for (int64_t i = 0; i < h[1]; ++i) g.cells[i] = { c[3 * i], c[3 * i + 1], c[3 * i + 2] };
  for (int64_t i = 0; i < h[2]; ++i) g.edges[i] = { e[4 * i], e[4 * i + 1], e[4 * i + 2], e[4 * i + 3] };
  for (int64_t i = 0; i < h[3]; ++i) g.facets[i] = { sg[2 * i], sg[2 * i + 1] };
  std::vector<std::array<double, 6>> oldCorners(h[4]);
  std::memcpy(oldCorners.data(), oxy.data(), oxy.size() * sizeof(double));
  const Shard sh = makeShard(g, ec, world, rank, &newOff, &newCell, &oldCorners, &oldIdx);

  mmb_context* ctx = nullptr;
  if (mmb_create(rank, 2, nullptr, &ctx) != MMB_OK) { std::fprintf(stderr, "rank %d: no device\n", rank); return 1; }
  uint8_t id[128];
  if (rank == 0) {
    CHECK(mmb_comm_unique_id(id));
    std::FILE* o = std::fopen((std::string(argv[4]) + ".tmp").c_str(), "wb"); std::fwrite(id, 1, 128, o); std::fclose(o);
    std::rename((std::string(argv[4]) + ".tmp").c_str(), argv[4]);
  } else {
    for (int t = 0; t < 3000; ++t) {
      std::FILE* i = std::fopen(argv[4], "rb");
      if (i) { const std::size_t n = std::fread(id, 1, 128, i); std::fclose(i); if (n == 128) break; }
      std::this_thread::sleep_for(std::chrono::milliseconds(10));
      if (t == 2999) { std::fprintf(stderr, "rank %d: no unique id\n", rank); return 1; }
    }
  }
  CHECK(mmb_comm_init(ctx, world, rank, id));
  const Snapshot<2>& L = sh.local;
  CHECK(mmb_upload_mesh(ctx, (int64_t)L.x.size(), L.x[0].data(), (int64_t)L.cells.size(), L.cells[0].data(), L.perm.data(), (int64_t)L.edges.size(),
                        L.edges.empty() ? nullptr : L.edges[0].data(), (int64_t)L.facets.size(), L.facets.empty() ? nullptr : L.facets[0].data()));
  CHECK(mmb_set_owned_cells(ctx, sh.ownLo, sh.ownHi));
  CHECK(mmb_halo_setup(ctx, (int64_t)sh.sendIdx.size(), sh.sendIdx.data(), (int64_t)sh.recvIdx.size(), sh.recvIdx.data()));
  CHECK(mmb_halo_counts(ctx, sh.sendCounts.data(), sh.recvCounts.data()));
  mmb_params prm; CHECK(mmb_get_params(ctx, &prm));
  prm.minH = prmv[0]; prm.maxH = prmv[1]; prm.factor = prmv[2]; prm.drop_area = 1e-12; prm.move_policy = (int32_t)h[5];
  CHECK(mmb_set_params(ctx, &prm));
  CHECK(mmb_upload_pairs(ctx, (int64_t)sh.newCell.size(), sh.newOff.data(), sh.newCell.data(), (int64_t)sh.oldIndex.size(),
                         sh.oldCorners.empty() ? nullptr : sh.oldCorners[0].data(), sh.oldIndex.data()));
  std::vector<double> ldofs(3 * sh.localCells.size()), lshift(2 * sh.localVertices.size(), 0.0);
  for (std::size_t i = 0; i < sh.localCells.size(); ++i) for (int k = 0; k < 3; ++k) ldofs[3 * i + k] = dofs[3 * (std::size_t)sh.localCells[i] + k];
  for (std::size_t i = 0; i < sh.localVertices.size(); ++i)
    if (sh.vertexOwner[i] == rank) { lshift[2 * i] = shifts[2 * (std::size_t)sh.localVertices[i]]; lshift[2 * i + 1] = shifts[2 * (std::size_t)sh.localVertices[i] + 1]; }
  CHECK(mmb_upload_dofs(ctx, 3, (int64_t)sh.localCells.size(), ldofs.data()));
  mmb_step_result r{};
  for (int rep = 0; rep < 3; ++rep) {            // eager, captured, replayed: each on the original coordinates
    CHECK(mmb_upload_coords(ctx, L.x[0].data()));
    CHECK(mmb_upload_dofs(ctx, 3, (int64_t)sh.localCells.size(), ldofs.data()));
    CHECK(mmb_adapt_step_sharded(ctx, lshift.data(), 3, &r));
  }
  std::vector<int8_t> marks(L.cells.size());
  CHECK(mmb_download_fields(ctx, 0, nullptr, marks.data(), nullptr, nullptr, nullptr, nullptr));
  CHECK(mmb_synchronize(ctx));
  long long msum = 0; for (int64_t i = sh.ownLo; i < sh.ownHi; ++i) msum += (long long)marks[(std::size_t)i] * (long long)((sh.localCells[(std::size_t)i] % 1000) + 1);
  std::printf("rank %d n_invalid %lld n_orient_nonpos %lld n_illegal %lld n_refine %lld n_coarsen %lld n_pieces %lld moved %lld max_dist %.17g mass %.17g cov %.17g marksum %lld\n",
              rank, (long long)r.n_invalid, (long long)r.n_orient_nonpos, (long long)r.n_illegal, (long long)r.n_refine, (long long)r.n_coarsen,
              (long long)r.n_pieces, (long long)r.moved, r.max_dist, r.mass_new, r.covered_area, msum);
  mmb_comm_destroy(ctx);
  mmb_destroy(ctx);
  return 0;
}
