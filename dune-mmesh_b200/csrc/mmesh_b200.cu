// mmesh_b200.cu -- C-ABI implementation (include/mmesh_b200.h) on top of the sm_100a kernels in kernels.cuh.
// Host side is plain C++/CUDA runtime: no torch, no CPU fallback.  Every entry point fails with
// MMB_ERR_CUDA when no usable device / kernel image exists.
#include "../../include/mmesh_b200.h"
#include "kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

using namespace mmb;

namespace {

struct TimedLaunch { const char* name; cudaEvent_t a, b; };

template <class T>
struct DBuf {
  T* p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  // grow to >= n elements keeping the first `keep` ones (device-to-device copy on `stream`)
  cudaError_t grow(size_t n, size_t keep, cudaStream_t stream) {
    if (n <= cap) return cudaSuccess;
    T* q = nullptr;
    const size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc((void**)&q, want * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p && keep) e = cudaMemcpyAsync(q, p, std::min(keep, cap) * sizeof(T), cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (p) cudaFree(p);
    p = q; cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

// number of projection chunks of the pipelined host-buffer step (the tail after the last upload is one chunk's compute +
// download); MMB_CHUNKS overrides
static const int MAX_CHUNKS = 64;
static int pipeline_chunks() {
  static int c = -1;
  if (c < 0) { const char* e = getenv("MMB_CHUNKS"); c = e ? atoi(e) : 32; c = std::max(1, std::min(c, MAX_CHUNKS)); }
  return c;
}

struct mmb_context {
  int device = 0, dim = 2;
  cudaStream_t stream = nullptr; bool own_stream = false;
  std::string err;
  int64_t nv = 0, nc = 0, ne = 0, ns = 0, nc_pad = 0;
  ncclComm_t comm = nullptr; int world = 1, rank = 0; std::vector<int64_t> halo_send_counts, halo_recv_counts;
  cudaStream_t s_comm = nullptr; std::vector<cudaEvent_t> comm_ev;
  cudaGraphExec_t step_graph = nullptr; int graph_ncomp = -1; mmb_params graph_prm{}; bool graph_valid = false; int64_t graph_launches = 0; int sharded_steps = 0;
  bool sharded = false;             // one of several ranks: the gate of the fused step's move comes from an all-reduce
  int64_t own_lo = 0, own_hi = 0;   // owned cell range (counts); halo cells outside it are evaluated but not counted
  DBuf<int32_t> halo_send_idx, halo_recv_idx; DBuf<double2> halo_send, halo_recv; int64_t n_halo_send = 0, n_halo_recv = 0;
  DBuf<double> reduce_vec, red_slices, pair_partial; DBuf<unsigned> red_ticket; DBuf<uint8_t> pair_pieces;
  mmb_params prm{};
  // mesh
  DBuf<double> x, shift, dist;
  DBuf<int32_t> cv[4], cells_aos, facets, facets_sorted;
  DBuf<uint8_t> perm, valid_fp, edge_flag, longest;
  DBuf<int8_t> orient, mark, imark;
  DBuf<int4> edges;
  DBuf<int32_t> wl, list_out;
  DBuf<uint32_t> block_count;
  DBuf<int32_t> tile_lo_cells, tile_lo_edges;
  DBuf<uint32_t> flip_log; std::vector<int64_t> flip_round_off; int64_t flip_log_total = 0;   // last mmb_restore_delaunay
  DBuf<int32_t> cell_nb; bool have_nb = false; DBuf<uint8_t> v_if, cand_flag; DBuf<unsigned long long> facet_keys; int64_t n_facet_keys = 0;
  DBuf<unsigned long long> ht_key; DBuf<int32_t> ht_val, ht_slot;   // inserted_ of the 3-D candidate loop
  DBuf<uint8_t> v_bnd; DBuf<int32_t> v_ifcount, v_level, first_claim, first_seg, chosen; bool cand_ready = false, have_levels = false;
  bool hp_active = false, hp_pipe_dofs = false; int hp_ncomp = 0;   // host-buffer step in flight (begin .. finish)
  DBuf<Counters> counters;
  bool has_shift = false, dist_valid = false, mesh_ok = false;
  // bvh
  int L = 1;
  DBuf<SegLeaf> seg_leaf; DBuf<float4> box2; DBuf<int32_t> seg_hint; DBuf<unsigned> ticket;
  DBuf<TriLeaf> tri_leaf; DBuf<Box3> box3; DBuf<int32_t> seed_grid; bool hints_cold = true;
  // projection
  int64_t nnew = 0, npairs = 0, n_old = 0;
  DBuf<long long> new_off; DBuf<int32_t> new_cell, old_idx; DBuf<double> old_xy, dofs_old, dofs_new, part_a, part_b, pair_vol;
  int dofs_ncomp = 0; int64_t max_old_idx = -1;
  std::vector<int64_t> chunk_i; std::vector<int32_t> chunk_cell_lo, chunk_cell_hi; bool chunks_ok = false;
  // pipelined DOF upload: chunk k reads old DOFs from its home index range [dof_lo[k], dof_lo[k+1]) plus a few remote ones
  bool dof_pipe_ok = false; std::vector<int64_t> dof_lo; std::vector<int32_t> remote_idx; DBuf<int32_t> d_remote_idx;
  DBuf<double> d_remote_stage; double* h_remote_stage = nullptr; size_t h_remote_cap = 0;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr; std::vector<cudaEvent_t> pipe_ev;
  // 3-D trace
  DBuf<int32_t> tr_in, tr_out; DBuf<double> tr_u, tr_ug, tr_a, tr_b, tr_c;
  // misc
  DBuf<double> pred_pts; DBuf<int8_t> pred_sign;
  Counters* h_counters = nullptr;   // pinned
  bool timing = false;
  std::vector<TimedLaunch> timed;
  std::vector<cudaEvent_t> event_pool;
  int64_t launches = 0;
  int num_sms = 148;
};

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                     \
      return MMB_ERR_CUDA;                                                                               \
    }                                                                                                    \
  } while (0)
#define REQUIRE(cond, code, msg)                                                                         \
  do { if (!(cond)) { ctx->err = msg; return code; } } while (0)

static cudaEvent_t get_event(mmb_context* ctx) {
  if (!ctx->event_pool.empty()) { cudaEvent_t e = ctx->event_pool.back(); ctx->event_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct LaunchScope {
  mmb_context* ctx; const char* name; cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(mmb_context* c, const char* n) : ctx(c), name(n) {
    ctx->launches++;
    if (ctx->timing) { a = get_event(ctx); b = get_event(ctx); cudaEventRecord(a, ctx->stream); }
  }
  ~LaunchScope() { if (ctx->timing) { cudaEventRecord(b, ctx->stream); ctx->timed.push_back({ name, a, b }); } }
};
#define LAUNCH(name, kernel, grid, block, ...)                                                           \
  do {                                                                                                   \
    LaunchScope ls_(ctx, name);                                                                          \
    kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__);                                            \
  } while (0)

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)std::max<int64_t>(1, (n + block - 1) / block); }
static DevParams dev_params(const mmb_params& p) {
  DevParams d; d.minH = p.minH; d.maxH = p.maxH; d.factor = p.factor; d.distProportion = p.distProportion;
  d.edgeRatio = p.edgeRatio; d.radiusRatio = p.radiusRatio; d.clip_eps = p.clip_eps; d.drop_area = p.drop_area;
  d.clip_translate = p.clip_translate; return d;
}
static int fetch_counters(mmb_context* ctx) {
  CK(cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MMB_OK;
}
// predicates whose inputs exceed the exponent range of the exact stage (predicates.cuh: pred_rescale) were counted on the
// device; their signs were set to 0 and must not be trusted
static int check_predicate_domain(mmb_context* ctx) {
  if (ctx->h_counters->pad_[1] != 0) {
    ctx->err = "exact predicate: " + std::to_string((unsigned long long)ctx->h_counters->pad_[1]) +
               " input(s) outside the exponent range of the exact stage (largest / smallest nonzero magnitude > 2^480 / 2^300 / 2^212 "
               "for orient2d / orient3d / in-circle, or non-finite coordinates); CGAL's rational fallback has no such limit";
    return MMB_ERR_CAPACITY;
  }
  return MMB_OK;
}
#define TRY(expr) do { int rc_ = (expr); if (rc_ != MMB_OK) return rc_; } while (0)

// ---------------------------------------------------------------------------------------------------------
extern "C" {

int mmb_create(int device, int dim, void* stream, mmb_context** out) {
  if (!out || (dim != 2 && dim != 3)) return MMB_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return MMB_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return MMB_ERR_CUDA;
  // fail loudly when this build has no kernel image for the device (sm_100a only)
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, k_move) != cudaSuccess) { cudaGetLastError(); return MMB_ERR_CUDA; }
  mmb_context* ctx = new mmb_context();
  ctx->device = device; ctx->dim = dim;
  if (stream) ctx->stream = (cudaStream_t)stream;
  else { if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MMB_ERR_CUDA; } ctx->own_stream = true; }
  cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (ctx->counters.ensure(1) != cudaSuccess || cudaMallocHost((void**)&ctx->h_counters, sizeof(Counters)) != cudaSuccess) {
    delete ctx; return MMB_ERR_CUDA;
  }
  cudaMemsetAsync(ctx->counters.p, 0, sizeof(Counters), ctx->stream);
  ctx->prm = mmb_params{ 0.0, 0.0, 1.0, 1.0, 4.0, 30.0, 1e-14, 1e-8, 0, 0 };
  *out = ctx;
  return MMB_OK;
}

int mmb_destroy(mmb_context* ctx) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& t : ctx->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  ctx->x.release(); ctx->shift.release(); ctx->dist.release();
  for (auto& b : ctx->cv) b.release();
  ctx->cells_aos.release(); ctx->facets.release(); ctx->facets_sorted.release();
  ctx->perm.release(); ctx->valid_fp.release(); ctx->edge_flag.release(); ctx->longest.release();
  ctx->orient.release(); ctx->mark.release(); ctx->imark.release(); ctx->edges.release();
  ctx->wl.release(); ctx->list_out.release(); ctx->block_count.release(); ctx->counters.release();
  ctx->tile_lo_cells.release(); ctx->tile_lo_edges.release();
  ctx->flip_log.release(); ctx->cell_nb.release(); ctx->v_if.release(); ctx->cand_flag.release(); ctx->facet_keys.release(); ctx->ht_key.release(); ctx->ht_val.release(); ctx->ht_slot.release();
  ctx->v_bnd.release(); ctx->v_ifcount.release(); ctx->v_level.release(); ctx->first_claim.release(); ctx->first_seg.release(); ctx->chosen.release();
  ctx->seg_leaf.release(); ctx->box2.release(); ctx->seg_hint.release(); ctx->ticket.release(); ctx->tri_leaf.release(); ctx->box3.release(); ctx->seed_grid.release();
  ctx->new_off.release(); ctx->new_cell.release(); ctx->old_idx.release(); ctx->old_xy.release();
  ctx->dofs_old.release(); ctx->dofs_new.release(); ctx->part_a.release(); ctx->part_b.release(); ctx->pair_vol.release();
  ctx->tr_in.release(); ctx->tr_out.release(); ctx->tr_u.release(); ctx->tr_ug.release();
  ctx->tr_a.release(); ctx->tr_b.release(); ctx->tr_c.release(); ctx->pred_pts.release(); ctx->pred_sign.release();
  ctx->halo_send_idx.release(); ctx->halo_recv_idx.release(); ctx->halo_send.release(); ctx->halo_recv.release(); ctx->reduce_vec.release(); ctx->red_slices.release(); ctx->red_ticket.release(); ctx->pair_partial.release(); ctx->pair_pieces.release();
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  if (ctx->h_remote_stage) cudaFreeHost(ctx->h_remote_stage);
  ctx->d_remote_idx.release(); ctx->d_remote_stage.release();
  mmb_comm_destroy(ctx);
  if (ctx->s_comm) cudaStreamDestroy(ctx->s_comm);
  for (auto e : ctx->comm_ev) cudaEventDestroy(e);
  if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
  if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
  for (auto e : ctx->pipe_ev) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return MMB_OK;
}

const char* mmb_last_error(const mmb_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* mmb_stream(const mmb_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int mmb_synchronize(mmb_context* ctx) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  CK(cudaStreamSynchronize(ctx->stream));
  return MMB_OK;
}
int64_t mmb_launch_count(const mmb_context* ctx) { return ctx ? ctx->launches : 0; }
#ifdef MMB_DEV_VARIANTS
// development counters of the last kernels (instrumented builds only; not part of the C-ABI header)
int mmb_dev_counters(mmb_context* ctx, unsigned long long out[8]) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  if (cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return MMB_ERR_CUDA;
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < 8; ++i) out[i] = ctx->h_counters->dbg[i];
  return MMB_OK;
}
#endif

int mmb_host_register(mmb_context* ctx, void* ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return MMB_ERR_INVALID_ARG;
  const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) { cudaGetLastError(); if (ctx) ctx->err = std::string("cudaHostRegister: ") + cudaGetErrorString(e); return MMB_ERR_CUDA; }
  return MMB_OK;
}
int mmb_host_unregister(mmb_context* ctx, void* ptr) {
  if (!ptr) return MMB_ERR_INVALID_ARG;
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); if (ctx) ctx->err = std::string("cudaHostUnregister: ") + cudaGetErrorString(e); return MMB_ERR_CUDA; }
  return MMB_OK;
}
int mmb_set_params(mmb_context* ctx, const mmb_params* p) {
  if (!ctx || !p) return MMB_ERR_INVALID_ARG;
  ctx->prm = *p; return MMB_OK;
}
int mmb_get_params(const mmb_context* ctx, mmb_params* p) {
  if (!ctx || !p) return MMB_ERR_INVALID_ARG;
  *p = ctx->prm; return MMB_OK;
}

// Morton key of a point inside a bounding box (host; upload time only)
static uint64_t morton2(double x, double y, const double* lo, const double* inv) {
  auto q = [](double t) { t = std::min(std::max(t, 0.0), 1.0); return (uint32_t)(t * 65535.0); };
  auto spread = [](uint64_t v) {
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull; v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full; v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull; return v;
  };
  return spread(q((x - lo[0]) * inv[0])) | (spread(q((y - lo[1]) * inv[1])) << 1);
}
static uint64_t morton3(const double* c, const double* lo, const double* inv) {
  auto q = [](double t) { t = std::min(std::max(t, 0.0), 1.0); return (uint64_t)(t * 1023.0); };
  auto spread = [](uint64_t v) {
    v = (v | (v << 16)) & 0x030000FFull; v = (v | (v << 8)) & 0x0300F00Full;
    v = (v | (v << 4)) & 0x030C30C3ull; v = (v | (v << 2)) & 0x09249249ull; return v;
  };
  return spread(q((c[0] - lo[0]) * inv[0])) | (spread(q((c[1] - lo[1]) * inv[1])) << 1) | (spread(q((c[2] - lo[2]) * inv[2])) << 2);
}

// Everything that depends on the interface facets: original order (marks) + Morton order (BVH leaves), interface-vertex
// flags, sorted interface-edge keys, BVH storage, and the distance hints (reset: they index the Morton order).
static int setup_facets(mmb_context* ctx, int64_t ns, const int32_t* facets_host, const double* x_host) {
  const int dim = ctx->dim;
  const int64_t nv = ctx->nv;
  ctx->ns = ns;
  // interface facets: original order (marks) + Morton order (BVH leaves)
  CK(ctx->facets.ensure((size_t)ns * dim + 2)); CK(ctx->facets_sorted.ensure((size_t)ns * dim + 2)); CK(ctx->imark.ensure(ns + 4));
  int L = 1; while (L < ns) L <<= 1;
  ctx->L = L;
  if (ns) {
    CK(cudaMemcpyAsync(ctx->facets.p, facets_host, (size_t)ns * dim * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 }, inv[3];
    std::vector<double> mid((size_t)ns * dim);
    for (int64_t s = 0; s < ns; ++s)
      for (int j = 0; j < dim; ++j) {
        double m = 0;
        for (int c = 0; c < dim; ++c) m += x_host[(size_t)facets_host[s * dim + c] * dim + j];
        m /= dim; mid[s * dim + j] = m; lo[j] = std::min(lo[j], m); hi[j] = std::max(hi[j], m);
      }
    for (int j = 0; j < dim; ++j) inv[j] = hi[j] > lo[j] ? 1.0 / (hi[j] - lo[j]) : 0.0;
    std::vector<std::pair<uint64_t, int32_t>> keyed(ns);
    for (int64_t s = 0; s < ns; ++s)
      keyed[s] = { dim == 2 ? morton2(mid[2 * s], mid[2 * s + 1], lo, inv) : morton3(&mid[3 * s], lo, inv), (int32_t)s };
    std::sort(keyed.begin(), keyed.end());
    std::vector<int32_t> sorted((size_t)ns * dim);
    for (int64_t s = 0; s < ns; ++s)
      for (int c = 0; c < dim; ++c) sorted[s * dim + c] = facets_host[(size_t)keyed[s].second * dim + c];
    CK(cudaMemcpyAsync(ctx->facets_sorted.p, sorted.data(), sorted.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // `sorted` is a stack-owned staging buffer
  }
  ctx->cand_ready = false;
  {
    // per-vertex interface flag + sorted (min,max) keys of the interface edges: MMesh::isInterface(edge) on the device
    // (2-D: the facets themselves, mmesh.hh:607-621; 3-D: the three edges of every interface triangle, :624-645)
    const int64_t nk = dim == 2 ? ns : 3 * ns;
    CK(ctx->v_if.ensure(nv + 1)); CK(ctx->cand_flag.ensure((size_t)std::max(ctx->nc_pad, ns) + 4)); CK(ctx->facet_keys.ensure(nk + 1));
    CK(cudaMemsetAsync(ctx->v_if.p, 0, (size_t)nv, ctx->stream));
    ctx->n_facet_keys = nk;
    if (ns) {
      LAUNCH("flag_interface_vertices", k_flag_interface_vertices, grid_for(dim * ns, 256), 256, ctx->facets.p, dim * ns, ctx->v_if.p);
      std::vector<unsigned long long> keys; keys.reserve((size_t)nk);
      auto key = [](int32_t a, int32_t b) { const unsigned x = (unsigned)std::min(a, b), y = (unsigned)std::max(a, b); return ((unsigned long long)x << 32) | y; };
      for (int64_t s2 = 0; s2 < ns; ++s2) {
        const int32_t* f = facets_host + (size_t)s2 * dim;
        if (dim == 2) keys.push_back(key(f[0], f[1]));
        else { keys.push_back(key(f[0], f[1])); keys.push_back(key(f[0], f[2])); keys.push_back(key(f[1], f[2])); }
      }
      std::sort(keys.begin(), keys.end());
      CK(cudaMemcpy(ctx->facet_keys.p, keys.data(), (size_t)nk * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
  }
  if (dim == 2) {
    CK(ctx->seg_leaf.ensure(L)); CK(ctx->box2.ensure(2 * (size_t)L)); CK(ctx->seg_hint.ensure(nv)); CK(ctx->ticket.ensure(1));
    CK(cudaMemsetAsync(ctx->seg_hint.p, 0xFF, (size_t)nv * sizeof(int32_t), ctx->stream));    // -1: no hint yet
    ctx->hints_cold = true;
    CK(cudaMemsetAsync(ctx->ticket.p, 0, sizeof(unsigned), ctx->stream));
  }
  else {
    CK(ctx->tri_leaf.ensure(L)); CK(ctx->box3.ensure(2 * (size_t)L)); CK(ctx->seg_hint.ensure(nv)); CK(ctx->ticket.ensure(1));
    CK(cudaMemsetAsync(ctx->seg_hint.p, 0xFF, (size_t)nv * sizeof(int32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->ticket.p, 0, sizeof(unsigned), ctx->stream));
  }
  return MMB_OK;
}

int mmb_upload_mesh(mmb_context* ctx, int64_t nv, const double* x_host, int64_t nc, const int32_t* cells_host,
                    const uint8_t* perm_host, int64_t ne, const int32_t* edges_host, int64_t ns,
                    const int32_t* facets_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(nv > 0 && nc >= 0 && ne >= 0 && ns >= 0 && x_host && (nc == 0 || (cells_host && perm_host)),
          MMB_ERR_INVALID_ARG, "mmb_upload_mesh: bad sizes or null arrays");
  REQUIRE(nv < (1ll << 31) && nc < (1ll << 31) && ne < (1ll << 31) && ns < (1ll << 24), MMB_ERR_CAPACITY,
          "mmb_upload_mesh: int32 index space exceeded");
  REQUIRE((ne == 0 || edges_host) && (ns == 0 || facets_host), MMB_ERR_INVALID_ARG, "mmb_upload_mesh: null edges/facets");
  CK(cudaSetDevice(ctx->device));
  const int dim = ctx->dim, k = dim + 1;
  ctx->mesh_ok = false;
  // the kernels trust the snapshot: reject out-of-range indices here instead of reading out of bounds on the device
  {
    int64_t bad = 0;
    for (int64_t i = 0; i < nc * k; ++i) bad += (cells_host[i] < 0 || cells_host[i] >= nv);
    for (int64_t i = 0; i < nc; ++i) {
      const unsigned p = perm_host[i];
      unsigned seen = 0;
      for (int j = 0; j < k; ++j) seen |= 1u << ((p >> (2 * j)) & 3u);
      bad += (seen != (1u << k) - 1u);                    // perm must be a permutation of 0..dim
    }
    if (dim == 2) for (int64_t i = 0; i < ne; ++i) {
      const int32_t* e = edges_host + 4 * i;
      bad += (e[0] < 0 || e[0] >= nv || e[1] < 0 || e[1] >= nv || e[2] < 0 || e[2] >= nv || e[3] < -1 || e[3] >= nv);
    }
    for (int64_t i = 0; i < ns * dim; ++i) bad += (facets_host[i] < 0 || facets_host[i] >= nv);
    REQUIRE(bad == 0, MMB_ERR_INVALID_ARG, "mmb_upload_mesh: vertex index out of range or invalid corner permutation");
  }
  ctx->nv = nv; ctx->nc = nc; ctx->ne = ne; ctx->ns = ns;
  ctx->own_lo = 0; ctx->own_hi = nc; ctx->n_halo_send = ctx->n_halo_recv = 0;
  ctx->nc_pad = (nc + 3) / 4 * 4;
  const size_t nx = (size_t)nv * dim + 2;
  CK(ctx->x.ensure(nx)); CK(ctx->shift.ensure(nx)); CK(ctx->dist.ensure(nv));
  CK(cudaMemsetAsync(ctx->x.p, 0, nx * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(ctx->shift.p, 0, nx * sizeof(double), ctx->stream));
  CK(cudaMemcpyAsync(ctx->x.p, x_host, (size_t)nv * dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  for (int j = 0; j < k; ++j) CK(ctx->cv[j].ensure(ctx->nc_pad + 4));
  CK(ctx->cells_aos.ensure((size_t)nc * k + 4));
  CK(ctx->perm.ensure(ctx->nc_pad + 4)); CK(ctx->valid_fp.ensure(ctx->nc_pad + 4)); CK(ctx->orient.ensure(ctx->nc_pad + 4));
  CK(ctx->mark.ensure(ctx->nc_pad + 4)); CK(ctx->longest.ensure(ctx->nc_pad + 4));
  CK(ctx->wl.ensure(std::max(std::max(ctx->nc_pad, ne), nv) + 4)); CK(ctx->list_out.ensure(std::max(ctx->nc_pad, ne) + 4));
  CK(ctx->block_count.ensure((std::max(ctx->nc_pad, ne) + COMPACT_TILE - 1) / COMPACT_TILE + 1));
  if (nc) {
    CK(cudaMemcpyAsync(ctx->cells_aos.p, cells_host, (size_t)nc * k * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->perm.p, perm_host, (size_t)nc, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH("cells_to_planes", k_cells_to_planes, grid_for(ctx->nc_pad, 256), 256, ctx->cells_aos.p, k, nc, ctx->nc_pad,
           ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, k == 4 ? ctx->cv[3].p : nullptr);
  }
  if (dim == 2) {
    CK(ctx->edges.ensure(ne + 1)); CK(ctx->edge_flag.ensure(ne + 4));
    if (ne) CK(cudaMemcpyAsync(ctx->edges.p, edges_host, (size_t)ne * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
    // Hilbert tile windows (see kernels.cuh): one per TILE cells / TILE edges
    const int64_t ntc = (ctx->nc_pad + TILE - 1) / TILE, nte = (ne + TILE - 1) / TILE;
    CK(ctx->tile_lo_cells.ensure(ntc + 1)); CK(ctx->tile_lo_edges.ensure(nte + 1));
    CK(cudaMemsetAsync(ctx->tile_lo_cells.p, 0, (ntc + 1) * sizeof(int32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->tile_lo_edges.p, 0, (nte + 1) * sizeof(int32_t), ctx->stream));
    if (nc) LAUNCH("tile_windows", k_tile_windows, (unsigned)ntc, 256, ctx->cv[0].p, 1, nc, nv, ctx->tile_lo_cells.p);
    if (ne) LAUNCH("tile_windows", k_tile_windows, (unsigned)nte, 256, (const int32_t*)ctx->edges.p, 4, ne, nv, ctx->tile_lo_edges.p);
  }
  ctx->have_nb = false; ctx->have_levels = false;
  TRY(setup_facets(ctx, ns, facets_host, x_host));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  ctx->has_shift = false; ctx->dist_valid = false; ctx->mesh_ok = true;
  ctx->nnew = ctx->npairs = 0; ctx->graph_valid = false; ctx->sharded_steps = 0;
  return MMB_OK;
}

int mmb_upload_coords(mmb_context* ctx, const double* x_host) {
  if (!ctx || !x_host) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  CK(cudaMemcpyAsync(ctx->x.p, x_host, (size_t)ctx->nv * ctx->dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->dist_valid = false;
  return MMB_OK;
}
int mmb_download_coords(mmb_context* ctx, double* x_host) {
  if (!ctx || !x_host) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  CK(cudaMemcpyAsync(x_host, ctx->x.p, (size_t)ctx->nv * ctx->dim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MMB_OK;
}
// Incremental snapshot update after a small TDS edit (what adapt_() changes, mmesh.hh:1338-1355): only the listed vertices,
// cells and edges are transferred; the resident arrays grow in place; per-vertex distance hints of untouched vertices survive
// when the interface is unchanged, so the next Distance::update starts warm.
int mmb_patch_mesh(mmb_context* ctx, int64_t nv_new, int64_t n_vert, const int32_t* vert_idx_host, const double* vert_x_host,
                   int64_t nc_new, int64_t n_cell, const int32_t* cell_idx_host, const int32_t* cell_v_host, const uint8_t* cell_perm_host,
                   int64_t ne_new, int64_t n_edge, const int32_t* edge_idx_host, const int32_t* edge_q_host,
                   int64_t ns_new, const int32_t* facets_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "mmb_patch_mesh needs an uploaded mesh");
  const int dim = ctx->dim, k = dim + 1;
  REQUIRE(nv_new > 0 && nc_new >= 0 && ne_new >= 0 && n_vert >= 0 && n_cell >= 0 && n_edge >= 0 && (n_vert == 0 || (vert_idx_host && vert_x_host)) &&
          (n_cell == 0 || (cell_idx_host && cell_v_host && cell_perm_host)) && (n_edge == 0 || (dim == 2 && edge_idx_host && edge_q_host)),
          MMB_ERR_INVALID_ARG, "mmb_patch_mesh: bad sizes or null arrays");
  REQUIRE(nv_new < (1ll << 31) && nc_new < (1ll << 31) && ne_new < (1ll << 31), MMB_ERR_CAPACITY, "mmb_patch_mesh: int32 index space exceeded");
  REQUIRE(dim == 2 || ne_new == 0, MMB_ERR_INVALID_ARG, "mmb_patch_mesh: edges are a 2-D concept");
  const int64_t nv0 = ctx->nv, nc0 = ctx->nc, ne0 = ctx->ne;
  {  // the kernels trust the snapshot: validate the patch here; every index beyond the old sizes must be supplied
    int64_t bad = 0;
    std::vector<uint8_t> seen_v((size_t)std::max<int64_t>(nv_new - nv0, 0), 0), seen_c((size_t)std::max<int64_t>(nc_new - nc0, 0), 0),
        seen_e((size_t)std::max<int64_t>(ne_new - ne0, 0), 0);
    for (int64_t i = 0; i < n_vert; ++i) { const int64_t v = vert_idx_host[i]; if (v < 0 || v >= nv_new) ++bad; else if (v >= nv0) seen_v[(size_t)(v - nv0)] = 1; }
    for (int64_t i = 0; i < n_cell; ++i) {
      const int64_t c = cell_idx_host[i];
      if (c < 0 || c >= nc_new) ++bad; else if (c >= nc0) seen_c[(size_t)(c - nc0)] = 1;
      unsigned seen = 0;
      for (int j = 0; j < k; ++j) { bad += (cell_v_host[k * i + j] < 0 || cell_v_host[k * i + j] >= nv_new); seen |= 1u << ((cell_perm_host[i] >> (2 * j)) & 3u); }
      bad += (seen != (1u << k) - 1u);
    }
    for (int64_t i = 0; i < n_edge; ++i) {
      const int64_t e = edge_idx_host[i];
      if (e < 0 || e >= ne_new) ++bad; else if (e >= ne0) seen_e[(size_t)(e - ne0)] = 1;
      const int32_t* q = edge_q_host + 4 * i;
      bad += (q[0] < 0 || q[0] >= nv_new || q[1] < 0 || q[1] >= nv_new || q[2] < 0 || q[2] >= nv_new || q[3] < -1 || q[3] >= nv_new);
    }
    for (uint8_t f : seen_v) bad += !f;
    for (uint8_t f : seen_c) bad += !f;
    for (uint8_t f : seen_e) bad += !f;
    if (facets_host) for (int64_t i = 0; i < ns_new * dim; ++i) bad += (facets_host[i] < 0 || facets_host[i] >= nv_new);
    REQUIRE(bad == 0, MMB_ERR_INVALID_ARG, "mmb_patch_mesh: index out of range, invalid corner permutation, or a new index without data");
    REQUIRE(facets_host || ctx->ns == 0 || nv_new >= nv0, MMB_ERR_INVALID_ARG, "mmb_patch_mesh: shrinking the vertex set needs the facet list (its indices may be stale)");
  }
  cudaStream_t S = ctx->stream;
  const int64_t nc_pad_new = (nc_new + 3) / 4 * 4;
  // resident data that must survive: grow in place
  CK(ctx->x.grow((size_t)nv_new * dim + 2, (size_t)nv0 * dim, S)); CK(ctx->seg_hint.grow((size_t)nv_new, (size_t)nv0, S));
  CK(ctx->v_if.grow((size_t)nv_new + 1, (size_t)nv0, S));
  for (int j = 0; j < k; ++j) CK(ctx->cv[j].grow((size_t)nc_pad_new + 4, (size_t)nc0, S));
  CK(ctx->cells_aos.grow((size_t)nc_new * k + 4, (size_t)nc0 * k, S)); CK(ctx->perm.grow((size_t)nc_pad_new + 4, (size_t)nc0, S));
  if (dim == 2) CK(ctx->edges.grow((size_t)ne_new + 1, (size_t)ne0, S));
  // outputs / scratch: contents need not survive
  CK(ctx->shift.ensure((size_t)nv_new * dim + 2)); CK(ctx->dist.ensure((size_t)nv_new));
  CK(ctx->valid_fp.ensure(nc_pad_new + 4)); CK(ctx->orient.ensure(nc_pad_new + 4)); CK(ctx->mark.ensure(nc_pad_new + 4)); CK(ctx->longest.ensure(nc_pad_new + 4));
  CK(ctx->wl.ensure(std::max(std::max(nc_pad_new, ne_new), nv_new) + 4)); CK(ctx->list_out.ensure(std::max(nc_pad_new, ne_new) + 4));
  CK(ctx->block_count.ensure((std::max(nc_pad_new, ne_new) + COMPACT_TILE - 1) / COMPACT_TILE + 1));
  CK(ctx->cand_flag.ensure((size_t)std::max(nc_pad_new, std::max(ns_new, ctx->ns)) + 4));
  if (dim == 2) CK(ctx->edge_flag.ensure(ne_new + 4));
  if (nv_new > nv0) {
    CK(cudaMemsetAsync(ctx->seg_hint.p + nv0, 0xFF, (size_t)(nv_new - nv0) * sizeof(int32_t), S));
    CK(cudaMemsetAsync(ctx->v_if.p + nv0, 0, (size_t)(nv_new - nv0), S));
  }
  // scatter the patch (staged through temporary device buffers)
  DBuf<int32_t> d_idx, d_val; DBuf<double> d_x; DBuf<uint8_t> d_perm;
  auto release = [&]() { d_idx.release(); d_val.release(); d_x.release(); d_perm.release(); };
  auto fail = [&](cudaError_t e) { ctx->err = cudaGetErrorString(e); release(); return (int)MMB_ERR_CUDA; };
  cudaError_t e;
  if (n_vert) {
    if ((e = d_idx.ensure(n_vert)) || (e = d_x.ensure((size_t)n_vert * dim))) return fail(e);
    cudaMemcpyAsync(d_idx.p, vert_idx_host, (size_t)n_vert * 4, cudaMemcpyHostToDevice, S);
    cudaMemcpyAsync(d_x.p, vert_x_host, (size_t)n_vert * dim * 8, cudaMemcpyHostToDevice, S);
    LAUNCH("patch_vertices", k_patch_vertices, grid_for(n_vert, 256), 256, ctx->x.p, ctx->seg_hint.p, dim, d_idx.p, d_x.p, n_vert);
    if ((e = cudaStreamSynchronize(S)) != cudaSuccess) return fail(e);
  }
  if (n_cell) {
    if ((e = d_idx.ensure(n_cell)) || (e = d_val.ensure((size_t)n_cell * k)) || (e = d_perm.ensure(n_cell))) return fail(e);
    cudaMemcpyAsync(d_idx.p, cell_idx_host, (size_t)n_cell * 4, cudaMemcpyHostToDevice, S);
    cudaMemcpyAsync(d_val.p, cell_v_host, (size_t)n_cell * k * 4, cudaMemcpyHostToDevice, S);
    cudaMemcpyAsync(d_perm.p, cell_perm_host, (size_t)n_cell, cudaMemcpyHostToDevice, S);
    LAUNCH("patch_cells", k_patch_cells, grid_for(n_cell, 256), 256, ctx->cells_aos.p, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, k == 4 ? ctx->cv[3].p : nullptr,
           ctx->perm.p, k, d_idx.p, d_val.p, d_perm.p, n_cell);
    if ((e = cudaStreamSynchronize(S)) != cudaSuccess) return fail(e);
  }
  for (int j = 0; j < k; ++j)                                // padding cells reference vertex 0 (4 cells per thread in k_validate2d)
    CK(cudaMemsetAsync(ctx->cv[j].p + nc_new, 0, (size_t)(nc_pad_new + 4 - nc_new) * sizeof(int32_t), S));
  if (n_edge) {
    if ((e = d_idx.ensure(n_edge)) || (e = d_val.ensure((size_t)n_edge * 4))) return fail(e);
    cudaMemcpyAsync(d_idx.p, edge_idx_host, (size_t)n_edge * 4, cudaMemcpyHostToDevice, S);
    cudaMemcpyAsync(d_val.p, edge_q_host, (size_t)n_edge * 16, cudaMemcpyHostToDevice, S);
    LAUNCH("patch_edges", k_patch_edges, grid_for(n_edge, 256), 256, ctx->edges.p, d_idx.p, (const int4*)d_val.p, n_edge);
    if ((e = cudaStreamSynchronize(S)) != cudaSuccess) return fail(e);
  }
  release();
  ctx->nv = nv_new; ctx->nc = nc_new; ctx->ne = ne_new; ctx->nc_pad = nc_pad_new;
  ctx->own_lo = 0; ctx->own_hi = nc_new;
  if (dim == 2) {                                              // Hilbert tile windows of the new tables
    const int64_t ntc = (nc_pad_new + TILE - 1) / TILE, nte = (ne_new + TILE - 1) / TILE;
    CK(ctx->tile_lo_cells.ensure(ntc + 1)); CK(ctx->tile_lo_edges.ensure(nte + 1));
    CK(cudaMemsetAsync(ctx->tile_lo_cells.p, 0, (ntc + 1) * sizeof(int32_t), S));
    CK(cudaMemsetAsync(ctx->tile_lo_edges.p, 0, (nte + 1) * sizeof(int32_t), S));
    if (nc_new) LAUNCH("tile_windows", k_tile_windows, (unsigned)ntc, 256, ctx->cv[0].p, 1, nc_new, nv_new, ctx->tile_lo_cells.p);
    if (ne_new) LAUNCH("tile_windows", k_tile_windows, (unsigned)nte, 256, (const int32_t*)ctx->edges.p, 4, ne_new, nv_new, ctx->tile_lo_edges.p);
  }
  if (facets_host || (ns_new == 0 && ctx->ns != 0)) {          // the interface changed: rebuild what hangs on it (hints start cold)
    std::vector<double> xh((size_t)nv_new * dim);
    CK(cudaMemcpyAsync(xh.data(), ctx->x.p, xh.size() * sizeof(double), cudaMemcpyDeviceToHost, S));
    CK(cudaStreamSynchronize(S));
    TRY(setup_facets(ctx, ns_new, facets_host, xh.data()));
  } else {
    REQUIRE(ns_new == ctx->ns, MMB_ERR_INVALID_ARG, "mmb_patch_mesh: ns changed but no facet list given");
    if (n_vert) ctx->hints_cold = true;                        // patched vertices lost their hints: they take a seed-grid facet
  }
  CK(cudaStreamSynchronize(S));
  CK(cudaGetLastError());
  ctx->has_shift = false; ctx->dist_valid = false; ctx->have_nb = false; ctx->have_levels = false; ctx->cand_ready = false;
  ctx->nnew = ctx->npairs = 0; ctx->graph_valid = false; ctx->sharded_steps = 0; ctx->n_halo_send = ctx->n_halo_recv = 0;
  return MMB_OK;
}
int mmb_upload_shifts(mmb_context* ctx, const double* shifts_host) {
  if (!ctx || !shifts_host) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  CK(cudaMemcpyAsync(ctx->shift.p, shifts_host, (size_t)ctx->nv * ctx->dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  ctx->has_shift = true;
  return MMB_OK;
}
int mmb_scale_shifts(mmb_context* ctx, double a) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->has_shift, MMB_ERR_STATE, "no resident shifts");
  const int64_t n2 = ((int64_t)ctx->nv * ctx->dim + 1) / 2;
  LAUNCH("scale_shifts", k_scale, std::min<unsigned>(grid_for(n2, 256), ctx->num_sms * 16), 256, (double2*)ctx->shift.p, a, n2);
  CK(cudaGetLastError());
  return MMB_OK;
}
static int stage_shifts(mmb_context* ctx, const double* shifts_host) {
  if (shifts_host) return mmb_upload_shifts(ctx, shifts_host);
  REQUIRE(ctx->has_shift, MMB_ERR_STATE, "shifts == NULL but no resident shifts (mmb_upload_shifts)");
  return MMB_OK;
}

// ---- move ---------------------------------------------------------------------------------------------
// gated = the move of the fused step under move_policy 0: withheld on the device when the trial validity pass of this
// step found an inverted cell (counters.n_invalid_global)
static int do_move(mmb_context* ctx, bool gated = false) {
  const int64_t n2 = ((int64_t)ctx->nv * ctx->dim + 1) / 2;
  LAUNCH("move", k_move, std::min<unsigned>(grid_for(n2, 256), ctx->num_sms * 16), 256, (double2*)ctx->x.p,
         (const double2*)ctx->shift.p, n2, gated && ctx->prm.move_policy == 0 ? ctx->counters.p : (const Counters*)nullptr);
  ctx->dist_valid = false;
  return MMB_OK;
}
int mmb_move_vertices(mmb_context* ctx, const double* shifts_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  TRY(stage_shifts(ctx, shifts_host));
  TRY(do_move(ctx));
  CK(cudaGetLastError());
  return MMB_OK;
}

// ---- trial move + validity ------------------------------------------------------------------------------
static int do_validate(mmb_context* ctx, bool with_shift) {
  Counters* c = ctx->counters.p;
  const int publish = ctx->sharded ? 0 : 1;      // sharded: the caller all-reduces n_invalid into n_invalid_global
  const unsigned fb = ctx->num_sms * 4;
  if (ctx->dim == 2) {
    const int64_t nc4 = ctx->nc_pad / 4;
    if (with_shift)
      LAUNCH("validate2d", k_validate2d<true>, grid_for(nc4, 256), 256, (const double2*)ctx->x.p, (const double2*)ctx->shift.p,
             (const int4*)ctx->cv[0].p, (const int4*)ctx->cv[1].p, (const int4*)ctx->cv[2].p, ctx->tile_lo_cells.p, ctx->nv, nc4, ctx->nc,
             ctx->own_lo, ctx->own_hi, (uint32_t*)ctx->valid_fp.p, (uint32_t*)ctx->orient.p, ctx->wl.p, c);
    else
      LAUNCH("validate2d", k_validate2d<false>, grid_for(nc4, 256), 256, (const double2*)ctx->x.p, (const double2*)ctx->shift.p,
             (const int4*)ctx->cv[0].p, (const int4*)ctx->cv[1].p, (const int4*)ctx->cv[2].p, ctx->tile_lo_cells.p, ctx->nv, nc4, ctx->nc,
             ctx->own_lo, ctx->own_hi, (uint32_t*)ctx->valid_fp.p, (uint32_t*)ctx->orient.p, ctx->wl.p, c);
    if (with_shift)
      LAUNCH("exact_orient2d", k_exact_orient2d<true>, fb, 128, (const double2*)ctx->x.p, (const double2*)ctx->shift.p,
             ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, ctx->wl.p, ctx->own_lo, ctx->own_hi, ctx->orient.p, c, publish);
    else
      LAUNCH("exact_orient2d", k_exact_orient2d<false>, fb, 128, (const double2*)ctx->x.p, (const double2*)ctx->shift.p,
             ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, ctx->wl.p, ctx->own_lo, ctx->own_hi, ctx->orient.p, c, publish);
  } else {
    if (with_shift)
      LAUNCH("validate3d", k_validate3d<true>, grid_for(ctx->nc, 256), 256, ctx->x.p, ctx->shift.p, ctx->cv[0].p, ctx->cv[1].p,
             ctx->cv[2].p, ctx->cv[3].p, ctx->nc, ctx->valid_fp.p, ctx->orient.p, ctx->wl.p, c);
    else
      LAUNCH("validate3d", k_validate3d<false>, grid_for(ctx->nc, 256), 256, ctx->x.p, ctx->shift.p, ctx->cv[0].p, ctx->cv[1].p,
             ctx->cv[2].p, ctx->cv[3].p, ctx->nc, ctx->valid_fp.p, ctx->orient.p, ctx->wl.p, c);
    if (with_shift)
      LAUNCH("exact_orient3d", k_exact_orient3d<true>, fb, 128, ctx->x.p, ctx->shift.p, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p,
             ctx->cv[3].p, ctx->wl.p, ctx->orient.p, c, publish);
    else
      LAUNCH("exact_orient3d", k_exact_orient3d<false>, fb, 128, ctx->x.p, ctx->shift.p, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p,
             ctx->cv[3].p, ctx->wl.p, ctx->orient.p, c, publish);
  }
  return MMB_OK;
}
static int zero_counters(mmb_context* ctx) {
  CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(Counters), ctx->stream));
  return MMB_OK;
}
int mmb_trial_move_validate(mmb_context* ctx, const double* shifts_host, uint8_t* valid_fp_host, int8_t* orient_sign_host,
                            int64_t* n_invalid) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  TRY(stage_shifts(ctx, shifts_host));
  // keep max_dist across calls: only the validity counters are reset
  CK(cudaMemsetAsync(&ctx->counters.p->n_invalid, 0, 2 * sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(&ctx->counters.p->wl_orient, 0, sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(&ctx->counters.p->n_exact_orient, 0, sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(&ctx->counters.p->pad_[1], 0, sizeof(unsigned long long), ctx->stream));
  TRY(do_validate(ctx, true));
  if (valid_fp_host) CK(cudaMemcpyAsync(valid_fp_host, ctx->valid_fp.p, (size_t)ctx->nc, cudaMemcpyDeviceToHost, ctx->stream));
  if (orient_sign_host) CK(cudaMemcpyAsync(orient_sign_host, ctx->orient.p, (size_t)ctx->nc, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaGetLastError());
  if (n_invalid || valid_fp_host || orient_sign_host || ctx->dim == 3) {
    TRY(fetch_counters(ctx));
    if (n_invalid) *n_invalid = (int64_t)ctx->h_counters->n_invalid;
    TRY(check_predicate_domain(ctx));
    if (ctx->dim == 3 && ctx->h_counters->n_invalid > 0) {
      ctx->err = "Vertices could not be moved, because the removal of vertices is not supported in 3D!";
      return MMB_ERR_INVALID_CELL_3D;
    }
  }
  return MMB_OK;
}

// ordered compaction helper -> ctx->list_out; total in *n_total
static int compact_flags(mmb_context* ctx, const uint8_t* flags, int64_t n, uint8_t match, int32_t* out_host, int64_t cap,
                         int64_t* n_total) {
  const int nb = (int)((n + COMPACT_TILE - 1) / COMPACT_TILE);
  if (nb == 0) { if (n_total) *n_total = 0; return MMB_OK; }
  LAUNCH("compact_count", k_compact_count, nb, 256, flags, n, match, ctx->block_count.p);
  LAUNCH("compact_scan", k_compact_scan, 1, 1024, ctx->block_count.p, nb, &ctx->counters.p->pad_[0]);
  LAUNCH("compact_scatter", k_compact_scatter, nb, 256, flags, n, match, ctx->block_count.p, ctx->list_out.p,
         (int64_t)ctx->list_out.cap);
  CK(cudaGetLastError());
  TRY(fetch_counters(ctx));
  const int64_t tot = (int64_t)ctx->h_counters->pad_[0];
  if (n_total) *n_total = tot;
  if (out_host && cap > 0 && tot > 0) {
    CK(cudaMemcpyAsync(out_host, ctx->list_out.p, (size_t)std::min(cap, tot) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return MMB_OK;
}
int mmb_invalid_cells(mmb_context* ctx, int32_t* cells_host, int64_t cap, int64_t* n_total) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  return compact_flags(ctx, ctx->valid_fp.p, ctx->nc, 0, cells_host, cap, n_total);
}
int mmb_marked_cells(mmb_context* ctx, int which, int32_t* cells_host, int64_t cap, int64_t* n_total) {
  if (!ctx || (which != 1 && which != -1)) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  return compact_flags(ctx, (const uint8_t*)ctx->mark.p, ctx->nc, which == 1 ? 1 : 0xFF, cells_host, cap, n_total);
}

// ---- legality -------------------------------------------------------------------------------------------
// shifted = true evaluates the flags on x (+) shift, the coordinates moveVertices is about to write (same IEEE additions,
// so the flags are bit-identical to running after the move); the pipelined host step uses it to get the flags out early
static int do_legality(mmb_context* ctx, bool shifted = false) {
  if (ctx->ne == 0) return MMB_OK;
  const int gated = ctx->prm.move_policy == 0 ? 1 : 0;    // a withheld move leaves x in place: flags on x then
  if (shifted) {
    LAUNCH("legality", k_legality<true>, grid_for(ctx->ne, TILE), 256, (const double2*)ctx->x.p, (const double2*)ctx->shift.p,
           (const int4*)ctx->edges.p, ctx->ne, ctx->tile_lo_edges.p, ctx->nv, ctx->edge_flag.p, ctx->wl.p, ctx->counters.p, gated);
    LAUNCH("exact_incircle", k_exact_incircle<true>, ctx->num_sms * 4, 128, (const double2*)ctx->x.p,
           (const double2*)ctx->shift.p, (const int4*)ctx->edges.p, ctx->wl.p, ctx->edge_flag.p, ctx->counters.p, gated);
    return MMB_OK;
  }
  LAUNCH("legality", k_legality<false>, grid_for(ctx->ne, TILE), 256, (const double2*)ctx->x.p, (const double2*)ctx->shift.p,
         (const int4*)ctx->edges.p, ctx->ne, ctx->tile_lo_edges.p, ctx->nv, ctx->edge_flag.p, ctx->wl.p, ctx->counters.p, 0);
  LAUNCH("exact_incircle", k_exact_incircle<false>, ctx->num_sms * 4, 128, (const double2*)ctx->x.p,
         (const double2*)ctx->shift.p, (const int4*)ctx->edges.p, ctx->wl.p, ctx->edge_flag.p, ctx->counters.p, 0);
  return MMB_OK;
}
int mmb_legality(mmb_context* ctx, uint8_t* flags_host, int64_t* n_illegal) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->dim == 2, MMB_ERR_STATE, "legality needs an uploaded 2-D mesh");
  CK(cudaMemsetAsync(&ctx->counters.p->n_illegal, 0, sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(&ctx->counters.p->wl_incircle, 0, sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(&ctx->counters.p->n_exact_incircle, 0, sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(&ctx->counters.p->pad_[1], 0, sizeof(unsigned long long), ctx->stream));
  TRY(do_legality(ctx));
  if (flags_host && ctx->ne) CK(cudaMemcpyAsync(flags_host, ctx->edge_flag.p, (size_t)ctx->ne, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaGetLastError());
  if (flags_host || n_illegal) { TRY(fetch_counters(ctx)); if (n_illegal) *n_illegal = (int64_t)ctx->h_counters->n_illegal; TRY(check_predicate_domain(ctx)); }
  return MMB_OK;
}

// ---- distance -------------------------------------------------------------------------------------------
static int do_distance(mmb_context* ctx, bool reset_max = true) {
  if (reset_max) CK(cudaMemsetAsync(&ctx->counters.p->max_dist_bits, 0, sizeof(unsigned long long), ctx->stream));
#ifdef MMB_DEV_VARIANTS
  CK(cudaMemsetAsync(ctx->counters.p->dbg, 0, sizeof(ctx->counters.p->dbg), ctx->stream));
#endif
  const int ns = (int)ctx->ns, L = ctx->L;
  if (ctx->dim == 2) {
    if (ns) {
      LAUNCH("bvh_build2d", k_bvh_build2d, L < 256 ? 1 : L / 256, 256, (const double2*)ctx->x.p, (const int2*)ctx->facets_sorted.p,
             ns, L, ctx->seg_leaf.p, ctx->box2.p, ctx->ticket.p, ctx->counters.p);
    }
    int logL = 0; while ((1 << logL) < L) ++logL;
    if (ctx->hints_cold && ns) {                               // first update after an upload: no hints yet
      CK(ctx->seed_grid.ensure(SEED_GX * SEED_GY));
      LAUNCH("distance_seeds", k_distance_seeds, grid_for(SEED_GX * SEED_GY, 128), 128, ctx->seg_leaf.p, ctx->box2.p, ns, L, ctx->seed_grid.p);
    }
    LAUNCH("distance2d", k_distance2d, grid_for(ctx->nv, 256), 256, (const double2*)ctx->x.p, ctx->nv, ctx->seg_leaf.p,
           ctx->box2.p, ns, L, logL, ctx->dist.p, ctx->seg_hint.p, ctx->hints_cold ? (const int32_t*)ctx->seed_grid.p : (const int32_t*)nullptr,
           ctx->wl.p, ctx->counters.p);
    if (ns && logL > DIST_GROUP_LOG)      // warps whose candidate list overflowed: one warp per vertex (usually a handful of warps)
      LAUNCH("distance2d_wide", k_distance2d_wide, ctx->num_sms * 2, 256, (const double2*)ctx->x.p, ctx->seg_leaf.p, ctx->box2.p, ns, L, ctx->dist.p,
             ctx->seg_hint.p, ctx->hints_cold ? (const int32_t*)ctx->seed_grid.p : (const int32_t*)nullptr, ctx->wl.p, ctx->counters.p);
    ctx->hints_cold = false;
  } else {
    if (ns) {
      LAUNCH("bvh_build3d", k_bvh_build3d, L < 256 ? 1 : L / 256, 256, ctx->x.p, ctx->facets_sorted.p, ns, L, ctx->tri_leaf.p,
             ctx->box3.p, ctx->ticket.p);
    }
    LAUNCH("distance3d", k_distance3d, grid_for(ctx->nv, 256), 256, ctx->x.p, ctx->nv, ctx->tri_leaf.p, ctx->box3.p, ns, L,
           ctx->dist.p, ctx->seg_hint.p, ctx->counters.p);
  }
  ctx->dist_valid = true;
  return MMB_OK;
}
int mmb_distance_update(mmb_context* ctx, double* dist_host, double* max_out) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  TRY(do_distance(ctx));
  if (dist_host) CK(cudaMemcpyAsync(dist_host, ctx->dist.p, (size_t)ctx->nv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaGetLastError());
  if (dist_host || max_out) {
    TRY(fetch_counters(ctx));
    if (max_out) { double m; memcpy(&m, &ctx->h_counters->max_dist_bits, 8); *max_out = m; }
  }
  return MMB_OK;
}
int mmb_distance_set(mmb_context* ctx, const double* dist_host) {
  if (!ctx || !dist_host) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  CK(cudaMemcpyAsync(ctx->dist.p, dist_host, (size_t)ctx->nv * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  double m = 0.0;
  for (int64_t i = 0; i < ctx->nv; ++i) m = std::max(dist_host[i], m);
  CK(cudaMemcpyAsync(&ctx->counters.p->max_dist_bits, &m, 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->dist_valid = true;
  return MMB_OK;
}

// ---- indicator ------------------------------------------------------------------------------------------
int mmb_indicator_init(mmb_context* ctx) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "indicator_init needs an uploaded mesh");
  Counters* c = ctx->counters.p;
  const unsigned long long init[4] = { ~0ull, 0ull, ~0ull, 0ull };
  CK(cudaMemcpyAsync(c->minmax_bits, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->dim == 2) {
    if (ctx->ns)
      LAUNCH("edge_minmax", k_edge_minmax, grid_for(ctx->ns, 256), 256, (const double2*)ctx->x.p, ctx->facets.p, 2, ctx->ns,
             &c->minmax_bits[0], &c->minmax_bits[1]);
    if (ctx->ne)
      LAUNCH("edge_minmax", k_edge_minmax, grid_for(ctx->ne, 256), 256, (const double2*)ctx->x.p, (const int32_t*)ctx->edges.p, 4,
             ctx->ne, &c->minmax_bits[2], &c->minmax_bits[3]);
  } else {                                                  // ratioindicator.hh:62-91 is dimension generic
    if (ctx->ns)
      LAUNCH("edge_minmax3d", k_edge_minmax3d, grid_for(ctx->ns, 256), 256, ctx->x.p, ctx->facets.p, 3, ctx->ns, &c->minmax_bits[0], &c->minmax_bits[1]);
    if (ctx->nc)
      LAUNCH("edge_minmax3d", k_edge_minmax3d, grid_for(ctx->nc, 256), 256, ctx->x.p, ctx->cells_aos.p, 4, ctx->nc, &c->minmax_bits[2], &c->minmax_bits[3]);
  }
  CK(cudaGetLastError());
  TRY(fetch_counters(ctx));
  auto as_d = [](unsigned long long b) { double d; memcpy(&d, &b, 8); return d; };
  const double K = 2.0, k = K / (2.0 * 4.0);
  double maxH = 0.0, minH = 1e100, maxh = 0.0, minh = 1e100;
  if (ctx->ns) { minH = std::min(minH, as_d(ctx->h_counters->minmax_bits[0])); maxH = std::max(maxH, as_d(ctx->h_counters->minmax_bits[1])); }
  if (ctx->dim == 2 ? ctx->ne > 0 : ctx->nc > 0) { minh = std::min(minh, as_d(ctx->h_counters->minmax_bits[2])); maxh = std::max(maxh, as_d(ctx->h_counters->minmax_bits[3])); }
  maxH = K * maxH; minH = k * minH;
  if (ctx->ns == 0) { maxH = K * maxh; minH = k * minh; }
  ctx->prm.maxH = maxH; ctx->prm.minH = minH; ctx->prm.factor = maxh / minh;
  return MMB_OK;
}
static int do_mark(mmb_context* ctx) {
  const DevParams dp = dev_params(ctx->prm);
  if (ctx->dim == 2)
    LAUNCH("mark2d", k_mark2d, grid_for(ctx->nc, 256), 256, (const double2*)ctx->x.p, ctx->dist.p, ctx->cv[0].p, ctx->cv[1].p,
           ctx->cv[2].p, ctx->perm.p, ctx->nc, ctx->own_lo, ctx->own_hi, dp, ctx->counters.p, ctx->mark.p, ctx->longest.p, ctx->counters.p);
  else
    LAUNCH("mark3d", k_mark3d, grid_for(ctx->nc, 256), 256, ctx->x.p, ctx->dist.p, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p,
           ctx->cv[3].p, ctx->perm.p, ctx->nc, dp, ctx->counters.p, ctx->mark.p, ctx->longest.p, ctx->counters.p);
  return MMB_OK;
}
int mmb_mark_elements(mmb_context* ctx, int run_distance, int8_t* marks_host, uint8_t* longest_edge_host, int64_t counts[2]) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  if (run_distance) TRY(do_distance(ctx));
  REQUIRE(ctx->dist_valid, MMB_ERR_STATE, "markElements before Distance::update");
  CK(cudaMemsetAsync(&ctx->counters.p->n_refine, 0, 2 * sizeof(unsigned long long), ctx->stream));
  TRY(do_mark(ctx));
  if (marks_host) CK(cudaMemcpyAsync(marks_host, ctx->mark.p, (size_t)ctx->nc, cudaMemcpyDeviceToHost, ctx->stream));
  if (longest_edge_host) CK(cudaMemcpyAsync(longest_edge_host, ctx->longest.p, (size_t)ctx->nc, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaGetLastError());
  if (marks_host || longest_edge_host || counts) {
    TRY(fetch_counters(ctx));
    if (counts) { counts[0] = (int64_t)ctx->h_counters->n_refine; counts[1] = (int64_t)ctx->h_counters->n_coarsen; }
  }
  return MMB_OK;
}
int mmb_mark_interface(mmb_context* ctx, int8_t* marks_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "mark_interface needs an uploaded mesh");
  REQUIRE(ctx->dist_valid, MMB_ERR_STATE, "mark_interface before Distance::update");
  if (ctx->ns == 0) return MMB_OK;
  if (ctx->dim == 2)
    LAUNCH("mark_interface2d", k_mark_interface2d, grid_for(ctx->ns, 256), 256, (const double2*)ctx->x.p, (const int2*)ctx->facets.p,
           (int)ctx->ns, dev_params(ctx->prm), ctx->counters.p, ctx->imark.p);
  else
    LAUNCH("mark_interface3d", k_mark_interface3d, grid_for(ctx->ns, 256), 256, ctx->x.p, ctx->facets.p, (int)ctx->ns, dev_params(ctx->prm),
           ctx->counters.p, ctx->imark.p);
  CK(cudaGetLastError());
  if (marks_host) {
    CK(cudaMemcpyAsync(marks_host, ctx->imark.p, (size_t)ctx->ns, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return MMB_OK;
}

// ---- projection -----------------------------------------------------------------------------------------
int mmb_upload_pairs(mmb_context* ctx, int64_t nnew, const int64_t* new_off_host, const int32_t* new_cell_host, int64_t npairs,
                     const double* old_xy_host, const int32_t* old_idx_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->dim == 2, MMB_ERR_STATE, "projection needs an uploaded 2-D mesh");
  REQUIRE(nnew >= 0 && npairs >= 0 && (nnew == 0 || (new_off_host && new_cell_host)) && (npairs == 0 || (old_xy_host && old_idx_host)),
          MMB_ERR_INVALID_ARG, "mmb_upload_pairs: null arrays");
  {
    int64_t bad = 0;
    if (nnew) bad += (new_off_host[0] != 0 || new_off_host[nnew] != npairs);
    for (int64_t i = 0; i < nnew; ++i) bad += (new_off_host[i + 1] < new_off_host[i] || new_cell_host[i] < 0 || new_cell_host[i] >= ctx->nc);
    for (int64_t p = 0; p < npairs; ++p) bad += (old_idx_host[p] < 0);
    REQUIRE(bad == 0, MMB_ERR_INVALID_ARG, "mmb_upload_pairs: malformed CSR (offsets, new cell or old index out of range)");
  }
  CK(ctx->new_off.ensure(nnew + 1)); CK(ctx->new_cell.ensure(nnew + 1)); CK(ctx->old_idx.ensure(npairs + 1));
  CK(ctx->old_xy.ensure(6 * (size_t)npairs + 1)); CK(ctx->pair_vol.ensure(npairs + 1));
  const int nb = (int)grid_for(4 * nnew, PROJ_THREADS) + 64;   // + slack for chunked launches
  CK(ctx->part_a.ensure(nb)); CK(ctx->part_b.ensure(nb));
  if (nnew) {
    CK(cudaMemcpyAsync(ctx->new_off.p, new_off_host, (size_t)(nnew + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->new_cell.p, new_cell_host, (size_t)nnew * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (npairs) {
    CK(cudaMemcpyAsync(ctx->old_idx.p, old_idx_host, (size_t)npairs * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->old_xy.p, old_xy_host, 6 * (size_t)npairs * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->nnew = nnew; ctx->npairs = npairs; ctx->graph_valid = false;
  ctx->max_old_idx = -1;
  for (int64_t p = 0; p < npairs; ++p) ctx->max_old_idx = std::max<int64_t>(ctx->max_old_idx, old_idx_host[p]);
  // chunk table for the pipelined host-buffer step: new cells are listed in ascending cell order (element iteration
  // order, mmesh.hh:1143), so chunk k of the CSR writes DOFs of cells [cell_lo[k], cell_hi[k]]
  ctx->chunk_i.clear(); ctx->chunk_cell_lo.clear(); ctx->chunk_cell_hi.clear();
  const int C = nnew >= (1 << 16) ? pipeline_chunks() : 1;
  bool ascending = true;
  for (int64_t i = 1; i < nnew && ascending; ++i) ascending = new_cell_host[i] > new_cell_host[i - 1];
  ctx->chunks_ok = ascending && nnew > 0;
  for (int k = 0; k <= C; ++k) ctx->chunk_i.push_back(nnew * k / C);
  for (int k = 0; k < C && nnew > 0; ++k) {
    ctx->chunk_cell_lo.push_back(new_cell_host[ctx->chunk_i[k]]);
    ctx->chunk_cell_hi.push_back(new_cell_host[ctx->chunk_i[k + 1] - 1]);
  }
  // Can the old DOFs be uploaded range by range behind the projection?  Split the old index space into C equal ranges;
  // chunk k's home range is range k.  Every reference outside the home range goes to a (unique, small) remote list that
  // is uploaded first.  If the numbering of old and new cells is unrelated the remote list is large -> monolithic upload.
  ctx->dof_pipe_ok = false; ctx->dof_lo.clear(); ctx->remote_idx.clear();
  if (ctx->chunks_ok && C > 1 && npairs > 0) {
    int64_t n_idx = 0;
    for (int64_t p = 0; p < npairs; ++p) n_idx = std::max<int64_t>(n_idx, old_idx_host[p] + 1);
    for (int k = 0; k <= C; ++k) ctx->dof_lo.push_back(n_idx * k / C);
    std::vector<uint8_t> seen((size_t)n_idx, 0);
    bool ok = true;
    for (int k = 0; k < C && ok; ++k) {
      const int64_t lo = ctx->dof_lo[k], hi = ctx->dof_lo[k + 1];
      for (int64_t p = new_off_host[ctx->chunk_i[k]]; p < new_off_host[ctx->chunk_i[k + 1]]; ++p) {
        const int32_t o = old_idx_host[p];
        if ((o < lo || o >= hi) && !seen[(size_t)o]) { seen[(size_t)o] = 1; ctx->remote_idx.push_back(o); }
      }
      ok = (int64_t)ctx->remote_idx.size() * 16 <= n_idx;
    }
    if (ok) {
      CK(ctx->d_remote_idx.ensure(ctx->remote_idx.size() + 1));
      if (!ctx->remote_idx.empty())
        CK(cudaMemcpy(ctx->d_remote_idx.p, ctx->remote_idx.data(), ctx->remote_idx.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      ctx->dof_pipe_ok = true;
    } else ctx->remote_idx.clear();
  }
  return MMB_OK;
}
#define PROJ_ARGS (const double2*)ctx->x.p, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, ctx->perm.p, ctx->new_off.p, ctx->new_cell.p, \
                  i0, i1, ctx->old_xy.p, ctx->old_idx.p, ctx->dofs_old.p, ctx->dofs_new.p, dp, ctx->part_a.p + pb0, ctx->part_b.p + pb0, ctx->counters.p
#ifdef MMB_DEV_VARIANTS
static int proj_variant() { const char* e = getenv("MMB_PROJ_VARIANT"); return e ? atoi(e) : 0; }
extern "C++" {
template <int NT, int MINB, bool FMA>
static int launch_project_tile(mmb_context* ctx, const ProjArgs& pa, int pb0, unsigned* nb_out) {
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_project_tile<NT, MINB, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProjTile<NT>)));
    attr_set = true;
  }
  const unsigned nb = grid_for(pa.nnew - pa.i_begin, NT / 4);
  {
    LaunchScope ls_(ctx, "project_p1");
    k_project_tile<NT, MINB, FMA><<<nb, NT, sizeof(ProjTile<NT>), ctx->stream>>>(pa, ctx->part_a.p + pb0, ctx->part_b.p + pb0, ctx->counters.p);
  }
  *nb_out = nb;
  return MMB_OK;
}
template <int NT, int MINB, bool FMA>
static int launch_project_pairs(mmb_context* ctx, const ProjArgs& pa, int pb0, unsigned* nb_out) {
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_project_pairs<NT, MINB, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProjTile<NT>)));
    attr_set = true;
  }
  CK(ctx->pair_partial.ensure(4 * (size_t)ctx->npairs + 4)); CK(ctx->pair_pieces.ensure((size_t)ctx->npairs + 4));
  const unsigned nbp = grid_for(pa.nnew - pa.i_begin, NT / 4);
  {
    LaunchScope ls_(ctx, "project_p1");
    k_project_pairs<NT, MINB, FMA><<<nbp, NT, sizeof(ProjTile<NT>), ctx->stream>>>(pa, (double4*)ctx->pair_partial.p, ctx->pair_pieces.p);
  }
  const unsigned nb = grid_for(pa.nnew - pa.i_begin, 128);
  LAUNCH("project_p1_cells", k_project_cells, nb, 128, pa, (const double4*)ctx->pair_partial.p, ctx->pair_pieces.p, ctx->part_a.p + pb0, ctx->part_b.p + pb0, ctx->counters.p);
  *nb_out = nb;
  return MMB_OK;
}
}  // extern "C++"
#endif
// projection of the new cells [i0, i1) of the pair CSR; per-block partial sums go to part_[pb0 ...); returns #blocks
static int project_range(mmb_context* ctx, int ncomp, int64_t i0, int64_t i1, int pb0, int* nblocks) {
  const DevParams dp = dev_params(ctx->prm);
  unsigned nb;
  if (ncomp == 1) {
    nb = grid_for(i1 - i0, PROJ_THREADS);
    LAUNCH("project_p0", (k_project<1, 1, 4, 1>), nb, PROJ_THREADS, PROJ_ARGS);
  } else {
    ProjArgs pa;
    pa.xy = (const double2*)ctx->x.p; pa.cv0 = ctx->cv[0].p; pa.cv1 = ctx->cv[1].p; pa.cv2 = ctx->cv[2].p; pa.perm = ctx->perm.p;
    pa.new_off = ctx->new_off.p; pa.new_cell = ctx->new_cell.p; pa.i_begin = i0; pa.nnew = i1;
    pa.old_xy = ctx->old_xy.p; pa.old_idx = ctx->old_idx.p; pa.dofs_old = ctx->dofs_old.p; pa.dofs_new = ctx->dofs_new.p; pa.prm = dp;
    // One thread per new cell walking its pairs in order (k_project).  Instrumented builds (-DMMB_DEV_VARIANTS) also carry
    // the alternatives that were measured slower on B200 (profiles/README.md): the tile kernel of project.cuh (pair slots,
    // transposed columns, dense list before the fan) and the hoisted-constant / merged-crossing / FMA forms of k_project.
#ifdef MMB_DEV_VARIANTS
    switch (proj_variant()) {
      case 10: TRY((launch_project_tile<128, 7, false>(ctx, pa, pb0, &nb))); break;
      case 11: TRY((launch_project_tile<128, 7, true>(ctx, pa, pb0, &nb))); break;
      case 12: TRY((launch_project_tile<256, 3, false>(ctx, pa, pb0, &nb))); break;
      case 13: TRY((launch_project_tile<256, 3, true>(ctx, pa, pb0, &nb))); break;
      case 30: TRY((launch_project_pairs<128, 7, false>(ctx, pa, pb0, &nb))); break;
      case 31: TRY((launch_project_pairs<128, 7, true>(ctx, pa, pb0, &nb))); break;
      case 32: TRY((launch_project_pairs<128, 6, false>(ctx, pa, pb0, &nb))); break;
      case 33: TRY((launch_project_pairs<128, 5, false>(ctx, pa, pb0, &nb))); break;
      case 34: TRY((launch_project_pairs<256, 3, false>(ctx, pa, pb0, &nb))); break;
      case 35: TRY((launch_project_pairs<64, 14, false>(ctx, pa, pb0, &nb))); break;
      case 21: nb = grid_for(i1 - i0, PROJ_THREADS); LAUNCH("project_p1", (k_project<3, 1, 5, 2>), nb, PROJ_THREADS, PROJ_ARGS); break;
      case 22: nb = grid_for(i1 - i0, PROJ_THREADS); LAUNCH("project_p1", (k_project<3, 1, 5, 3>), nb, PROJ_THREADS, PROJ_ARGS); break;
      default: nb = grid_for(i1 - i0, PROJ_THREADS); LAUNCH("project_p1", (k_project<3, 1, 5, 1>), nb, PROJ_THREADS, PROJ_ARGS); break;
    }
#else
    (void)pa;
    nb = grid_for(i1 - i0, PROJ_THREADS);
    LAUNCH("project_p1", (k_project<3, 1, 5, 1>), nb, PROJ_THREADS, PROJ_ARGS);
#endif
  }
  *nblocks = (int)nb;
  return MMB_OK;
}
static int reduce_partials(mmb_context* ctx, int n) {
  CK(ctx->red_slices.ensure(2 * REDUCE_MAX_BLOCKS));
  if (!ctx->red_ticket.p) { CK(ctx->red_ticket.ensure(1)); CK(cudaMemsetAsync(ctx->red_ticket.p, 0, sizeof(unsigned), ctx->stream)); }
  const int g = std::max(1, std::min(REDUCE_MAX_BLOCKS, (n + 2047) / 2048));
  LAUNCH("reduce_partials", k_reduce_partials, g, 256, ctx->part_a.p, ctx->part_b.p, n, ctx->red_slices.p, ctx->red_slices.p + REDUCE_MAX_BLOCKS,
         ctx->red_ticket.p, ctx->counters.p);
  return MMB_OK;
}
static int do_project(mmb_context* ctx, int ncomp) {
  if (ctx->nnew == 0) return MMB_OK;
  REQUIRE(ctx->max_old_idx < ctx->n_old, MMB_ERR_INVALID_ARG, "projection: a pair references an old cell beyond the DOF array");
  int nb = 0;
  TRY(project_range(ctx, ncomp, 0, ctx->nnew, 0, &nb));
  TRY(reduce_partials(ctx, nb));
  return MMB_OK;
}
static int stage_dofs(mmb_context* ctx, int ncomp, int64_t n_old, const double* dofs_old_host) {
  REQUIRE(ncomp == 1 || ncomp == 3, MMB_ERR_INVALID_ARG, "ncomp must be 1 (P0) or 3 (DG-P1)");
  if (dofs_old_host) {
    REQUIRE(n_old > 0, MMB_ERR_INVALID_ARG, "n_old must be positive");
    CK(ctx->dofs_old.ensure((size_t)n_old * ncomp));
    CK(cudaMemcpyAsync(ctx->dofs_old.p, dofs_old_host, (size_t)n_old * ncomp * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_old = n_old; ctx->dofs_ncomp = ncomp;
  } else REQUIRE(ctx->dofs_ncomp == ncomp, MMB_ERR_STATE, "dofs_old == NULL but no resident DOFs of this ncomp");
  if (ctx->dofs_new.cap < (size_t)ctx->nc * ncomp) {
    CK(ctx->dofs_new.ensure((size_t)ctx->nc * ncomp));
    CK(cudaMemsetAsync(ctx->dofs_new.p, 0, (size_t)ctx->nc * ncomp * sizeof(double), ctx->stream));
  }
  return MMB_OK;
}
int mmb_project(mmb_context* ctx, int ncomp, int64_t n_old, const double* dofs_old_host, double* dofs_new_host, double* mass_new,
                double* covered_area, int64_t* n_pieces) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->dim == 2, MMB_ERR_STATE, "projection needs an uploaded 2-D mesh");
  TRY(stage_dofs(ctx, ncomp, n_old, dofs_old_host));
  CK(cudaMemsetAsync(&ctx->counters.p->n_pieces, 0, sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(&ctx->counters.p->mass_new, 0, 2 * sizeof(double), ctx->stream));      // nnew == 0 leaves them at zero
  TRY(do_project(ctx, ncomp));
  if (dofs_new_host) CK(cudaMemcpyAsync(dofs_new_host, ctx->dofs_new.p, (size_t)ctx->nc * ncomp * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaGetLastError());
  if (dofs_new_host || mass_new || covered_area || n_pieces) {
    TRY(fetch_counters(ctx));
    if (mass_new) *mass_new = ctx->h_counters->mass_new;
    if (covered_area) *covered_area = ctx->h_counters->covered_area;
    if (n_pieces) *n_pieces = (int64_t)ctx->h_counters->n_pieces;
  }
  return MMB_OK;
}
int mmb_intersection_volumes(mmb_context* ctx, double* vol_host) {
  if (!ctx || !vol_host) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->dim == 2 && ctx->nnew > 0, MMB_ERR_STATE, "no pairs uploaded");
  LAUNCH("isect_volume", k_isect_volume, grid_for(ctx->nnew, 128), 128, (const double2*)ctx->x.p, ctx->cv[0].p, ctx->cv[1].p,
         ctx->cv[2].p, ctx->new_off.p, ctx->new_cell.p, ctx->nnew, ctx->old_xy.p, ctx->pair_vol.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(vol_host, ctx->pair_vol.p, (size_t)ctx->npairs * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MMB_OK;
}

int mmb_project_interface_p0(mmb_context* ctx, int64_t nnew, const double* new_len_host, const int64_t* new_off_host,
                             int64_t npairs, const double* old_len_host, const int32_t* old_idx_host, int64_t n_old,
                             const double* dofs_old_host, double* dofs_new_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(nnew > 0 && npairs >= 0 && n_old > 0 && new_len_host && new_off_host && dofs_old_host && dofs_new_host &&
          (npairs == 0 || (old_len_host && old_idx_host)), MMB_ERR_INVALID_ARG, "mmb_project_interface_p0: bad arguments");
  DBuf<double> nl, ol, dO, dN; DBuf<long long> off; DBuf<int32_t> oi;
  auto fail = [&](cudaError_t e) { ctx->err = cudaGetErrorString(e); nl.release(); ol.release(); dO.release(); dN.release(); off.release(); oi.release(); return MMB_ERR_CUDA; };
  cudaError_t e;
  if ((e = nl.ensure(nnew)) || (e = ol.ensure(npairs + 1)) || (e = dO.ensure(n_old)) || (e = dN.ensure(nnew)) || (e = off.ensure(nnew + 1)) || (e = oi.ensure(npairs + 1))) return fail(e);
  cudaMemcpyAsync(nl.p, new_len_host, nnew * 8, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(off.p, new_off_host, (nnew + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (npairs) { cudaMemcpyAsync(ol.p, old_len_host, npairs * 8, cudaMemcpyHostToDevice, ctx->stream); cudaMemcpyAsync(oi.p, old_idx_host, npairs * 4, cudaMemcpyHostToDevice, ctx->stream); }
  cudaMemcpyAsync(dO.p, dofs_old_host, n_old * 8, cudaMemcpyHostToDevice, ctx->stream);
  LAUNCH("project_interface_p0", k_project_interface_p0, grid_for(nnew, 256), 256, nl.p, off.p, nnew, ol.p, oi.p, dO.p, dN.p);
  cudaMemcpyAsync(dofs_new_host, dN.p, nnew * 8, cudaMemcpyDeviceToHost, ctx->stream);
  e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(e);
  nl.release(); ol.release(); dO.release(); dN.release(); off.release(); oi.release();
  return MMB_OK;
}
int mmb_project_interface_p1(mmb_context* ctx, int64_t nnew, const double* new_xy_host, const int64_t* new_off_host,
                             int64_t npairs, const double* old_xy_host, const int32_t* old_idx_host, int64_t n_old,
                             const double* dofs_old_host, double* dofs_new_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(nnew > 0 && npairs >= 0 && n_old > 0 && new_xy_host && new_off_host && dofs_old_host && dofs_new_host &&
          (npairs == 0 || (old_xy_host && old_idx_host)), MMB_ERR_INVALID_ARG, "mmb_project_interface_p1: bad arguments");
  REQUIRE(new_off_host[0] == 0 && new_off_host[nnew] == npairs, MMB_ERR_INVALID_ARG, "mmb_project_interface_p1: malformed CSR");
  for (int64_t i = 0; i < nnew; ++i) REQUIRE(new_off_host[i + 1] >= new_off_host[i], MMB_ERR_INVALID_ARG, "mmb_project_interface_p1: malformed CSR");
  for (int64_t p = 0; p < npairs; ++p) REQUIRE(old_idx_host[p] >= 0 && old_idx_host[p] < n_old, MMB_ERR_INVALID_ARG, "mmb_project_interface_p1: old index out of range");
  DBuf<double> nx, ox, dO, dN; DBuf<long long> off; DBuf<int32_t> oi;
  auto release = [&]() { nx.release(); ox.release(); dO.release(); dN.release(); off.release(); oi.release(); };
  auto fail = [&](cudaError_t e) { ctx->err = cudaGetErrorString(e); release(); return (int)MMB_ERR_CUDA; };
  cudaError_t e;
  if ((e = nx.ensure(4 * nnew)) || (e = ox.ensure(4 * npairs + 4)) || (e = dO.ensure(2 * n_old)) || (e = dN.ensure(2 * nnew)) || (e = off.ensure(nnew + 1)) || (e = oi.ensure(npairs + 1))) return fail(e);
  cudaMemcpyAsync(nx.p, new_xy_host, (size_t)nnew * 32, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(off.p, new_off_host, (size_t)(nnew + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (npairs) { cudaMemcpyAsync(ox.p, old_xy_host, (size_t)npairs * 32, cudaMemcpyHostToDevice, ctx->stream); cudaMemcpyAsync(oi.p, old_idx_host, (size_t)npairs * 4, cudaMemcpyHostToDevice, ctx->stream); }
  cudaMemcpyAsync(dO.p, dofs_old_host, (size_t)n_old * 16, cudaMemcpyHostToDevice, ctx->stream);
  LAUNCH("project_interface_p1", k_project_interface_p1, grid_for(nnew, 256), 256, (const double4*)nx.p, off.p, nnew, (const double4*)ox.p, oi.p,
         (const double2*)dO.p, (double2*)dN.p);
  cudaMemcpyAsync(dofs_new_host, dN.p, (size_t)nnew * 16, cudaMemcpyDeviceToHost, ctx->stream);
  e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(e);
  release();
  return MMB_OK;
}
int mmb_refinement_points(mmb_context* ctx, int64_t n, const int32_t* cells_host, double* points_host, int32_t* edge_vertices_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "refinement points need an uploaded mesh with marks");
  REQUIRE(n >= 0 && (n == 0 || (cells_host && points_host)), MMB_ERR_INVALID_ARG, "mmb_refinement_points: bad arguments");
  if (n == 0) return MMB_OK;
  for (int64_t i = 0; i < n; ++i)
    REQUIRE(cells_host[i] >= 0 && cells_host[i] < ctx->nc, MMB_ERR_INVALID_ARG, "mmb_refinement_points: cell index out of range");
  CK(ctx->list_out.ensure(n));
  if (ctx->dim == 3) {
    DBuf<double> out3; DBuf<int2> ev3;
    CK(out3.ensure(3 * (size_t)n)); CK(ev3.ensure(n));
    CK(cudaMemcpyAsync(ctx->list_out.p, cells_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH("refinement_points3d", k_refinement_points3d, grid_for(n, 256), 256, ctx->x.p, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, ctx->cv[3].p,
           ctx->perm.p, ctx->longest.p, ctx->list_out.p, n, out3.p, ev3.p);
    cudaMemcpyAsync(points_host, out3.p, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (edge_vertices_host) cudaMemcpyAsync(edge_vertices_host, ev3.p, n * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e3 = cudaStreamSynchronize(ctx->stream);
    out3.release(); ev3.release();
    if (e3 != cudaSuccess) { ctx->err = cudaGetErrorString(e3); return MMB_ERR_CUDA; }
    return MMB_OK;
  }
  DBuf<double2> out; DBuf<int2> ev;
  CK(out.ensure(n)); CK(ev.ensure(n));
  CK(cudaMemcpyAsync(ctx->list_out.p, cells_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH("refinement_points", k_refinement_points, grid_for(n, 256), 256, (const double2*)ctx->x.p, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p,
         ctx->perm.p, ctx->longest.p, ctx->list_out.p, n, out.p, ev.p);
  cudaMemcpyAsync(points_host, out.p, n * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream);
  if (edge_vertices_host) cudaMemcpyAsync(edge_vertices_host, ev.p, n * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  out.release(); ev.release();
  if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return MMB_ERR_CUDA; }
  return MMB_OK;
}

int mmb_upload_neighbours(mmb_context* ctx, const int32_t* cell_nb_host) {
  if (!ctx || !cell_nb_host) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->dim == 2, MMB_ERR_STATE, "neighbours need an uploaded 2-D mesh");
  for (int64_t i = 0; i < 3 * ctx->nc; ++i)
    REQUIRE(cell_nb_host[i] >= -1 && cell_nb_host[i] < ctx->nc, MMB_ERR_INVALID_ARG, "mmb_upload_neighbours: neighbour index out of range");
  CK(ctx->cell_nb.ensure(3 * (size_t)ctx->nc + 1));
  CK(cudaMemcpyAsync(ctx->cell_nb.p, cell_nb_host, 3 * (size_t)ctx->nc * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_nb = true;
  return MMB_OK;
}
int mmb_refinement_candidates(mmb_context* ctx, int32_t* cells_host, int64_t cap, int64_t* n_total) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  if (ctx->mesh_ok && ctx->dim == 3) {
    // size the table from the number of marked cells (every marked cell claims at most one edge)
    int64_t n_marked = 0;
    TRY(compact_flags(ctx, (const uint8_t*)ctx->mark.p, ctx->nc, 1, nullptr, 0, &n_marked));
    size_t T = 1024; while (T < 2 * (size_t)n_marked + 16) T <<= 1;
    REQUIRE(T <= (1ull << 31), MMB_ERR_CAPACITY, "refinement candidates: too many marked cells");
    CK(ctx->ht_key.ensure(T)); CK(ctx->ht_val.ensure(T)); CK(ctx->ht_slot.ensure((size_t)ctx->nc + 1));
    CK(cudaMemsetAsync(ctx->ht_key.p, 0xff, T * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->ht_val.p, 0x7f, T * sizeof(int32_t), ctx->stream));
    LAUNCH("refine3d_claim", k_refine3d_claim, grid_for(ctx->nc, 256), 256, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, ctx->cv[3].p, ctx->perm.p,
           ctx->mark.p, ctx->longest.p, ctx->v_if.p, ctx->facet_keys.p, (int)ctx->n_facet_keys, ctx->nc, ctx->ht_key.p, ctx->ht_val.p,
           (unsigned)(T - 1), ctx->ht_slot.p);
    LAUNCH("refine3d_flag", k_refine3d_flag, grid_for(ctx->nc, 256), 256, ctx->ht_slot.p, ctx->ht_val.p, ctx->nc, ctx->cand_flag.p);
    return compact_flags(ctx, ctx->cand_flag.p, ctx->nc, 1, cells_host, cap, n_total);
  }
  REQUIRE(ctx->mesh_ok && ctx->dim == 2 && ctx->have_nb, MMB_ERR_STATE, "refinement candidates need a 2-D mesh with neighbours (mmb_upload_neighbours) and marks");
  LAUNCH("refine_candidates", k_refine_candidates, grid_for(ctx->nc, 256), 256, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, ctx->perm.p,
         ctx->mark.p, ctx->longest.p, ctx->cell_nb.p, ctx->v_if.p, ctx->facet_keys.p, (int)ctx->ns, ctx->nc, ctx->cand_flag.p);
  return compact_flags(ctx, ctx->cand_flag.p, ctx->nc, 1, cells_host, cap, n_total);
}

// per-mesh state of the coarsening candidates, built on first use (not part of the per-step footprint)
static int ensure_candidate_state(mmb_context* ctx) {
  if (ctx->cand_ready) return MMB_OK;
  const size_t nv = (size_t)ctx->nv;
  CK(ctx->v_bnd.ensure(nv + 1)); CK(ctx->v_ifcount.ensure(nv + 1)); CK(ctx->v_level.ensure(nv + 1));
  CK(ctx->first_claim.ensure(nv + 1)); CK(ctx->first_seg.ensure(nv + 1));
  CK(ctx->chosen.ensure((size_t)std::max(ctx->nc, ctx->ns) + 1));
  CK(cudaMemsetAsync(ctx->v_bnd.p, 0, nv, ctx->stream));
  CK(cudaMemsetAsync(ctx->v_ifcount.p, 0, nv * sizeof(int32_t), ctx->stream));
  if (!ctx->have_levels) CK(cudaMemsetAsync(ctx->v_level.p, 0, nv * sizeof(int32_t), ctx->stream));   // VertexInfo default (mmesh.hh:132)
  if (ctx->ne) LAUNCH("flag_boundary_vertices", k_flag_boundary_vertices, grid_for(ctx->ne, 256), 256, (const int4*)ctx->edges.p, ctx->ne, ctx->v_bnd.p);
  if (ctx->ns) LAUNCH("count_interface_elements", k_count_interface_elements, grid_for(2 * ctx->ns, 256), 256, ctx->facets.p, 2 * ctx->ns, ctx->v_ifcount.p);
  CK(cudaGetLastError());
  ctx->cand_ready = true;
  return MMB_OK;
}
int mmb_upload_vertex_levels(mmb_context* ctx, const int32_t* level_host) {
  if (!ctx || !level_host) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->dim == 2, MMB_ERR_STATE, "vertex levels need an uploaded 2-D mesh");
  for (int64_t i = 0; i < ctx->nv; ++i)
    REQUIRE(level_host[i] >= 0, MMB_ERR_INVALID_ARG, "mmb_upload_vertex_levels: insertionLevel is unsigned in the reference");
  CK(ctx->v_level.ensure((size_t)ctx->nv + 1));
  CK(cudaMemcpyAsync(ctx->v_level.p, level_host, (size_t)ctx->nv * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_levels = true;
  return MMB_OK;
}
static int bulk_coarsen_choice(mmb_context* ctx) {
  TRY(ensure_candidate_state(ctx));
  CK(cudaMemsetAsync(ctx->first_claim.p, 0x7f, (size_t)ctx->nv * sizeof(int32_t), ctx->stream));
  LAUNCH("coarsen_choice", k_coarsen_choice, grid_for(ctx->nc, 256), 256, (const double2*)ctx->x.p, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p,
         ctx->perm.p, ctx->mark.p, ctx->v_if.p, ctx->v_bnd.p, ctx->v_level.p, ctx->nc, ctx->chosen.p, ctx->first_claim.p);
  return MMB_OK;
}
int mmb_coarsening_candidates(mmb_context* ctx, int32_t* cells_host, int32_t* vertices_host, int64_t cap, int64_t* n_total) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->dim == 2 && ctx->ne > 0, MMB_ERR_STATE, "coarsening candidates need a 2-D mesh with edges and marks");
  TRY(bulk_coarsen_choice(ctx));
  LAUNCH("coarsen_flag", k_coarsen_flag, grid_for(ctx->nc, 256), 256, ctx->chosen.p, ctx->first_claim.p, ctx->nc, ctx->cand_flag.p);
  int64_t tot = 0;
  TRY(compact_flags(ctx, ctx->cand_flag.p, ctx->nc, 1, cells_host, cap, &tot));
  if (n_total) *n_total = tot;
  const int64_t n = std::min(cap, tot);
  if (vertices_host && n > 0) {
    CK(ctx->wl.ensure((size_t)n));
    LAUNCH("gather_i32", k_gather_i32, grid_for(n, 256), 256, ctx->chosen.p, ctx->list_out.p, n, ctx->wl.p);
    CK(cudaMemcpyAsync(vertices_host, ctx->wl.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return MMB_OK;
}
int mmb_interface_candidates(mmb_context* ctx, int32_t* insert_segs_host, int64_t cap_insert, int64_t* n_insert,
                             int32_t* remove_segs_host, int32_t* remove_vertices_host, int64_t cap_remove, int64_t* n_remove) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->dim == 2 && ctx->ne > 0, MMB_ERR_STATE, "interface candidates need a 2-D mesh with edges and marks");
  if (n_insert) *n_insert = 0;
  if (n_remove) *n_remove = 0;
  if (ctx->ns == 0) return MMB_OK;
  // insert_: every interface element with mark +1 (its edge id cannot be in inserted_: the bulk loop skips interface edges)
  TRY(compact_flags(ctx, (const uint8_t*)ctx->imark.p, ctx->ns, 1, insert_segs_host, cap_insert, n_insert));
  // remove_: removed_ already holds what the bulk loop chose
  TRY(bulk_coarsen_choice(ctx));
  CK(cudaMemsetAsync(ctx->first_seg.p, 0x7f, (size_t)ctx->nv * sizeof(int32_t), ctx->stream));
  LAUNCH("if_coarsen_choice", k_if_coarsen_choice, grid_for(ctx->ns, 256), 256, (const int2*)ctx->facets.p, ctx->imark.p, ctx->v_ifcount.p,
         ctx->first_claim.p, (int)ctx->ns, ctx->chosen.p, ctx->first_seg.p);
  LAUNCH("coarsen_flag", k_coarsen_flag, grid_for(ctx->ns, 256), 256, ctx->chosen.p, ctx->first_seg.p, ctx->ns, ctx->cand_flag.p);
  int64_t tot = 0;
  TRY(compact_flags(ctx, ctx->cand_flag.p, ctx->ns, 1, remove_segs_host, cap_remove, &tot));
  if (n_remove) *n_remove = tot;
  const int64_t n = std::min(cap_remove, tot);
  if (remove_vertices_host && n > 0) {
    CK(ctx->wl.ensure((size_t)n));
    LAUNCH("gather_i32", k_gather_i32, grid_for(n, 256), 256, ctx->chosen.p, ctx->list_out.p, n, ctx->wl.p);
    CK(cudaMemcpyAsync(remove_vertices_host, ctx->wl.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return MMB_OK;
}

int mmb_component_numbers(mmb_context* ctx, int64_t nz, const int64_t* zone_off_host, const int32_t* zone_cells_host,
                          int64_t component_count, int32_t* zone_number_host, int32_t* cell_number_host, int64_t* component_count_out) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "component numbers need an uploaded mesh");
  REQUIRE(nz >= 0 && component_count >= 1 && (nz == 0 || (zone_off_host && zone_number_host)), MMB_ERR_INVALID_ARG, "mmb_component_numbers: bad arguments");
  const int64_t ne = nz ? zone_off_host[nz] : 0;
  REQUIRE(nz == 0 || (zone_off_host[0] == 0 && ne >= 0 && (ne == 0 || zone_cells_host)), MMB_ERR_INVALID_ARG, "mmb_component_numbers: malformed CSR");
  for (int64_t z = 0; z < nz; ++z) REQUIRE(zone_off_host[z + 1] >= zone_off_host[z], MMB_ERR_INVALID_ARG, "mmb_component_numbers: malformed CSR");
  for (int64_t e = 0; e < ne; ++e) REQUIRE(zone_cells_host[e] >= 0 && zone_cells_host[e] < ctx->nc, MMB_ERR_INVALID_ARG, "mmb_component_numbers: cell index out of range");
  REQUIRE(component_count + nz < 0x7f7f7f7fll, MMB_ERR_CAPACITY, "mmb_component_numbers: component numbers exceed int32");
  const size_t nc = (size_t)ctx->nc;
  DBuf<long long> off; DBuf<int32_t> cells, first, label, clabel; DBuf<uint32_t> fresh; DBuf<uint8_t> fflag; DBuf<int> changed;
  auto release = [&]() { off.release(); cells.release(); first.release(); label.release(); clabel.release(); fresh.release(); fflag.release(); changed.release(); };
  auto fail = [&](cudaError_t e) { ctx->err = cudaGetErrorString(e); release(); return (int)MMB_ERR_CUDA; };
  cudaError_t e;
  if ((e = off.ensure(nz + 1)) || (e = cells.ensure(ne + 1)) || (e = first.ensure(nc + 1)) || (e = label.ensure(nz + 1)) ||
      (e = clabel.ensure(nc + 1)) || (e = fresh.ensure(nz + 1)) || (e = fflag.ensure(nz + 1)) || (e = changed.ensure(1))) return fail(e);
  cudaStream_t S = ctx->stream;
  cudaMemsetAsync(clabel.p, 0x7f, nc * sizeof(int32_t), S);            // 0x7f7f7f7f: larger than any label; normalised below
  cudaMemsetAsync(first.p, 0x7f, nc * sizeof(int32_t), S);
  int64_t count_out = component_count;
  if (nz) {
    cudaMemcpyAsync(off.p, zone_off_host, (size_t)(nz + 1) * sizeof(long long), cudaMemcpyHostToDevice, S);
    if (ne) cudaMemcpyAsync(cells.p, zone_cells_host, (size_t)ne * sizeof(int32_t), cudaMemcpyHostToDevice, S);
    const unsigned g = grid_for(nz, 256);
    LAUNCH("cc_first", k_cc_first, g, 256, off.p, cells.p, nz, first.p);
    LAUNCH("cc_fresh", k_cc_fresh, g, 256, off.p, cells.p, nz, first.p, fresh.p, fflag.p);
    LAUNCH("compact_scan", k_compact_scan, 1, 1024, fresh.p, (int)nz, &ctx->counters.p->pad_[0]);
    LAUNCH("cc_init", k_cc_init, g, 256, fresh.p, fflag.p, nz, (int)component_count, label.p);
    for (int it = 0; it < 1 << 20; ++it) {
      cudaMemsetAsync(changed.p, 0, sizeof(int), S);
      LAUNCH("cc_to_cells", k_cc_to_cells, g, 256, off.p, cells.p, nz, label.p, clabel.p);
      LAUNCH("cc_to_zones", k_cc_to_zones, g, 256, off.p, cells.p, nz, clabel.p, label.p, changed.p);
      int h = 0;
      cudaMemcpyAsync(&h, changed.p, sizeof(int), cudaMemcpyDeviceToHost, S);
      if ((e = cudaStreamSynchronize(S)) != cudaSuccess) return fail(e);
      if (!h) break;
    }
    cudaMemcpyAsync(zone_number_host, label.p, (size_t)nz * sizeof(int32_t), cudaMemcpyDeviceToHost, S);
    if (fetch_counters(ctx) != MMB_OK) { release(); return MMB_ERR_CUDA; }
    count_out += (int64_t)ctx->h_counters->pad_[0];
  }
  if (cell_number_host) {
    // cells outside every zone keep componentNumber 0
    {
      LaunchScope ls_(ctx, "cc_finish");
      k_cc_finish<<<grid_for((int64_t)nc, 256), 256, 0, S>>>(clabel.p, (int64_t)nc);
    }
    cudaMemcpyAsync(cell_number_host, clabel.p, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, S);
  }
  e = cudaStreamSynchronize(S);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(e);
  if (component_count_out) *component_count_out = count_out;
  release();
  return MMB_OK;
}

// ---- whole step -----------------------------------------------------------------------------------------
static int step_core(mmb_context* ctx, const double* shifts_host, int ncomp, int64_t n_old, const double* dofs_old_host) {
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  REQUIRE(ncomp == 0 || ctx->dim == 2, MMB_ERR_INVALID_ARG, "projection is 2-D only (geometryInFather, entity.hh:573-594)");
  TRY(stage_shifts(ctx, shifts_host));
  if (ncomp) TRY(stage_dofs(ctx, ncomp, n_old, dofs_old_host));
  TRY(zero_counters(ctx));
  TRY(do_distance(ctx));            // markElements: indicator_.update()
  TRY(do_mark(ctx));                //               bulk marks (+ longest edge)
  if (ctx->dim == 2 && ctx->ns)
    LAUNCH("mark_interface2d", k_mark_interface2d, grid_for(ctx->ns, 256), 256, (const double2*)ctx->x.p, (const int2*)ctx->facets.p,
           (int)ctx->ns, dev_params(ctx->prm), ctx->counters.p, ctx->imark.p);
  TRY(do_validate(ctx, true));      // ensureVertexMovement (trial move, coordinates untouched)
  if (ncomp) TRY(do_project(ctx, ncomp));   // adapt(handle): old children -> new cells, before the real move
  TRY(do_move(ctx, true));          // moveVertices (withheld when a cell would invert, unless move_policy == 1)
  if (ctx->dim == 2) TRY(do_legality(ctx));   // in-circle flags of the moved mesh (input of the next host TDS pass)
  CK(cudaGetLastError());
  return MMB_OK;
}
static int step_result(mmb_context* ctx, mmb_step_result* res) {
  if (res) {
    TRY(fetch_counters(ctx));
    const Counters& c = *ctx->h_counters;
    res->n_invalid = (int64_t)c.n_invalid; res->n_orient_nonpos = (int64_t)c.n_orient_nonpos; res->n_illegal = (int64_t)c.n_illegal;
    res->n_refine = (int64_t)c.n_refine; res->n_coarsen = (int64_t)c.n_coarsen;
    res->n_exact_orient = (int64_t)c.n_exact_orient; res->n_exact_incircle = (int64_t)c.n_exact_incircle;
    res->n_pieces = (int64_t)c.n_pieces; memcpy(&res->max_dist, &c.max_dist_bits, 8);
    res->mass_new = c.mass_new; res->covered_area = c.covered_area;
    const bool withheld = ctx->prm.move_policy == 0 && c.n_invalid_global > 0;
    res->moved = withheld ? 0 : 1;
    TRY(check_predicate_domain(ctx));
    if (ctx->dim == 3 && c.n_invalid > 0) {
      ctx->err = "Vertices could not be moved, because the removal of vertices is not supported in 3D!";
      return MMB_ERR_INVALID_CELL_3D;
    }
    if (withheld) {
      ctx->err = "moveVertices withheld: the trial move inverts cells (ensureVertexMovement == true); remove the vertices first";
      return MMB_MOVE_WITHHELD;
    }
  }
  return MMB_OK;
}
int mmb_adapt_step(mmb_context* ctx, const double* shifts_host, int ncomp, mmb_step_result* res) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  TRY(step_core(ctx, shifts_host, ncomp, 0, nullptr));
  return step_result(ctx, res);
}
// Host-buffer step, pipelined over three streams: the H2D of the shifts overlaps Distance::update + marks, the H2D of
// the old DOFs overlaps the validity pass, the projection runs in chunks whose new DOFs stream back while the next
// chunk is computed, and the flag arrays return behind the kernels that produced them.  (Pinned host buffers are
// needed for real overlap; pageable ones still work, serialised by the driver.)
// The three phases are exported separately so that a sharded caller can put its collectives between them
// (multigpu.py): begin -> [all-reduce MAX of the distance maximum] -> marks -> [halo exchange of the shifts] -> finish.
int mmb_step_host_begin(mmb_context* ctx, const double* shifts_host, int ncomp, int64_t n_old, const double* dofs_old_host,
                        int will_download_dofs) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  REQUIRE(ncomp == 0 || ctx->dim == 2, MMB_ERR_INVALID_ARG, "projection is 2-D only (geometryInFather, entity.hh:573-594)");
  REQUIRE(ncomp == 0 || ncomp == 1 || ncomp == 3, MMB_ERR_INVALID_ARG, "ncomp must be 0, 1 (P0) or 3 (DG-P1)");
  REQUIRE(shifts_host || ctx->has_shift, MMB_ERR_STATE, "shifts == NULL but no resident shifts");
  if (ncomp && ctx->nnew > 0)      // same bound as do_project: the pipelined path launches project_range directly
    REQUIRE(ctx->max_old_idx < (dofs_old_host ? n_old : ctx->n_old), MMB_ERR_INVALID_ARG,
            "projection: a pair references an old cell beyond the DOF array");
  if (!ctx->s_h2d) { CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking)); }
  while (ctx->pipe_ev.size() < 8 + 2 * (size_t)MAX_CHUNKS) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->pipe_ev.push_back(e); }
  cudaStream_t S0 = ctx->stream, S1 = ctx->s_h2d, S2 = ctx->s_d2h;
  cudaEvent_t* ev = ctx->pipe_ev.data();
  const size_t nc = (size_t)ctx->nc;
  // buffers must exist before any asynchronous copy targets them
  if (ncomp) {
    if (dofs_old_host) { REQUIRE(n_old > 0, MMB_ERR_INVALID_ARG, "n_old must be positive"); CK(ctx->dofs_old.ensure((size_t)n_old * ncomp)); }
    else REQUIRE(ctx->dofs_ncomp == ncomp, MMB_ERR_STATE, "dofs_old == NULL but no resident DOFs of this ncomp");
    if (ctx->dofs_new.cap < nc * ncomp) { CK(ctx->dofs_new.ensure(nc * ncomp)); CK(cudaMemsetAsync(ctx->dofs_new.p, 0, nc * ncomp * sizeof(double), S0)); }
  }
  CK(cudaEventRecord(ev[0], S0));                                  // everything issued so far (previous step) is ordered first
  CK(cudaStreamWaitEvent(S1, ev[0], 0)); CK(cudaStreamWaitEvent(S2, ev[0], 0));
  // S1: inputs
  if (shifts_host) CK(cudaMemcpyAsync(ctx->shift.p, shifts_host, (size_t)ctx->nv * ctx->dim * sizeof(double), cudaMemcpyHostToDevice, S1));
  ctx->has_shift = true;
  CK(cudaEventRecord(ev[1], S1));
  const bool pipe_dofs = ncomp && dofs_old_host && will_download_dofs && ctx->nnew && ctx->chunks_ok && ctx->dof_pipe_ok &&
                         (ctx->dof_lo.empty() || ctx->dof_lo.back() <= n_old);
  if (ncomp && dofs_old_host) {
    if (pipe_dofs) {
      // remote DOFs first (packed on the host, scattered on the device), then one index range per projection chunk
      const size_t nr = ctx->remote_idx.size();
      if (nr) {
        if (ctx->h_remote_cap < nr * ncomp) {
          if (ctx->h_remote_stage) cudaFreeHost(ctx->h_remote_stage);
          CK(cudaMallocHost((void**)&ctx->h_remote_stage, nr * 3 * sizeof(double))); ctx->h_remote_cap = nr * 3;
        }
        CK(ctx->d_remote_stage.ensure(nr * 3));
        for (size_t r = 0; r < nr; ++r)
          for (int c = 0; c < ncomp; ++c) ctx->h_remote_stage[r * ncomp + c] = dofs_old_host[(size_t)ctx->remote_idx[r] * ncomp + c];
        CK(cudaMemcpyAsync(ctx->d_remote_stage.p, ctx->h_remote_stage, nr * ncomp * sizeof(double), cudaMemcpyHostToDevice, S1));
        { LaunchScope ls_(ctx, "scatter_dofs");
          k_scatter_dofs<<<grid_for((int64_t)nr * ncomp, 256), 256, 0, S1>>>(ctx->dofs_old.p, ctx->d_remote_idx.p, (int64_t)nr, ncomp, ctx->d_remote_stage.p); }
      }
      const int C = (int)ctx->chunk_cell_lo.size();
      for (int k = 0; k < C; ++k) {
        const size_t lo = (size_t)ctx->dof_lo[k], hi = k == C - 1 ? (size_t)n_old : (size_t)ctx->dof_lo[k + 1];
        if (hi > lo) CK(cudaMemcpyAsync(ctx->dofs_old.p + lo * ncomp, dofs_old_host + lo * ncomp, (hi - lo) * ncomp * sizeof(double), cudaMemcpyHostToDevice, S1));
        CK(cudaEventRecord(ev[8 + MAX_CHUNKS + k], S1));
      }
    } else {
      CK(cudaMemcpyAsync(ctx->dofs_old.p, dofs_old_host, (size_t)n_old * ncomp * sizeof(double), cudaMemcpyHostToDevice, S1));
    }
    ctx->n_old = n_old; ctx->dofs_ncomp = ncomp;
  }
  CK(cudaEventRecord(ev[2], S1));
  ctx->hp_active = true; ctx->hp_pipe_dofs = pipe_dofs; ctx->hp_ncomp = ncomp;
  // S0: Distance::update needs neither shifts nor DOFs
  TRY(zero_counters(ctx));
  TRY(do_distance(ctx));
  CK(cudaGetLastError());
  return MMB_OK;
}
int mmb_step_host_marks(mmb_context* ctx, int8_t* marks_host, uint8_t* longest_edge_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->hp_active, MMB_ERR_STATE, "mmb_step_host_marks without mmb_step_host_begin");
  cudaStream_t S0 = ctx->stream, S2 = ctx->s_d2h;
  cudaEvent_t* ev = ctx->pipe_ev.data();
  const size_t nc = (size_t)ctx->nc;
  TRY(do_mark(ctx));
  if (ctx->dim == 2 && ctx->ns)
    LAUNCH("mark_interface2d", k_mark_interface2d, grid_for(ctx->ns, 256), 256, (const double2*)ctx->x.p, (const int2*)ctx->facets.p,
           (int)ctx->ns, dev_params(ctx->prm), ctx->counters.p, ctx->imark.p);
  CK(cudaEventRecord(ev[3], S0));
  CK(cudaStreamWaitEvent(S2, ev[3], 0));
  if (marks_host) CK(cudaMemcpyAsync(marks_host, ctx->mark.p, nc, cudaMemcpyDeviceToHost, S2));
  if (longest_edge_host) CK(cudaMemcpyAsync(longest_edge_host, ctx->longest.p, nc, cudaMemcpyDeviceToHost, S2));
  CK(cudaStreamWaitEvent(S0, ev[1], 0));                           // shifts have landed: whatever follows on S0 may read them
  CK(cudaGetLastError());
  return MMB_OK;
}
int mmb_step_host_finish(mmb_context* ctx, double* dofs_new_host, uint8_t* valid_fp_host, int8_t* orient_sign_host,
                         uint8_t* edge_flags_host, int reduce_vector, mmb_step_result* res) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->hp_active, MMB_ERR_STATE, "mmb_step_host_finish without mmb_step_host_begin");
  ctx->hp_active = false;
  cudaStream_t S0 = ctx->stream, S2 = ctx->s_d2h;
  cudaEvent_t* ev = ctx->pipe_ev.data();
  const size_t nc = (size_t)ctx->nc;
  const int ncomp = ctx->hp_ncomp;
  const bool pipe_dofs = ctx->hp_pipe_dofs && dofs_new_host;
  TRY(do_validate(ctx, true));
  CK(cudaEventRecord(ev[4], S0));
  CK(cudaStreamWaitEvent(S2, ev[4], 0));
  if (valid_fp_host) CK(cudaMemcpyAsync(valid_fp_host, ctx->valid_fp.p, nc, cudaMemcpyDeviceToHost, S2));
  if (orient_sign_host) CK(cudaMemcpyAsync(orient_sign_host, ctx->orient.p, nc, cudaMemcpyDeviceToHost, S2));
  if (ctx->dim == 2) {                                             // legality of the moved mesh, evaluated on x (+) shift so
    TRY(do_legality(ctx, true));                                   // that its flags leave before the projection's DOFs
    CK(cudaEventRecord(ev[5], S0));
    CK(cudaStreamWaitEvent(S2, ev[5], 0));
    if (edge_flags_host && ctx->ne) CK(cudaMemcpyAsync(edge_flags_host, ctx->edge_flag.p, (size_t)ctx->ne, cudaMemcpyDeviceToHost, S2));
  }
  if (ncomp && ctx->nnew) {
    if (!pipe_dofs) CK(cudaStreamWaitEvent(S0, ev[2], 0));         // old DOFs have landed (monolithic upload / all ranges)
    const bool chunked = ctx->chunks_ok && dofs_new_host;
    const int C = chunked ? (int)ctx->chunk_cell_lo.size() : 1;
    int pb = 0;
    for (int k = 0; k < C; ++k) {
      const int64_t i0 = chunked ? ctx->chunk_i[k] : 0, i1 = chunked ? ctx->chunk_i[k + 1] : ctx->nnew;
      int nb = 0;
      if (pipe_dofs) CK(cudaStreamWaitEvent(S0, ev[8 + MAX_CHUNKS + k], 0));   // this chunk's home range (and the remote DOFs) have landed
      TRY(project_range(ctx, ncomp, i0, i1, pb, &nb));
      pb += nb;
      if (dofs_new_host) {
        CK(cudaEventRecord(ev[8 + k], S0));
        CK(cudaStreamWaitEvent(S2, ev[8 + k], 0));
        // chunk k owns cells (hi of chunk k-1, hi of chunk k]; cells outside the projected set keep their old values
        const size_t lo = chunked ? (k == 0 ? 0 : (size_t)ctx->chunk_cell_hi[k - 1] + 1) : 0;
        const size_t hi = chunked ? (k == C - 1 ? nc : (size_t)ctx->chunk_cell_hi[k] + 1) : nc;
        if (hi > lo)
          CK(cudaMemcpyAsync(dofs_new_host + lo * ncomp, ctx->dofs_new.p + lo * ncomp, (hi - lo) * ncomp * sizeof(double), cudaMemcpyDeviceToHost, S2));
      }
    }
    TRY(reduce_partials(ctx, pb));
  } else if (ncomp && dofs_new_host) {
    CK(cudaStreamWaitEvent(S2, ev[4], 0));
    CK(cudaMemcpyAsync(dofs_new_host, ctx->dofs_new.p, nc * ncomp * sizeof(double), cudaMemcpyDeviceToHost, S2));
  }
  TRY(do_move(ctx, true));
  CK(cudaEventRecord(ev[6], S2));
  if (reduce_vector) {                                             // sharded caller: counts + mass for the SUM all-reduce
    CK(ctx->reduce_vec.ensure(16));
    LAUNCH("reduce_vector", k_reduce_vector, 1, 32, ctx->counters.p, ctx->reduce_vec.p);
  }
  CK(cudaStreamWaitEvent(S0, ev[6], 0));                           // the step is complete on S0 only when all copies are
  CK(cudaGetLastError());
  if (res) return step_result(ctx, res);
  if (!reduce_vector) CK(cudaStreamSynchronize(S0));
  return MMB_OK;
}
int mmb_adapt_step_host(mmb_context* ctx, const double* shifts_host, int ncomp, int64_t n_old, const double* dofs_old_host,
                        double* dofs_new_host, int8_t* marks_host, uint8_t* longest_edge_host, uint8_t* valid_fp_host,
                        int8_t* orient_sign_host, uint8_t* edge_flags_host, mmb_step_result* res) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  TRY(mmb_step_host_begin(ctx, shifts_host, ncomp, n_old, dofs_old_host, dofs_new_host != nullptr));
  const int rc = mmb_step_host_marks(ctx, marks_host, longest_edge_host);
  if (rc != MMB_OK) { ctx->hp_active = false; return rc; }
  return mmb_step_host_finish(ctx, dofs_new_host, valid_fp_host, orient_sign_host, edge_flags_host, 0, res);
}

// ---- 3-D P0 trace / skeleton ----------------------------------------------------------------------------
int mmb_trace_skeleton_p0(mmb_context* ctx, int64_t nf, const int32_t* inside_host, const int32_t* outside_host,
                          const double* u_bulk_host, int64_t n_bulk, const double* u_gamma_host, double* trace_in_host,
                          double* trace_out_host, double* skeleton_host) {
  if (!ctx || nf <= 0 || !inside_host || !outside_host || !u_bulk_host || !u_gamma_host || n_bulk <= 0) return MMB_ERR_INVALID_ARG;
  CK(ctx->tr_in.ensure(nf)); CK(ctx->tr_out.ensure(nf)); CK(ctx->tr_u.ensure(n_bulk)); CK(ctx->tr_ug.ensure(nf));
  CK(ctx->tr_a.ensure(nf)); CK(ctx->tr_b.ensure(nf)); CK(ctx->tr_c.ensure(nf));
  CK(cudaMemcpyAsync(ctx->tr_in.p, inside_host, nf * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tr_out.p, outside_host, nf * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tr_u.p, u_bulk_host, n_bulk * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tr_ug.p, u_gamma_host, nf * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH("trace_skeleton", k_trace_skeleton, grid_for(nf, 256), 256, nf, ctx->tr_in.p, ctx->tr_out.p, ctx->tr_u.p, ctx->tr_ug.p,
         ctx->tr_a.p, ctx->tr_b.p, ctx->tr_c.p);
  CK(cudaGetLastError());
  if (trace_in_host) CK(cudaMemcpyAsync(trace_in_host, ctx->tr_a.p, nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (trace_out_host) CK(cudaMemcpyAsync(trace_out_host, ctx->tr_b.p, nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (skeleton_host) CK(cudaMemcpyAsync(skeleton_host, ctx->tr_c.p, nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MMB_OK;
}

// ---- predicate batches ----------------------------------------------------------------------------------
int mmb_predicate_batch(mmb_context* ctx, int which, int64_t n, const double* pts_host, int8_t* sign_host, int64_t* n_exact) {
  if (!ctx || which < 0 || which > 2 || n <= 0 || !pts_host || !sign_host) return MMB_ERR_INVALID_ARG;
  REQUIRE(n < (1ll << 31), MMB_ERR_CAPACITY, "batch too large");
  const int k = which == 0 ? 6 : (which == 1 ? 8 : 12);
  CK(ctx->pred_pts.ensure((size_t)n * k)); CK(ctx->pred_sign.ensure(n)); CK(ctx->wl.ensure(n + 4));
  CK(cudaMemcpyAsync(ctx->pred_pts.p, pts_host, (size_t)n * k * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(&ctx->counters.p->wl_orient, 0, sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(&ctx->counters.p->n_exact_orient, 0, sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(&ctx->counters.p->pad_[1], 0, sizeof(unsigned long long), ctx->stream));
  LAUNCH("pred_filter", k_pred_filter, grid_for(n, 256), 256, which, ctx->pred_pts.p, n, ctx->pred_sign.p, ctx->wl.p, ctx->counters.p);
  LAUNCH("pred_exact", k_pred_exact, ctx->num_sms * 4, 128, which, ctx->pred_pts.p, ctx->wl.p, ctx->pred_sign.p, ctx->counters.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(sign_host, ctx->pred_sign.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  TRY(fetch_counters(ctx));
  if (n_exact) *n_exact = (int64_t)ctx->h_counters->n_exact_orient;
  return check_predicate_domain(ctx);
}

// asynchronous field transfers on the context's stream (host buffers should be pinned); mmb_synchronize completes them
int mmb_upload_dofs(mmb_context* ctx, int ncomp, int64_t n_old, const double* dofs_old_host) {
  if (!ctx || !dofs_old_host) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  return stage_dofs(ctx, ncomp, n_old, dofs_old_host);
}
int mmb_download_fields(mmb_context* ctx, int ncomp, double* dofs_new_host, int8_t* marks_host, uint8_t* longest_edge_host,
                        uint8_t* valid_fp_host, int8_t* orient_sign_host, uint8_t* edge_flags_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  const size_t nc = (size_t)ctx->nc;
  if (marks_host) CK(cudaMemcpyAsync(marks_host, ctx->mark.p, nc, cudaMemcpyDeviceToHost, ctx->stream));
  if (longest_edge_host) CK(cudaMemcpyAsync(longest_edge_host, ctx->longest.p, nc, cudaMemcpyDeviceToHost, ctx->stream));
  if (valid_fp_host) CK(cudaMemcpyAsync(valid_fp_host, ctx->valid_fp.p, nc, cudaMemcpyDeviceToHost, ctx->stream));
  if (orient_sign_host) CK(cudaMemcpyAsync(orient_sign_host, ctx->orient.p, nc, cudaMemcpyDeviceToHost, ctx->stream));
  if (edge_flags_host && ctx->ne) CK(cudaMemcpyAsync(edge_flags_host, ctx->edge_flag.p, (size_t)ctx->ne, cudaMemcpyDeviceToHost, ctx->stream));
  if (dofs_new_host && ncomp) {
    REQUIRE(ctx->dofs_new.cap >= nc * ncomp, MMB_ERR_STATE, "no projected DOFs of this ncomp");
    CK(cudaMemcpyAsync(dofs_new_host, ctx->dofs_new.p, nc * ncomp * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  return MMB_OK;
}

// ---- f4: Delaunay restoration by edge flips on the resident snapshot (flip.cuh) ------------------------------------------
int mmb_restore_delaunay(mmb_context* ctx, const int64_t* vertex_id_host, int32_t max_rounds, mmb_flip_result* out) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  if (out) std::memset(out, 0, sizeof(*out));
  REQUIRE(ctx->mesh_ok && ctx->dim == 2 && ctx->have_nb, MMB_ERR_STATE, "mmb_restore_delaunay needs a 2-D mesh with neighbours (mmb_upload_neighbours)");
  REQUIRE(ctx->own_lo == 0 && ctx->own_hi == ctx->nc, MMB_ERR_STATE, "mmb_restore_delaunay works on an unsharded snapshot");
  REQUIRE(max_rounds >= 0, MMB_ERR_INVALID_ARG, "mmb_restore_delaunay: max_rounds < 0");
  const int64_t nc = ctx->nc, nv = ctx->nv;
  REQUIRE(3 * nc < 0xffffffffll, MMB_ERR_CAPACITY, "mmb_restore_delaunay: edge ids exceed 32 bits");
  if (nc == 0) { if (out) out->converged = 1; return MMB_OK; }
  DBuf<unsigned> claim; DBuf<uint8_t> ill, sel, act, act_next; DBuf<FlipCounts> fc; DBuf<int64_t> vid;
  auto release = [&]() { claim.release(); ill.release(); sel.release(); act.release(); act_next.release(); fc.release(); vid.release(); };
  auto fail = [&](cudaError_t e) { ctx->err = std::string("mmb_restore_delaunay: ") + cudaGetErrorString(e); release(); return (int)MMB_ERR_CUDA; };
  cudaError_t e;
  if ((e = claim.ensure((size_t)nc)) || (e = ill.ensure((size_t)nc)) || (e = sel.ensure((size_t)nc)) || (e = act.ensure((size_t)nc)) ||
      (e = act_next.ensure((size_t)nc)) || (e = fc.ensure(1)) || (vertex_id_host && (e = vid.ensure((size_t)nv)))) return fail(e);
  cudaStream_t S = ctx->stream;
  if (vertex_id_host && (e = cudaMemcpyAsync(vid.p, vertex_id_host, (size_t)nv * sizeof(int64_t), cudaMemcpyHostToDevice, S))) return fail(e);
  cudaMemsetAsync(fc.p, 0, sizeof(FlipCounts), S);
  cudaMemsetAsync(act.p, 1, (size_t)nc, S);
  if ((e = ctx->flip_log.ensure((size_t)nc))) return fail(e);      // room for nc flips; beyond that only the count is kept
  ctx->flip_round_off.clear(); ctx->flip_log_total = 0;
  FlipCounts h; std::memset(&h, 0, sizeof(h));
  int rounds = 0, converged = 0;
  bool full = true;                                           // the sweep covers every cell (first round, and the one that reports)
  for (;;) {
    // mark: counts of this sweep (flips accumulate)
    cudaMemsetAsync(fc.p, 0, offsetof(FlipCounts, flips), S);
    cudaMemsetAsync(claim.p, 0xff, (size_t)nc * sizeof(unsigned), S);
    LAUNCH("flip_mark", k_flip_mark, grid_for(nc, 128), 128, ctx->x.p, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, ctx->cell_nb.p, ctx->v_if.p,
           ctx->facet_keys.p, (int)ctx->n_facet_keys, nc, act.p, claim.p, ill.p, fc.p);
    if ((e = cudaMemcpyAsync(&h, fc.p, sizeof(FlipCounts), cudaMemcpyDeviceToHost, S)) || (e = cudaStreamSynchronize(S))) return fail(e);
    if (h.broken) { release(); REQUIRE(false, MMB_ERR_INVALID_ARG, "mmb_restore_delaunay: the neighbour table is not symmetric"); }
    if (h.domain) { release(); REQUIRE(false, MMB_ERR_CAPACITY, "mmb_restore_delaunay: coordinates outside the exact predicates' exponent range"); }
    if (h.inverted) {
      release();
      REQUIRE(false, MMB_ERR_STATE, "mmb_restore_delaunay: " + std::to_string(h.inverted) + " cells are not positively oriented; flips need a valid triangulation (ensureVertexMovement first)");
    }
    const bool done = h.illegal == 0, out_of_rounds = rounds >= max_rounds;
    if (done || out_of_rounds) {
      // the counts reported come from a sweep over ALL cells (a partial sweep sees every illegal edge, but not every illegal
      // interface edge); for `done` that sweep also confirms the result independently of the bookkeeping of active cells
      if (!full) { cudaMemsetAsync(act.p, 1, (size_t)nc, S); full = true; continue; }
      converged = done ? 1 : 0;
      break;
    }
    ctx->flip_round_off.push_back((int64_t)h.flips);          // flips of the earlier rounds: where this round's log entries start
    cudaMemsetAsync(act_next.p, 0, (size_t)nc, S);
    LAUNCH("flip_select", k_flip_select, grid_for(nc, 256), 256, ctx->cell_nb.p, claim.p, ill.p, nc, sel.p);
    LAUNCH("flip_apply", k_flip_apply, grid_for(nc, 256), 256, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, ctx->cell_nb.p, ctx->cells_aos.p,
           vertex_id_host ? vid.p : (const int64_t*)nullptr, ctx->perm.p, sel.p, ill.p, act_next.p, nc, fc.p, ctx->flip_log.p,
           (unsigned long long)nc);
    std::swap(act.p, act_next.p); std::swap(act.cap, act_next.cap);
    full = false;
    ++rounds;
  }
  ctx->flip_round_off.push_back((int64_t)h.flips); ctx->flip_log_total = (int64_t)h.flips;
  if (out) {
    out->n_flips = (int64_t)h.flips; out->n_illegal_left = (int64_t)h.illegal; out->n_constrained_illegal = (int64_t)h.constrained;
    out->rounds = rounds; out->converged = converged;
  }
  if (h.flips) {
    // derived arrays: edge list (count is invariant under flips), tile windows; everything computed from the old topology is stale
    const int64_t ntc = (ctx->nc_pad + TILE - 1) / TILE, nte = (ctx->ne + TILE - 1) / TILE;
    LAUNCH("tile_windows", k_tile_windows, (unsigned)ntc, 256, ctx->cv[0].p, 1, nc, nv, ctx->tile_lo_cells.p);
    ctx->nnew = ctx->npairs = 0; ctx->graph_valid = false; ctx->cand_ready = false;
  }
  if (h.flips && ctx->ne) {
    const int nb_blocks = (int)((nc + COMPACT_TILE - 1) / COMPACT_TILE);
    LAUNCH("edges_count", k_edges_count, nb_blocks, 256, ctx->cell_nb.p, nc, ctx->block_count.p);
    LAUNCH("compact_scan", k_compact_scan, 1, 1024, ctx->block_count.p, nb_blocks, &ctx->counters.p->pad_[0]);
    LAUNCH("edges_emit", k_edges_emit, nb_blocks, 256, ctx->cv[0].p, ctx->cv[1].p, ctx->cv[2].p, ctx->cell_nb.p, nc, ctx->block_count.p,
           (int4*)ctx->edges.p, ctx->ne);
    const int64_t nte = (ctx->ne + TILE - 1) / TILE;
    LAUNCH("tile_windows", k_tile_windows, (unsigned)nte, 256, (const int32_t*)ctx->edges.p, 4, ctx->ne, nv, ctx->tile_lo_edges.p);
    if ((e = cudaGetLastError())) return fail(e);
    if (int rc = fetch_counters(ctx)) { release(); return rc; }
    if ((int64_t)ctx->h_counters->pad_[0] != ctx->ne) {
      release();
      REQUIRE(false, MMB_ERR_STATE, "mmb_restore_delaunay: the rebuilt edge list has " + std::to_string(ctx->h_counters->pad_[0]) + " edges, the snapshot " + std::to_string(ctx->ne));
    }
  }
  if ((e = cudaStreamSynchronize(S))) return fail(e);
  release();
  return MMB_OK;
}
int mmb_flip_log(mmb_context* ctx, uint32_t* edge_id_host, int64_t cap, int64_t* n_total, int64_t* round_off_host, int32_t round_cap,
                 int32_t* n_rounds) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(cap >= 0 && round_cap >= 0 && (cap == 0 || edge_id_host) && (round_cap == 0 || round_off_host), MMB_ERR_INVALID_ARG, "mmb_flip_log: bad arguments");
  const int64_t rounds = ctx->flip_round_off.empty() ? 0 : (int64_t)ctx->flip_round_off.size() - 1;
  if (n_total) *n_total = ctx->flip_log_total;
  if (n_rounds) *n_rounds = (int32_t)rounds;
  for (int64_t r = 0; r <= rounds && r < round_cap && !ctx->flip_round_off.empty(); ++r) round_off_host[r] = ctx->flip_round_off[(size_t)r];
  const int64_t n = std::min<int64_t>(std::min<int64_t>(cap, ctx->flip_log_total), (int64_t)ctx->flip_log.cap);
  if (n > 0) {
    CK(cudaMemcpyAsync(edge_id_host, ctx->flip_log.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return MMB_OK;
}
int mmb_download_topology(mmb_context* ctx, int32_t* cells_host, uint8_t* perm_host, int32_t* cell_nb_host, int32_t* edges_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  const size_t nc = (size_t)ctx->nc, k = (size_t)ctx->dim + 1;
  if (cells_host && nc) CK(cudaMemcpyAsync(cells_host, ctx->cells_aos.p, nc * k * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (perm_host && nc) CK(cudaMemcpyAsync(perm_host, ctx->perm.p, nc, cudaMemcpyDeviceToHost, ctx->stream));
  if (cell_nb_host) {
    REQUIRE(ctx->have_nb, MMB_ERR_STATE, "no neighbours uploaded");
    if (nc) CK(cudaMemcpyAsync(cell_nb_host, ctx->cell_nb.p, 3 * nc * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (edges_host && ctx->ne) CK(cudaMemcpyAsync(edges_host, ctx->edges.p, (size_t)ctx->ne * sizeof(int4), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MMB_OK;
}

// ---- multi-GPU plumbing: ownership, halo pack/unpack, phased step ------------------------------------------------
int mmb_set_owned_cells(mmb_context* ctx, int64_t lo, int64_t hi) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && 0 <= lo && lo <= hi && hi <= ctx->nc, MMB_ERR_INVALID_ARG, "bad owned range");
  ctx->own_lo = lo; ctx->own_hi = hi;
  return MMB_OK;
}
int mmb_halo_setup(mmb_context* ctx, int64_t n_send, const int32_t* send_idx_host, int64_t n_recv, const int32_t* recv_idx_host) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && n_send >= 0 && n_recv >= 0 && (n_send == 0 || send_idx_host) && (n_recv == 0 || recv_idx_host),
          MMB_ERR_INVALID_ARG, "mmb_halo_setup: bad arguments");
  CK(ctx->halo_send_idx.ensure(n_send + 1)); CK(ctx->halo_recv_idx.ensure(n_recv + 1));
  CK(ctx->halo_send.ensure(n_send + 1)); CK(ctx->halo_recv.ensure(n_recv + 1)); CK(ctx->reduce_vec.ensure(16));
  if (n_send) CK(cudaMemcpyAsync(ctx->halo_send_idx.p, send_idx_host, n_send * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (n_recv) CK(cudaMemcpyAsync(ctx->halo_recv_idx.p, recv_idx_host, n_recv * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->n_halo_send = n_send; ctx->n_halo_recv = n_recv;
  return MMB_OK;
}
int mmb_halo_pack_shifts(mmb_context* ctx) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->dim == 2 && ctx->has_shift, MMB_ERR_STATE, "halo pack needs a 2-D mesh with resident shifts");
  if (ctx->n_halo_send)
    LAUNCH("halo_pack", k_gather2, grid_for(ctx->n_halo_send, 256), 256, (const double2*)ctx->shift.p, ctx->halo_send_idx.p,
           ctx->n_halo_send, ctx->halo_send.p);
  CK(cudaGetLastError());
  return MMB_OK;
}
int mmb_halo_unpack_shifts(mmb_context* ctx) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->dim == 2 && ctx->has_shift, MMB_ERR_STATE, "halo unpack needs a 2-D mesh with resident shifts");
  if (ctx->n_halo_recv)
    LAUNCH("halo_unpack", k_scatter2, grid_for(ctx->n_halo_recv, 256), 256, (double2*)ctx->shift.p, ctx->halo_recv_idx.p,
           ctx->n_halo_recv, (const double2*)ctx->halo_recv.p);
  CK(cudaGetLastError());
  return MMB_OK;
}
int mmb_device_buffer(mmb_context* ctx, int which, void** ptr, int64_t* bytes) {
  if (!ctx || !ptr || !bytes) return MMB_ERR_INVALID_ARG;
  switch (which) {
    case 0: *ptr = ctx->halo_send.p; *bytes = ctx->n_halo_send * (int64_t)sizeof(double2); break;
    case 1: *ptr = ctx->halo_recv.p; *bytes = ctx->n_halo_recv * (int64_t)sizeof(double2); break;
    case 2: *ptr = &ctx->counters.p->max_dist_bits; *bytes = 8; break;
    case 3: *ptr = ctx->reduce_vec.p; *bytes = 16 * 8; break;
    default: return MMB_ERR_INVALID_ARG;
  }
  return MMB_OK;
}
// phase A: counters reset + Distance::update on the local vertices (local max in the counters)
int mmb_step_phase_distance(mmb_context* ctx) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok, MMB_ERR_STATE, "no mesh uploaded");
  TRY(zero_counters(ctx));
  TRY(do_distance(ctx));
  CK(cudaGetLastError());
  return MMB_OK;
}
// phases after the distance update.  which = 1: trial validity + projection (independent of the global max distance, so
// the MAX all-reduce can run concurrently); 2: marks (need the global max); 3: moveVertices + legality + the reduce
// vector; 0: all three in the reference order (marks first).
int mmb_step_phase(mmb_context* ctx, int which, int ncomp) {
  if (!ctx || which < 0 || which > 3) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->has_shift, MMB_ERR_STATE, "no mesh / shifts");
  auto marks = [&]() -> int {
    TRY(do_mark(ctx));
    if (ctx->dim == 2 && ctx->ns)
      LAUNCH("mark_interface2d", k_mark_interface2d, grid_for(ctx->ns, 256), 256, (const double2*)ctx->x.p, (const int2*)ctx->facets.p,
             (int)ctx->ns, dev_params(ctx->prm), ctx->counters.p, ctx->imark.p);
    return MMB_OK;
  };
  auto validate_project = [&]() -> int {
    if (ncomp) TRY(stage_dofs(ctx, ncomp, 0, nullptr));
    TRY(do_validate(ctx, true));
    if (ncomp) TRY(do_project(ctx, ncomp));
    return MMB_OK;
  };
  auto move_legality = [&]() -> int {
    TRY(do_move(ctx, true));
    if (ctx->dim == 2) TRY(do_legality(ctx));
    CK(ctx->reduce_vec.ensure(16));
    LAUNCH("reduce_vector", k_reduce_vector, 1, 32, ctx->counters.p, ctx->reduce_vec.p);
    return MMB_OK;
  };
  if (which == 0 || which == 2) TRY(marks());
  if (which == 0 || which == 1) TRY(validate_project());
  if (which == 0 || which == 3) TRY(move_legality());
  CK(cudaGetLastError());
  return MMB_OK;
}
int mmb_step_phase_rest(mmb_context* ctx, int ncomp) { return mmb_step_phase(ctx, 0, ncomp); }

// ---- NCCL inside the library ---------------------------------------------------------------------------------------
} // extern "C"
namespace {
struct NcclApi {
  void* h = nullptr; bool ok = false; std::string err;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr; decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr; decltype(&ncclAllReduce) AllReduce = nullptr; decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr; decltype(&ncclGroupStart) GroupStart = nullptr; decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi& nccl_api() {
  static NcclApi a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);          // a process that already uses NCCL (torch) gets that copy
  if (!a.h) { a.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return a; }
#define NCCL_SYM(name) a.name = (decltype(a.name))dlsym(a.h, "nccl" #name); if (!a.name) { a.err = "libnccl.so.2 lacks nccl" #name; return a; }
  NCCL_SYM(GetUniqueId) NCCL_SYM(CommInitRank) NCCL_SYM(CommDestroy) NCCL_SYM(AllReduce) NCCL_SYM(Send) NCCL_SYM(Recv)
  NCCL_SYM(GroupStart) NCCL_SYM(GroupEnd) NCCL_SYM(GetErrorString)
#undef NCCL_SYM
  a.ok = true;
  return a;
}
}  // namespace
#define NK(call)                                                                                         \
  do {                                                                                                   \
    ncclResult_t r_ = (call);                                                                            \
    if (r_ != ncclSuccess) { ctx->err = std::string(#call) + ": " + nccl_api().GetErrorString(r_); return MMB_ERR_NCCL; } \
  } while (0)
extern "C" {

int mmb_comm_unique_id(uint8_t id_out[128]) {
  if (!id_out) return MMB_ERR_INVALID_ARG;
  NcclApi& N = nccl_api();
  if (!N.ok) return MMB_ERR_NCCL;
  ncclUniqueId id;
  if (N.GetUniqueId(&id) != ncclSuccess) return MMB_ERR_NCCL;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id_out, &id, 128);
  return MMB_OK;
}
int mmb_comm_init(mmb_context* ctx, int world, int rank, const uint8_t id_bytes[128]) {
  if (!ctx || !id_bytes || world < 1 || rank < 0 || rank >= world) return MMB_ERR_INVALID_ARG;
  NcclApi& N = nccl_api();
  if (!N.ok) { ctx->err = N.err; return MMB_ERR_NCCL; }
  REQUIRE(!ctx->comm, MMB_ERR_STATE, "mmb_comm_init: the context already has a communicator");
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId id; memcpy(&id, id_bytes, 128);
  NK(N.CommInitRank(&ctx->comm, world, id, rank));
  ctx->world = world; ctx->rank = rank; ctx->sharded = world > 1;
  if (!ctx->s_comm) CK(cudaStreamCreateWithFlags(&ctx->s_comm, cudaStreamNonBlocking));
  while (ctx->comm_ev.size() < 10) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->comm_ev.push_back(e); }
  CK(ctx->reduce_vec.ensure(16));
  CK(cudaMemsetAsync(ctx->reduce_vec.p, 0, 16 * sizeof(double), ctx->stream));
  ctx->graph_valid = false; ctx->sharded_steps = 0;
  return MMB_OK;
}
int mmb_comm_destroy(mmb_context* ctx) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  if (ctx->step_graph) { cudaGraphExecDestroy(ctx->step_graph); ctx->step_graph = nullptr; ctx->graph_valid = false; }
  if (ctx->comm) { cudaStreamSynchronize(ctx->stream); if (ctx->s_comm) cudaStreamSynchronize(ctx->s_comm); nccl_api().CommDestroy(ctx->comm); ctx->comm = nullptr; }
  ctx->sharded = false; ctx->world = 1; ctx->rank = 0;
  return MMB_OK;
}
int mmb_halo_counts(mmb_context* ctx, const int64_t* send_counts, const int64_t* recv_counts) {
  if (!ctx || !send_counts || !recv_counts) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->comm, MMB_ERR_STATE, "mmb_halo_counts before mmb_comm_init");
  int64_t ts = 0, tr = 0;
  for (int p = 0; p < ctx->world; ++p) { REQUIRE(send_counts[p] >= 0 && recv_counts[p] >= 0, MMB_ERR_INVALID_ARG, "negative halo count"); ts += send_counts[p]; tr += recv_counts[p]; }
  REQUIRE(ts == ctx->n_halo_send && tr == ctx->n_halo_recv, MMB_ERR_INVALID_ARG, "halo counts do not add up to the lists of mmb_halo_setup");
  ctx->halo_send_counts.assign(send_counts, send_counts + ctx->world); ctx->halo_recv_counts.assign(recv_counts, recv_counts + ctx->world);
  ctx->graph_valid = false;
  return MMB_OK;
}
// everything one sharded step puts on the two streams (kernels on ctx->stream, collectives on ctx->s_comm)
static int sharded_enqueue(mmb_context* ctx, int ncomp) {
  NcclApi& N = nccl_api();
  cudaStream_t S0 = ctx->stream, Sc = ctx->s_comm;
  cudaEvent_t* ev = ctx->comm_ev.data();
  Counters* c = ctx->counters.p;
  TRY(zero_counters(ctx));
  if (ctx->n_halo_send)
    LAUNCH("halo_pack", k_gather2, grid_for(ctx->n_halo_send, 256), 256, (const double2*)ctx->shift.p, ctx->halo_send_idx.p, ctx->n_halo_send, ctx->halo_send.p);
  CK(cudaEventRecord(ev[0], S0));
  CK(cudaStreamWaitEvent(Sc, ev[0], 0));
  {                                                             // halo exchange of the shifts: grouped point-to-point
    NK(N.GroupStart());
    int64_t so = 0, ro = 0;
    for (int p = 0; p < ctx->world; ++p) {
      const int64_t ns = ctx->halo_send_counts[(size_t)p], nr = ctx->halo_recv_counts[(size_t)p];
      if (ns) NK(N.Send(ctx->halo_send.p + so, (size_t)ns * 2, ncclDouble, p, ctx->comm, Sc));
      if (nr) NK(N.Recv(ctx->halo_recv.p + ro, (size_t)nr * 2, ncclDouble, p, ctx->comm, Sc));
      so += ns; ro += nr;
    }
    NK(N.GroupEnd());
  }
  CK(cudaEventRecord(ev[1], Sc));
  TRY(do_distance(ctx, false));                                 // needs no shifts: runs under the exchange
  CK(cudaEventRecord(ev[2], S0));
  CK(cudaStreamWaitEvent(Sc, ev[2], 0));
  NK(N.AllReduce(&c->max_dist_bits, &c->max_dist_bits, 1, ncclUint64, ncclMax, ctx->comm, Sc));     // non-negative doubles order like integers
  CK(cudaEventRecord(ev[3], Sc));
  CK(cudaStreamWaitEvent(S0, ev[1], 0));
  if (ctx->n_halo_recv)
    LAUNCH("halo_unpack", k_scatter2, grid_for(ctx->n_halo_recv, 256), 256, (double2*)ctx->shift.p, ctx->halo_recv_idx.p, ctx->n_halo_recv, (const double2*)ctx->halo_recv.p);
  TRY(do_validate(ctx, true));
  CK(cudaEventRecord(ev[4], S0));
  CK(cudaStreamWaitEvent(Sc, ev[4], 0));
  NK(N.AllReduce(&c->n_invalid, &c->n_invalid_global, 1, ncclUint64, ncclSum, ctx->comm, Sc));      // the gate of the move, same on all ranks
  CK(cudaEventRecord(ev[5], Sc));
  if (ncomp) TRY(do_project(ctx, ncomp));                       // needs neither the global maximum nor the gate
  CK(cudaStreamWaitEvent(S0, ev[3], 0));
  TRY(do_mark(ctx));
  if (ctx->dim == 2 && ctx->ns)
    LAUNCH("mark_interface2d", k_mark_interface2d, grid_for(ctx->ns, 256), 256, (const double2*)ctx->x.p, (const int2*)ctx->facets.p,
           (int)ctx->ns, dev_params(ctx->prm), ctx->counters.p, ctx->imark.p);
  CK(cudaStreamWaitEvent(S0, ev[5], 0));
  TRY(do_move(ctx, true));
  if (ctx->dim == 2) TRY(do_legality(ctx));
  LAUNCH("reduce_vector", k_reduce_vector, 1, 32, ctx->counters.p, ctx->reduce_vec.p);
  CK(cudaEventRecord(ev[6], S0));
  CK(cudaStreamWaitEvent(Sc, ev[6], 0));
  NK(N.AllReduce(ctx->reduce_vec.p, ctx->reduce_vec.p, 16, ncclDouble, ncclSum, ctx->comm, Sc));
  CK(cudaEventRecord(ev[7], Sc));
  CK(cudaStreamWaitEvent(S0, ev[7], 0));
  CK(cudaGetLastError());
  return MMB_OK;
}
int mmb_adapt_step_sharded(mmb_context* ctx, const double* shifts_host, int ncomp, mmb_step_result* res) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  REQUIRE(ctx->mesh_ok && ctx->comm, MMB_ERR_STATE, "mmb_adapt_step_sharded needs an uploaded shard and mmb_comm_init");
  REQUIRE(ncomp == 0 || ((ncomp == 1 || ncomp == 3) && ctx->dim == 2), MMB_ERR_INVALID_ARG, "ncomp must be 0, 1 (P0) or 3 (DG-P1); projection is 2-D");
  REQUIRE((int64_t)ctx->halo_send_counts.size() == ctx->world && (int64_t)ctx->halo_recv_counts.size() == ctx->world, MMB_ERR_STATE,
          "mmb_adapt_step_sharded before mmb_halo_setup / mmb_halo_counts");
  TRY(stage_shifts(ctx, shifts_host));
  if (ncomp) TRY(stage_dofs(ctx, ncomp, 0, nullptr));
  if (ncomp && ctx->nnew) REQUIRE(ctx->max_old_idx < ctx->n_old, MMB_ERR_INVALID_ARG, "projection: a pair references an old cell beyond the DOF array");
  static const bool use_graph = []() { const char* e = getenv("MMB_GRAPH"); return !(e && atoi(e) == 0); }();
  // The first step runs eagerly: NCCL sets up its point-to-point channels on first use, which must not happen inside a
  // capture.  From the second step on the launch sequence (it depends only on the snapshot, the pairs, ncomp and the
  // parameters) is captured once and replayed.
  if (use_graph && !ctx->timing && ctx->sharded_steps > 0) {
    if (!ctx->graph_valid || ctx->graph_ncomp != ncomp || memcmp(&ctx->graph_prm, &ctx->prm, sizeof(mmb_params)) != 0) {
      if (ctx->step_graph) { cudaGraphExecDestroy(ctx->step_graph); ctx->step_graph = nullptr; }
      cudaGraph_t g = nullptr;
      const int64_t l0 = ctx->launches;
      CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
      const int rc = sharded_enqueue(ctx, ncomp);
      const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
      ctx->graph_launches = ctx->launches - l0; ctx->launches = l0;
      if (rc != MMB_OK) { if (g) cudaGraphDestroy(g); return rc; }
      if (ce != cudaSuccess) { ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); return MMB_ERR_CUDA; }
      CK(cudaGraphInstantiate(&ctx->step_graph, g, 0));
      cudaGraphDestroy(g);
      ctx->graph_valid = true; ctx->graph_ncomp = ncomp; ctx->graph_prm = ctx->prm;
    }
    CK(cudaGraphLaunch(ctx->step_graph, ctx->stream));
    ctx->launches += ctx->graph_launches;
    ctx->dist_valid = true;
  } else {
    TRY(sharded_enqueue(ctx, ncomp));
  }
  ctx->sharded_steps++;
  if (res) {
    double* hv = reinterpret_cast<double*>(ctx->h_counters);     // pinned staging: 16 doubles + the max word fit in Counters
    static_assert(sizeof(Counters) >= 17 * sizeof(double), "staging");
    CK(cudaMemcpyAsync(hv, ctx->reduce_vec.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hv + 16, &ctx->counters.p->max_dist_bits, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    res->n_invalid = (int64_t)hv[0]; res->n_orient_nonpos = (int64_t)hv[1]; res->n_illegal = (int64_t)hv[2];
    res->n_refine = (int64_t)hv[3]; res->n_coarsen = (int64_t)hv[4]; res->n_exact_orient = (int64_t)hv[5];
    res->n_exact_incircle = (int64_t)hv[6]; res->n_pieces = (int64_t)hv[7]; res->mass_new = hv[8]; res->covered_area = hv[9];
    res->max_dist = hv[16];
    const bool withheld = ctx->prm.move_policy == 0 && res->n_invalid > 0;
    res->moved = withheld ? 0 : 1;
    if (ctx->dim == 3 && res->n_invalid > 0) {
      ctx->err = "Vertices could not be moved, because the removal of vertices is not supported in 3D!";
      return MMB_ERR_INVALID_CELL_3D;
    }
    if (withheld) { ctx->err = "moveVertices withheld: the trial move inverts cells (ensureVertexMovement == true); remove the vertices first"; return MMB_MOVE_WITHHELD; }
  }
  return MMB_OK;
}

// ---- timing ---------------------------------------------------------------------------------------------
int mmb_timing_enable(mmb_context* ctx, int on) { if (!ctx) return MMB_ERR_INVALID_ARG; ctx->timing = on != 0; return MMB_OK; }
int mmb_timing_reset(mmb_context* ctx) {
  if (!ctx) return MMB_ERR_INVALID_ARG;
  cudaStreamSynchronize(ctx->stream);
  for (auto& t : ctx->timed) { ctx->event_pool.push_back(t.a); ctx->event_pool.push_back(t.b); }
  ctx->timed.clear();
  return MMB_OK;
}
int mmb_timing_get(mmb_context* ctx, int cap, const char** names, double* total_ms, int64_t* launches) {
  if (!ctx) return 0;
  cudaStreamSynchronize(ctx->stream);
  std::vector<const char*> order; std::map<std::string, std::pair<double, int64_t>> agg; std::map<std::string, const char*> nm;
  for (auto& t : ctx->timed) {
    float ms = 0.f; cudaEventElapsedTime(&ms, t.a, t.b);
    auto it = agg.find(t.name);
    if (it == agg.end()) { agg[t.name] = { ms, 1 }; nm[t.name] = t.name; order.push_back(t.name); }
    else { it->second.first += ms; it->second.second++; }
  }
  int n = 0;
  for (auto* name : order) {
    if (n < cap) { if (names) names[n] = name; if (total_ms) total_ms[n] = agg[name].first; if (launches) launches[n] = agg[name].second; }
    ++n;
  }
  return n;
}

}  // extern "C"
