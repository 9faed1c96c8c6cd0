// kernels.cuh -- sm_100a kernels of the moving-mesh hot path.  Every kernel cites the reference lines whose
// arithmetic it reproduces (paths relative to the reference checkout).  Compile with -fmad=false: the
// reference is an x86-64 build without FMA contraction and marks/flags must be bit-exact.
//
// Layout in HBM: coordinates / shifts as 16-byte double2 per vertex (one LDG.128 per gather),
// cell corners as three int32 planes (SoA; 4 cells per thread through int4 loads), 1-byte flags/marks packed
// 4 per 32-bit store, edges as int4 quads, interface leaves as 48-byte records in Morton order under an
// implicit binary AABB tree.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "predicates.cuh"
#include "clip.cuh"
#include "geom.cuh"
#include "project.cuh"
#include "flip.cuh"

namespace mmb {

struct Counters {
  unsigned long long n_invalid, n_orient_nonpos, n_illegal, n_refine, n_coarsen;
  unsigned long long wl_orient, wl_incircle;      // worklist cursors
  unsigned long long n_exact_orient, n_exact_incircle;
  unsigned long long n_pieces;
  unsigned long long max_dist_bits;               // non-negative double compared as integer
  unsigned long long minmax_bits[4];              // indicator init: iface min, iface max, bulk min, bulk max
  double mass_new, covered_area;
  unsigned long long pad_[2];                     // pad_[0]: total of the last compaction scan; pad_[1]: predicates outside the exact stage's exponent range
  unsigned long long dbg[8];                      // development counters (MMB_DEV_VARIANTS builds only)
  unsigned long long n_invalid_global;            // what gates the fused step's move: n_invalid of ALL ranks (a sharded caller
                                                  // all-reduces it; single GPU: a copy of n_invalid)
  unsigned long long wl_distance;                 // vertices handed to k_distance2d_wide (reset by k_bvh_build2d)
};

// (DevParams: geom.cuh)

__device__ __forceinline__ double2 ldg2(const double2* p) { return __ldg(p); }
__device__ __forceinline__ double4 ldg4(const double4* p) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum of up to 4 small counters packed as separate values; result valid in thread 0
template <int NT>
__device__ __forceinline__ void block_count_add(unsigned a, unsigned b, unsigned long long* ga, unsigned long long* gb) {
  __shared__ unsigned sa[NT / 32], sb[NT / 32];
  a = warp_sum(a); b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = l < NT / 32 ? sa[l] : 0u; b = l < NT / 32 ? sb[l] : 0u;
    a = warp_sum(a); b = warp_sum(b);
    if (l == 0) { if (a) atomicAdd(ga, (unsigned long long)a); if (b) atomicAdd(gb, (unsigned long long)b); }
  }
}
// warp-aggregated append of `item` to a worklist when `pred` holds
__device__ __forceinline__ void wl_push(bool pred, int item, int32_t* wl, unsigned long long* cursor) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) wl[base + __popc(m & ((1u << lane) - 1u))] = item;
}

// ------------------------------------------------------------------------------------------------------
// K6  MMesh::moveVertices (dune/mmesh/grid/mmesh.hh:1421-1430): x <- x (+) s, one IEEE add per component
// ------------------------------------------------------------------------------------------------------
// gate != nullptr: the fused step's move, withheld on the device when the trial validity pass found an inverted cell
// (ensureVertexMovement == true => the reference removes vertices before it moves, mmesh.hh:1109-1134)
__global__ void __launch_bounds__(256) k_move(double2* __restrict__ x, const double2* __restrict__ s, int64_t n2,
                                              const Counters* __restrict__ gate) {
  if (gate != nullptr && gate->n_invalid_global != 0ull) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 a = x[i];
    const double2 b = __ldg(s + i);
    a.x = a.x + b.x; a.y = a.y + b.y;
    x[i] = a;
  }
}

// mmesh.hh:1130: for (GlobalCoordinate& s : shifts) s *= -1.0;  (any scale; exact for +-1)
__global__ void __launch_bounds__(256) k_scale(double2* __restrict__ s, double a, int64_t n2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 v = s[i];
    v.x *= a; v.y *= a;
    s[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------------
// Hilbert tiles.  Cells (and edges) arrive in the host's Hilbert order, so the vertices referenced by a tile of
// TILE consecutive cells cluster in a short index window (98 % inside a TILE-wide window on the 16.8 M benchmark
// mesh).  Each CTA stages that window in shared memory with coalesced 16-byte loads and gathers from there;
// references outside the window fall back to a global (L2) load.  The window start is fixed at upload time:
// median of a 256-index sample of the tile, minus half a window.
// ------------------------------------------------------------------------------------------------------
constexpr int TILE = 1024;          // cells (or edges) per CTA = 256 threads x 4
constexpr int WIN = 1024;           // staged vertices per CTA
__global__ void __launch_bounds__(256)
k_tile_windows(const int32_t* __restrict__ first_index, int stride, int64_t n_items, int64_t nv, int32_t* __restrict__ tile_lo) {
  __shared__ int sv[256];
  const int64_t base = (int64_t)blockIdx.x * TILE;
  const int64_t item = base + 4 * (int64_t)threadIdx.x;            // one sample per thread
  const int a = item < n_items ? first_index[item * stride] : -1;
  sv[threadIdx.x] = a;
  __syncthreads();
  int valid = 0, rank = 0;
  for (int j = 0; j < 256; ++j) {
    const int b = sv[j];
    valid += (b >= 0);
    rank += (b >= 0) && (b < a || (b == a && j < (int)threadIdx.x));
  }
  if (a >= 0 && rank == valid / 2) {
    long long lo = (long long)a - WIN / 2;
    if (lo > nv - WIN) lo = nv - WIN;
    if (lo < 0) lo = 0;
    tile_lo[blockIdx.x] = (int32_t)lo;
  }
}

// ------------------------------------------------------------------------------------------------------
// K1  trial move + validity, 2-D.  mmesh.hh:1094-1134 (ensureVertexMovement): signedVolume_ <= 0.0 on
//     x (+) s in TDS vertex order (mmesh.hh:1169-1172 -> CGAL/Cartesian/function_objects.h:918-926), plus the
//     exact CGAL orientation sign (static filter here, worklist for the fallback kernel).
// ------------------------------------------------------------------------------------------------------
template <bool SHIFT>
__global__ void __launch_bounds__(256)
k_validate2d(const double2* __restrict__ xy, const double2* __restrict__ sh, const int4* __restrict__ cv0,
             const int4* __restrict__ cv1, const int4* __restrict__ cv2, const int32_t* __restrict__ tile_lo, int64_t nv,
             int64_t nc4, int64_t nc, int64_t own_lo, int64_t own_hi, uint32_t* __restrict__ valid4,
             uint32_t* __restrict__ orient4, int32_t* __restrict__ wl, Counters* __restrict__ cnt) {
  __shared__ double2 sxy[WIN];                              // trial coordinates x (+) s of the tile's vertex window
  const int lo = tile_lo[blockIdx.x];
#pragma unroll
  for (int j = threadIdx.x; j < WIN; j += 256) {
    const int v = lo + j;
    if (v < nv) {
      double2 p = ldg2(xy + v);
      if (SHIFT) { const double2 q = ldg2(sh + v); p.x = p.x + q.x; p.y = p.y + q.y; }   // mmesh.hh:1427 point = center() + shift
      sxy[j] = p;
    }
  }
  __syncthreads();
  auto fetch = [&](int v) -> double2 {
    const unsigned j = (unsigned)(v - lo);
    if (j < (unsigned)WIN) return sxy[j];
    double2 p = ldg2(xy + v);
    if (SHIFT) { const double2 q = ldg2(sh + v); p.x = p.x + q.x; p.y = p.y + q.y; }
    return p;
  };
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned ninv = 0, nnonpos = 0;
  bool need[4] = { false, false, false, false };
  if (g < nc4) {
    const int4 a = __ldg(cv0 + g), b = __ldg(cv1 + g), c = __ldg(cv2 + g);
    const int va[4] = { a.x, a.y, a.z, a.w }, vb[4] = { b.x, b.y, b.z, b.w }, vc[4] = { c.x, c.y, c.z, c.w };
    uint32_t vbits = 0, obits = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double2 P = fetch(va[j]), Q = fetch(vb[j]), R = fetch(vc[j]);
      double det;
      int s = orient2d_filter(P.x, P.y, Q.x, Q.y, R.x, R.y, det);
      const double area = det / 2;                           // Compute_area_2: determinant(...)/2
      const bool ok = !(area <= 0.0);                        // mmesh.hh:1102
      const bool live = 4 * g + j < nc;
      const bool own = 4 * g + j >= own_lo && 4 * g + j < own_hi;      // halo cells are evaluated but not counted
      need[j] = live && (s == SIGN_UNKNOWN);
      if (s == SIGN_UNKNOWN) s = 0;
      vbits |= (uint32_t)(ok ? 1u : 0u) << (8 * j);
      obits |= ((uint32_t)(uint8_t)(int8_t)s) << (8 * j);
      ninv += (own && !ok);
      nnonpos += (own && !need[j] && s <= 0);
    }
    valid4[g] = vbits;
    orient4[g] = obits;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) wl_push(need[j], (int)(4 * g + j), wl, &cnt->wl_orient);
  block_count_add<256>(ninv, nnonpos, &cnt->n_invalid, &cnt->n_orient_nonpos);
}

// exact fallback for the cells the static filter could not decide (interval, then expansions)
template <bool SHIFT>
__global__ void __launch_bounds__(128)
k_exact_orient2d(const double2* __restrict__ xy, const double2* __restrict__ sh, const int32_t* __restrict__ cv0,
                 const int32_t* __restrict__ cv1, const int32_t* __restrict__ cv2, const int32_t* __restrict__ wl,
                 int64_t own_lo, int64_t own_hi, int8_t* __restrict__ orient, Counters* __restrict__ cnt, int publish) {
  // the validity pass is complete (stream order): single-GPU contexts publish their own count as the gate of the move
  if (publish && blockIdx.x == 0 && threadIdx.x == 0) cnt->n_invalid_global = cnt->n_invalid;
  const unsigned long long n = cnt->wl_orient;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned nnonpos = 0, nexact = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = wl[i];
    double2 P = ldg2(xy + cv0[c]), Q = ldg2(xy + cv1[c]), R = ldg2(xy + cv2[c]);
    if (SHIFT) {
      const double2 sp = ldg2(sh + cv0[c]), sq = ldg2(sh + cv1[c]), sr = ldg2(sh + cv2[c]);
      P.x = P.x + sp.x; P.y = P.y + sp.y; Q.x = Q.x + sq.x; Q.y = Q.y + sq.y; R.x = R.x + sr.x; R.y = R.y + sr.y;
    }
    int s = orient2d_interval(P.x, P.y, Q.x, Q.y, R.x, R.y);
    if (s == SIGN_UNKNOWN) { s = orient2d_exact(P.x, P.y, Q.x, Q.y, R.x, R.y); nexact++; }
    if (s == SIGN_DOMAIN) { s = 0; atomicAdd(&cnt->pad_[1], 1ull); }
    orient[c] = (int8_t)s;
    nnonpos += (s <= 0 && c >= own_lo && c < own_hi);
  }
  block_count_add<128>(nnonpos, nexact, &cnt->n_orient_nonpos, &cnt->n_exact_orient);
}

// ------------------------------------------------------------------------------------------------------
// K2  per-edge Delaunay legality: illegal <=> side_of_oriented_circle(a,b,c; d) == ON_POSITIVE_SIDE
//     (CGAL/Delaunay_triangulation_2.h:663-681; predicate CGAL/internal/Static_filters/Side_of_oriented_circle_2.h)
// ------------------------------------------------------------------------------------------------------
template <bool SHIFT>
__global__ void __launch_bounds__(256)
k_legality(const double2* __restrict__ xy, const double2* __restrict__ sh, const int4* __restrict__ edges, int64_t ne,
           const int32_t* __restrict__ tile_lo, int64_t nv, uint8_t* __restrict__ flag, int32_t* __restrict__ wl,
           Counters* __restrict__ cnt, int gated) {
  __shared__ double2 sxy[WIN];
  const int lo = tile_lo[blockIdx.x];
  // SHIFT evaluates the flags on x (+) shift, the coordinates the step's move is about to write; when that move is
  // withheld (gated, an inverted cell was found) the flags describe the mesh as it stays: x
  const bool shifted = SHIFT && !(gated && cnt->n_invalid_global != 0ull);
#pragma unroll
  for (int j = threadIdx.x; j < WIN; j += 256) {
    const int v = lo + j;
    if (v < nv) {
      double2 p = ldg2(xy + v);
      if (shifted) { const double2 q = ldg2(sh + v); p.x = p.x + q.x; p.y = p.y + q.y; }
      sxy[j] = p;
    }
  }
  __syncthreads();
  auto fetch = [&](int v) -> double2 {
    const unsigned j = (unsigned)(v - lo);
    if (j < (unsigned)WIN) return sxy[j];
    double2 p = ldg2(xy + v);
    if (shifted) { const double2 q = ldg2(sh + v); p.x = p.x + q.x; p.y = p.y + q.y; }
    return p;
  };
  unsigned nill = 0;
  const int64_t e0 = (int64_t)blockIdx.x * TILE + threadIdx.x;       // edges e0, e0+256, e0+512, e0+768: coalesced int4 loads
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t e = e0 + 256 * r;
    bool need = false;
    if (e < ne) {
      const int4 q = __ldg(edges + e);
      uint8_t f = 2;
      if (q.w >= 0) {
        const double2 A = fetch(q.x), B = fetch(q.y), C = fetch(q.z), D = fetch(q.w);
        const int s = incircle_filter(A.x, A.y, B.x, B.y, C.x, C.y, D.x, D.y);
        need = (s == SIGN_UNKNOWN);
        f = (s > 0 && !need) ? 0 : 1;
        nill += (f == 0);
      }
      flag[e] = f;
    }
    wl_push(need, (int)e, wl, &cnt->wl_incircle);
  }
  block_count_add<256>(nill, 0u, &cnt->n_illegal, &cnt->n_illegal);
}
template <bool SHIFT>
__global__ void __launch_bounds__(128)
k_exact_incircle(const double2* __restrict__ xy, const double2* __restrict__ sh, const int4* __restrict__ edges,
                 const int32_t* __restrict__ wl, uint8_t* __restrict__ flag, Counters* __restrict__ cnt, int gated) {
  const unsigned long long n = cnt->wl_incircle;
  const bool shifted = SHIFT && !(gated && cnt->n_invalid_global != 0ull);
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned nill = 0, nexact = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int e = wl[i];
    const int4 q = __ldg(edges + e);
    double2 A = ldg2(xy + q.x), B = ldg2(xy + q.y), C = ldg2(xy + q.z), D = ldg2(xy + q.w);
    if (shifted) {
      const double2 sa = ldg2(sh + q.x), sb = ldg2(sh + q.y), sc = ldg2(sh + q.z), sd = ldg2(sh + q.w);
      A.x = A.x + sa.x; A.y = A.y + sa.y; B.x = B.x + sb.x; B.y = B.y + sb.y;
      C.x = C.x + sc.x; C.y = C.y + sc.y; D.x = D.x + sd.x; D.y = D.y + sd.y;
    }
    int s = incircle_interval(A.x, A.y, B.x, B.y, C.x, C.y, D.x, D.y);
    if (s == SIGN_UNKNOWN) { s = incircle_exact(A.x, A.y, B.x, B.y, C.x, C.y, D.x, D.y); nexact++; }
    if (s == SIGN_DOMAIN) { s = 0; atomicAdd(&cnt->pad_[1], 1ull); }
    flag[e] = (uint8_t)(s > 0 ? 0 : 1);
    nill += (s > 0);
  }
  block_count_add<128>(nill, nexact, &cnt->n_illegal, &cnt->n_exact_incircle);
}

// raw predicate batches (config 4 stress): which 0 orient2d, 1 incircle, 2 orient3d; two-phase like the mesh kernels
__global__ void __launch_bounds__(256)
k_pred_filter(int which, const double* __restrict__ pts, int64_t n, int8_t* __restrict__ sign, int32_t* __restrict__ wl,
              Counters* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool need = false;
  if (i < n) {
    int s; double det;
    if (which == 0) { const double* p = pts + 6 * i; s = orient2d_filter(p[0], p[1], p[2], p[3], p[4], p[5], det); }
    else if (which == 1) { const double* p = pts + 8 * i; s = incircle_filter(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]); }
    else { const double* p = pts + 12 * i; s = orient3d_filter(p, p + 3, p + 6, p + 9, det); }
    need = (s == SIGN_UNKNOWN);
    sign[i] = (int8_t)(need ? 0 : s);
  }
  wl_push(need, (int)i, wl, &cnt->wl_orient);
}
__global__ void __launch_bounds__(128)
k_pred_exact(int which, const double* __restrict__ pts, const int32_t* __restrict__ wl, int8_t* __restrict__ sign,
             Counters* __restrict__ cnt) {
  const unsigned long long n = cnt->wl_orient;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned nexact = 0;
  for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const int64_t i = wl[k];
    int s;
    if (which == 0) {
      const double* p = pts + 6 * i;
      s = orient2d_interval(p[0], p[1], p[2], p[3], p[4], p[5]);
      if (s == SIGN_UNKNOWN) { s = orient2d_exact(p[0], p[1], p[2], p[3], p[4], p[5]); nexact++; }
    } else if (which == 1) {
      const double* p = pts + 8 * i;
      s = incircle_interval(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
      if (s == SIGN_UNKNOWN) { s = incircle_exact(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]); nexact++; }
    } else {
      const double* p = pts + 12 * i;
      s = orient3d_interval(p, p + 3, p + 6, p + 9);
      if (s == SIGN_UNKNOWN) { s = orient3d_exact(p, p + 3, p + 6, p + 9); nexact++; }
    }
    if (s == SIGN_DOMAIN) { s = 0; atomicAdd(&cnt->pad_[1], 1ull); }
    sign[i] = (int8_t)s;
  }
  block_count_add<128>(nexact, 0u, &cnt->n_exact_orient, &cnt->n_exact_orient);
}

// ------------------------------------------------------------------------------------------------------
// K3  interface distance field, 2-D.  Per-pair arithmetic = dune/mmesh/remeshing/distance.hh:143-156 and
//     Plane (dune/mmesh/cgal/triangulationwrapper.hh:14-40); min over ALL facets (distance.hh:42-59,133-140) is
//     obtained with an AABB tree whose culling is conservative, so the minimum has the brute-force bits.
// ------------------------------------------------------------------------------------------------------
// (SegLeaf, seg_leaf_make, seg_distance: geom.cuh)

// AABB tree over the facets in Morton order of their (upload-time) midpoints: implicit heap, node i in [1, 2L),
// children 2i / 2i+1, leaves at L + s.  Boxes are float4 (minx, miny, maxx, maxy) rounded OUTWARD, so every fp32
// test below is a rigorous lower bound; the per-facet arithmetic stays fp64 and identical to the reference.
__device__ __forceinline__ float4 box_merge(float4 a, float4 b) {
  return make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
// One launch refits the whole tree: each block reduces 256 leaves in shared memory (8 levels); the block that
// finishes last (atomic ticket) reduces the remaining top levels.
__global__ void __launch_bounds__(256)
k_bvh_build2d(const double2* __restrict__ xy, const int2* __restrict__ segs_sorted, int ns, int L,
              SegLeaf* __restrict__ leaf, float4* __restrict__ box, unsigned* __restrict__ ticket, Counters* __restrict__ cnt) {
  __shared__ float4 sb[512];
  __shared__ bool last;
  const int t = threadIdx.x;
  if (blockIdx.x == 0 && t == 0) cnt->wl_distance = 0ull;   // cursor of the list k_distance2d fills for k_distance2d_wide
  const int W = L < 256 ? L : 256;                       // leaves per block (power of two)
  const int i = blockIdx.x * W + t;
  float4 b = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
  if (t < W && i < ns) {
    const int2 s = segs_sorted[i];
    const double2 v0 = ldg2(xy + s.x), v1 = ldg2(xy + s.y);
    const SegLeaf f = seg_leaf_make(v0.x, v0.y, v1.x, v1.y);
    leaf[i] = f;
    b = make_float4(__double2float_rd(fmin(f.c0x, f.c1x)), __double2float_rd(fmin(f.c0y, f.c1y)),
                    __double2float_ru(fmax(f.c0x, f.c1x)), __double2float_ru(fmax(f.c0y, f.c1y)));
  }
  if (t < W) { sb[W + t] = b; box[L + i] = b; }
  const int g0 = L / W + blockIdx.x;                     // global index of this block's subtree root
  for (int w = W >> 1, d = 1; w >= 1; w >>= 1, ++d) {
    __syncthreads();
    if (t < w) {
      const float4 m = box_merge(sb[2 * (w + t)], sb[2 * (w + t) + 1]);
      sb[w + t] = m;
      // local node (w + t) sits `lev` levels below the subtree root, offset t inside its level
      int lev = 0; for (int x = w; x > 1; x >>= 1) ++lev;
      box[((size_t)g0 << lev) + t] = m;
    }
  }
  if (L <= W) return;                                    // single block: done (root written above)
  __threadfence();
  __syncthreads();
  if (t == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  const int nb = L / W;                                  // top tree: leaves are the nb subtree roots at [nb, 2nb)
  for (int w = nb >> 1; w >= 1; w >>= 1) {
    for (int j = w + t; j < 2 * w; j += 256) {
      const float4 a = __ldcg(box + 2 * j), c = __ldcg(box + 2 * j + 1);
      box[j] = box_merge(a, c);
    }
    __threadfence_block();
    __syncthreads();
  }
  if (t == 0) *ticket = 0;
}
struct DistQuery {
  double px, py, best, diag;
  float pxlo, pxhi, pylo, pyhi, T2f;
  int arg;
  // conservative cull threshold: a node whose box distance exceeds T cannot hold a facet whose *computed* distance
  // is below `best` (computed and true distance differ by O(1e-15 (d + len)); diag >= any facet length)
  __device__ __forceinline__ void set_threshold() {
    const double T = best + 1e-12 * (best + diag);
    T2f = __double2float_ru(T * T * (1.0 + 1e-12));
  }
  // rigorous fp32 lower bound of the squared distance from the query point to a (outward-rounded) box
  __device__ __forceinline__ float lb2(const float4 b) const {
    const float dx = fmaxf(fmaxf(__fsub_rd(b.x, pxhi), __fsub_rd(pxlo, b.z)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rd(b.y, pyhi), __fsub_rd(pylo, b.w)), 0.f);
    return __fadd_rd(__fmul_rd(dx, dx), __fmul_rd(dy, dy));
  }
  __device__ __forceinline__ void eval_leaf(const SegLeaf* __restrict__ leaf, int s) {
    const SegLeaf f = leaf[s];
    const double d = seg_distance(f, px, py);
    if (d < best) { best = d; arg = s; set_threshold(); }          // distance.hh:137-138
  }
  // nearest-first depth-first search of the subtree rooted at `root`
  __device__ __forceinline__ void dfs(const SegLeaf* __restrict__ leaf, const float4* __restrict__ box, int ns, int L, int root) {
    int stack[32]; float sd2[32]; int sp = 0;
    int node = root;
    for (;;) {
      if (node >= L) {
        if (node - L < ns) eval_leaf(leaf, node - L);
        node = 0;
      } else {
        const float dl = lb2(__ldg(box + 2 * node)), dr = lb2(__ldg(box + 2 * node + 1));
        const bool lf = dl <= dr;
        const float dn = lf ? dl : dr, df = lf ? dr : dl;
        const int nn = 2 * node + (lf ? 0 : 1), nf = 2 * node + (lf ? 1 : 0);
        if (dn > T2f) node = 0;
        else { if (!(df > T2f)) { stack[sp] = nf; sd2[sp] = df; ++sp; } node = nn; }
      }
      if (node == 0) {
        while (sp > 0) { --sp; if (!(sd2[sp] > T2f)) { node = stack[sp]; break; } }
        if (node == 0) break;
      }
    }
  }
};
// Per warp of 32 Hilbert-consecutive vertices (a compact patch of the mesh):
//   1. every lane evaluates its hint facet (last step's nearest facet; on a cold start the warp leader's full search seeds
//      the others) -> an upper bound per lane; the warp keeps the largest one and the bounding box of its 32 points;
//   2. the warp walks the tree TOGETHER, level by level (lane pairs test the two children of 16 frontier nodes at a time),
//      keeping the nodes whose box can be closer to the warp's box than the warp's bound, down to the level whose nodes
//      hold DIST_GROUP leaves: a short list of candidate groups in shared memory;
//   3. every lane runs the same flat loop over that list with its own, much tighter bound: group box, then leaf boxes,
//      then the exact reference arithmetic for the few facets that survive.
// Culling is conservative at both levels (outward-rounded fp32 boxes, bounds rounded up, a point's distance to a box is
// never below the distance between the warp's box and that box), so the result is the minimum over ALL facets with the
// brute-force bits, like distance.hh:42-59.  A frontier that outgrows the list (bounds far larger than the facet
// spacing: no usable hint anywhere in the warp) falls back to the per-lane walk: hint leaf -> root, each sibling subtree
// searched once.
constexpr int DIST_GROUP_LOG = 3;            // candidate groups of 8 leaves
constexpr int DIST_GROUP = 1 << DIST_GROUP_LOG;
constexpr int DIST_QCAP = 128;               // frontier / candidate capacity per warp
__device__ __forceinline__ void dist_lane_walk(DistQuery& q, const SegLeaf* __restrict__ leaf, const float4* __restrict__ box,
                                               int ns, int L, int h) {
  for (int node = L + h; node > 1; node >>= 1) {
    const int sib = node ^ 1;
    if (!(q.lb2(__ldg(box + sib)) > q.T2f)) q.dfs(leaf, box, ns, L, sib);
  }
}
// Cold start (no hints yet: the first Distance::update after a snapshot upload): a coarse grid over the interface's bounding
// box, every grid cell holding the facet nearest to its centre (one full search per grid cell, 32 k searches).  A vertex
// takes its grid cell's facet as the first upper bound -- any facet is a valid one, the search below makes it exact.
constexpr int SEED_GX = 256, SEED_GY = 128;
__global__ void __launch_bounds__(128)
k_distance_seeds(const SegLeaf* __restrict__ leaf, const float4* __restrict__ box, int ns, int L, int32_t* __restrict__ seed) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= SEED_GX * SEED_GY) return;
  const float4 root = __ldg(box + 1);
  DistQuery q;
  q.px = (double)root.x + ((double)root.z - (double)root.x) * (((g % SEED_GX) + 0.5) / SEED_GX);
  q.py = (double)root.y + ((double)root.w - (double)root.y) * (((g / SEED_GX) + 0.5) / SEED_GY);
  q.pxlo = __double2float_rd(q.px); q.pxhi = __double2float_ru(q.px);
  q.pylo = __double2float_rd(q.py); q.pyhi = __double2float_ru(q.py);
  q.diag = ((double)root.z - (double)root.x) + ((double)root.w - (double)root.y);
  q.best = 1e100; q.arg = 0;
  q.set_threshold();
  q.dfs(leaf, box, ns, L, 1);
  seed[g] = q.arg;
}
__global__ void __launch_bounds__(256, 5)
k_distance2d(const double2* __restrict__ xy, int64_t nv, const SegLeaf* __restrict__ leaf,
             const float4* __restrict__ box, int ns, int L, int logL, double* __restrict__ dist, int32_t* __restrict__ hint,
             const int32_t* __restrict__ seed, int32_t* __restrict__ wide, Counters* __restrict__ cnt) {
  __shared__ int s_front[8][2][DIST_QCAP];
  __shared__ float4 s_gbox[8][DIST_QCAP];                 // boxes of the candidate groups (written by the lane that tested them)
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool live = v < nv && ns > 0;
  DistQuery q;
  q.best = 1e100; q.arg = -1;
  q.px = q.py = 0.0; q.diag = 0.0; q.pxlo = q.pylo = INFINITY; q.pxhi = q.pyhi = -INFINITY; q.T2f = 0.f;
  if (live) {
    const double2 p = ldg2(xy + v);
    q.px = p.x; q.py = p.y;
    q.pxlo = __double2float_rd(p.x); q.pxhi = __double2float_ru(p.x);
    q.pylo = __double2float_rd(p.y); q.pyhi = __double2float_ru(p.y);
    const float4 root = __ldg(box + 1);
    q.diag = ((double)root.z - (double)root.x) + ((double)root.w - (double)root.y);
    q.set_threshold();
  }
  int h = live ? hint[v] : 0;
  if (h >= ns) h = -1;
  if (live && h < 0) {                                   // cold start: the facet nearest to the centre of the vertex's seed-grid cell
    int sd = 0;
    if (seed != nullptr) {
      const float4 root = __ldg(box + 1);
      const double fx = (q.px - (double)root.x) / ((double)root.z - (double)root.x), fy = (q.py - (double)root.y) / ((double)root.w - (double)root.y);
      const int gx = min(max((int)(fx * SEED_GX), 0), SEED_GX - 1), gy = min(max((int)(fy * SEED_GY), 0), SEED_GY - 1);   // NaN (degenerate box) -> 0
      sd = seed[gy * SEED_GX + gx];
    }
    h = min(max(sd, 0), ns - 1);
  }
  if (live) q.eval_leaf(leaf, h);
  // the warp's bound and box (dead lanes are neutral: T2f = 0, empty box)
  float U2 = live ? q.T2f : 0.f, bxlo = q.pxlo, bxhi = q.pxhi, bylo = q.pylo, byhi = q.pyhi;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    U2 = fmaxf(U2, __shfl_xor_sync(0xffffffffu, U2, o));
    bxlo = fminf(bxlo, __shfl_xor_sync(0xffffffffu, bxlo, o)); bxhi = fmaxf(bxhi, __shfl_xor_sync(0xffffffffu, bxhi, o));
    bylo = fminf(bylo, __shfl_xor_sync(0xffffffffu, bylo, o)); byhi = fmaxf(byhi, __shfl_xor_sync(0xffffffffu, byhi, o));
  }
  const int stop = logL - DIST_GROUP_LOG;                 // level whose nodes hold DIST_GROUP leaves
  bool deferred = false;
  bool flat = __any_sync(0xffffffffu, live) && stop >= 1;
  // the walk starts with the 8 nodes of level 3 as its frontier (one test of their 16 children replaces four levels of
  // one- to eight-node iterations; a subtree without facets has an empty box and drops out at once)
  const int lev0 = stop > 3 ? 3 : 0;
  int ncur = 1 << lev0, cb = 0;
  if (flat) {
    if (lane < ncur) s_front[w][0][lane] = ncur + lane;
    __syncwarp();
    for (int lev = lev0; lev < stop && flat; ++lev) {
      int nnext = 0;
      for (int base = 0; base < ncur; base += 16) {
        const int k = base + (lane >> 1);
        bool pass = false; int child = 0;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < ncur) {
          child = 2 * s_front[w][cb][k] + (lane & 1);
          b = __ldg(box + child);
          const float dx = fmaxf(fmaxf(__fsub_rd(b.x, bxhi), __fsub_rd(bxlo, b.z)), 0.f);
          const float dy = fmaxf(fmaxf(__fsub_rd(b.y, byhi), __fsub_rd(bylo, b.w)), 0.f);
          pass = !(__fadd_rd(__fmul_rd(dx, dx), __fmul_rd(dy, dy)) > U2);
        }
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        const int pos = nnext + __popc(m & ((1u << lane) - 1u));
        if (pass && pos < DIST_QCAP) { s_front[w][cb ^ 1][pos] = child; s_gbox[w][pos] = b; }
        nnext += __popc(m);
      }
      __syncwarp();
      if (nnext > DIST_QCAP) flat = false;                // frontier outgrew the list: per-lane walk instead
      ncur = nnext; cb ^= 1;
    }
  }
#ifdef MMB_DEV_VARIANTS
  unsigned dbg_g = 0, dbg_l = 0, dbg_e = 0;
#endif
  if (flat) {
    if (live) {
      for (int k = 0; k < ncur; ++k) {
        if (q.lb2(s_gbox[w][k]) > q.T2f) continue;
#ifdef MMB_DEV_VARIANTS
        if (lane == __ffs(__activemask()) - 1) dbg_g++;
#endif
        const int s0 = (s_front[w][cb][k] << DIST_GROUP_LOG) - L;
        // the group's leaf boxes in one batch (independent loads), then the exact arithmetic for the survivors only
        const float4* lb = box + L + s0;
        unsigned need = 0;
#pragma unroll
        for (int c = 0; c < DIST_GROUP; ++c) need |= (!(q.lb2(__ldg(lb + c)) > q.T2f) ? 1u : 0u) << c;
        while (need) {
          const int c = __ffs(need) - 1; need &= need - 1;
          const int s = s0 + c;
          // (the bound has shrunk since the batch test only if an earlier facet of this group improved it: re-test is not
          // needed for correctness, an extra exact evaluation never changes the minimum)
          if (s < ns && s != h) {
#ifdef MMB_DEV_VARIANTS
            if (lane == __ffs(__activemask()) - 1) dbg_e++;
            dbg_l++;
#endif
            q.eval_leaf(leaf, s);
          }
        }
      }
    }
  } else if (wide != nullptr && stop >= 1) {
    // the candidate list outgrew its slots: the bounds of this warp are far larger than the facet spacing (a patch about
    // equally far from a long stretch of interface -- the centre of a closed curve -- or no usable hint).  Its vertices go
    // to k_distance2d_wide, one WARP per vertex; here they contribute nothing (distance, hint and maximum are written there)
    const unsigned m = __ballot_sync(0xffffffffu, live);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&cnt->wl_distance, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) wide[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)v;
    q.best = 0.0;
    deferred = true;
  } else if (live) {
    dist_lane_walk(q, leaf, box, ns, L, h);
  }
  if (live && !deferred) hint[v] = q.arg;
  if (v < nv && !deferred) dist[v] = q.best;
#ifdef MMB_DEV_VARIANTS
  {  // per warp: [0] warps, [1] flat warps, [2] sum of candidate groups, [3] group passes (warp union), [4] leaf evals (warp union), [5] leaf evals (lanes)
    const unsigned gs = warp_sum(dbg_g), es = warp_sum(dbg_e), ls = warp_sum(dbg_l);
    if (lane == 0) {
      atomicAdd(&cnt->dbg[0], 1ull); atomicAdd(&cnt->dbg[1], flat ? 1ull : 0ull); atomicAdd(&cnt->dbg[2], flat ? (unsigned long long)ncur : 0ull);
      atomicAdd(&cnt->dbg[3], (unsigned long long)gs); atomicAdd(&cnt->dbg[4], (unsigned long long)es); atomicAdd(&cnt->dbg[5], (unsigned long long)ls);
    }
  }
#endif
  // Distance::maximum (distance.hh:118-123): max starting from 0.0; non-negative doubles order like integers
  unsigned long long bits = (v < nv) ? (unsigned long long)__double_as_longlong(q.best) : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, bits, o); bits = t > bits ? t : bits; }
  __shared__ unsigned long long sm[8];
  if (lane == 0) sm[w] = bits;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long m = sm[0];
#pragma unroll
    for (int ww = 1; ww < 8; ++ww) m = sm[ww] > m ? sm[ww] : m;
    atomicMax(&cnt->max_dist_bits, m);
  }
}

// The vertices k_distance2d could not serve from a short candidate list, one warp per vertex (grid-stride over the list):
// the 32 lanes test the boxes of 32 leaf groups at a time against the vertex's bound, then take the passing groups four at a
// time -- one leaf per lane, box test, exact reference arithmetic -- and tighten the common bound with a warp minimum.  The
// work of such a vertex (every facet of a closed curve seen from its centre is about equally far) is spread over a warp and
// the vertices over the whole grid, instead of 32 lanes walking the tree one after the other.  Same conservative culls, same
// per-facet arithmetic: the minimum over all facets with the brute-force bits.
__global__ void __launch_bounds__(256, 4)
k_distance2d_wide(const double2* __restrict__ xy, const SegLeaf* __restrict__ leaf, const float4* __restrict__ box, int ns, int L,
                  double* __restrict__ dist, int32_t* __restrict__ hint, const int32_t* __restrict__ seed,
                  const int32_t* __restrict__ wide, Counters* __restrict__ cnt) {
  const unsigned long long n = cnt->wl_distance;
  const int lane = threadIdx.x & 31;
  const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  const int ng = (ns + DIST_GROUP - 1) >> DIST_GROUP_LOG, G0 = L >> DIST_GROUP_LOG;
  unsigned long long mx = 0ull;
  for (unsigned long long i = warp0; i < n; i += nwarps) {
    const int64_t v = wide[i];
    DistQuery q;
    const double2 p = ldg2(xy + v);
    q.px = p.x; q.py = p.y;
    q.pxlo = __double2float_rd(p.x); q.pxhi = __double2float_ru(p.x);
    q.pylo = __double2float_rd(p.y); q.pyhi = __double2float_ru(p.y);
    const float4 root = __ldg(box + 1);
    q.diag = ((double)root.z - (double)root.x) + ((double)root.w - (double)root.y);
    q.best = 1e100; q.arg = -1; q.set_threshold();
    int h = hint[v];
    if (h >= ns) h = -1;
    if (h < 0) {
      int sd = 0;
      if (seed != nullptr) {
        const double fx = (q.px - (double)root.x) / ((double)root.z - (double)root.x), fy = (q.py - (double)root.y) / ((double)root.w - (double)root.y);
        const int gx = min(max((int)(fx * SEED_GX), 0), SEED_GX - 1), gy = min(max((int)(fy * SEED_GY), 0), SEED_GY - 1);
        sd = seed[gy * SEED_GX + gx];
      }
      h = min(max(sd, 0), ns - 1);
    }
    q.eval_leaf(leaf, h);                                 // every lane holds the same bound from here on
    // two levels: the boxes of 32 "super groups" (tree nodes five levels above the groups, 32 groups each) per round, then
    // the groups of the super groups that pass.  A tree with fewer than 256 leaves has no such level: all groups then.
    const bool two_level = G0 >= 32;
    const int S0 = G0 >> 5, nsg = two_level ? (ng + 31) >> 5 : 1;
    for (int sb = 0; sb < nsg; sb += 32) {
      unsigned ms = 1u;
      if (two_level) { const int j = sb + lane; ms = __ballot_sync(0xffffffffu, j < nsg && !(q.lb2(__ldg(box + S0 + j)) > q.T2f)); }
      while (ms) {
        const int sj = __ffs(ms) - 1; ms &= ms - 1;
        const int b0 = two_level ? (sb + sj) << 5 : 0, b1 = two_level ? min(b0 + 32, ng) : ng;
        for (int base = b0; base < b1; base += 32) {
          const int g = base + lane;
          unsigned m = __ballot_sync(0xffffffffu, g < b1 && !(q.lb2(__ldg(box + G0 + g)) > q.T2f));
          while (m) {
            int gsel = -1;                                  // up to four passing groups, eight lanes each
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              const int b = m ? __ffs(m) - 1 : -1;
              if (m) m &= m - 1;
              if ((lane >> DIST_GROUP_LOG) == sl) gsel = b;
            }
            double d = 1e300; int s = -1;
            if (gsel >= 0) {
              s = ((base + gsel) << DIST_GROUP_LOG) + (lane & (DIST_GROUP - 1));
              if (s < ns && !(q.lb2(__ldg(box + L + s)) > q.T2f)) d = seg_distance(leaf[s], q.px, q.py); else s = -1;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {              // warp minimum, lowest facet index among equal distances
              const double od = __shfl_xor_sync(0xffffffffu, d, o); const int os = __shfl_xor_sync(0xffffffffu, s, o);
              if (od < d || (od == d && os >= 0 && (s < 0 || os < s))) { d = od; s = os; }
            }
            if (s >= 0 && d < q.best) { q.best = d; q.arg = s; q.set_threshold(); }
          }
        }
      }
    }
    if (lane == 0) { dist[v] = q.best; hint[v] = q.arg; }
    const unsigned long long bits = (unsigned long long)__double_as_longlong(q.best);
    mx = bits > mx ? bits : mx;
  }
  if (lane == 0 && mx) atomicMax(&cnt->max_dist_bits, mx);
}

// ------------------------------------------------------------------------------------------------------
// K4  RatioIndicator::operator() + LongestEdgeRefinement::refinement argmax, 2-D bulk cells.
//     dune/mmesh/remeshing/ratioindicator.hh:106-160, distance.hh:92-101, longestedgerefinement.hh:41-57,
//     circumcenter CGAL/constructions/kernel_ftC2.h:40-74, geometry dune/mmesh/grid/geometry.hh:76-104
//     (+ AffineGeometry of dune-geometry).  One evaluation per cell (the reference evaluates twice,
//     mmesh.hh:741-742).
// ------------------------------------------------------------------------------------------------------
// (edge_len2d, indicator2d: geom.cuh)
__global__ void __launch_bounds__(256)
k_mark2d(const double2* __restrict__ xy, const double* __restrict__ dist, const int32_t* __restrict__ cv0,
         const int32_t* __restrict__ cv1, const int32_t* __restrict__ cv2, const uint8_t* __restrict__ perm, int64_t nc,
         int64_t own_lo, int64_t own_hi, DevParams prm, const Counters* __restrict__ cin, int8_t* __restrict__ mark,
         uint8_t* __restrict__ longest, Counters* __restrict__ cnt) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned nref = 0, ncoa = 0;
  if (c < nc) {
    // RatioIndicator::update (ratioindicator.hh:94-97): maxDist_ = distProportion_ * distance_.maximum()
    const double maxDist = prm.distProportion * __longlong_as_double((long long)cin->max_dist_bits);
    const int v[3] = { __ldg(cv0 + c), __ldg(cv1 + c), __ldg(cv2 + c) };
    const unsigned p = __ldg(perm + c);
    const int k0 = v[p & 3], k1 = v[(p >> 2) & 3], k2 = v[(p >> 4) & 3];          // id-sorted corners
    const double2 c0 = ldg2(xy + k0), c1 = ldg2(xy + k1), c2 = ldg2(xy + k2);
    const double d0 = __ldg(dist + k0), d1 = __ldg(dist + k1), d2 = __ldg(dist + k2);
    int le;
    const int m = indicator2d(c0, c1, c2, d0, d1, d2, prm, maxDist, le);
    mark[c] = (int8_t)m;
    longest[c] = (uint8_t)le;
    const bool own = c >= own_lo && c < own_hi;
    nref = (own && m > 0); ncoa = (own && m < 0);
  }
  block_count_add<256>(nref, ncoa, &cnt->n_refine, &cnt->n_coarsen);
}
// interface elements through the same operator() (mmesh.hh:745-748; distance 0 => l = 0)
__global__ void __launch_bounds__(256)
k_mark_interface2d(const double2* __restrict__ xy, const int2* __restrict__ segs, int ns, DevParams prm,
                   const Counters* __restrict__ cin, int8_t* __restrict__ mark) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  const double maxDist = prm.distProportion * __longlong_as_double((long long)cin->max_dist_bits);
  const int2 sv = segs[s];
  const double2 a = ldg2(xy + sv.x), b = ldg2(xy + sv.y);
  int refine = 0, coarse = 0;
  const double dist = (0.0 < maxDist) ? 0.0 : maxDist;
  const double l = dist / maxDist;
  const double minH = (1. - l) * prm.minH + l * prm.factor * prm.minH;
  const double maxH = (1. - l) * prm.maxH + l * prm.factor * prm.maxH;
  const double len = edge_len2d(a, b);
  if (len < minH) coarse++;
  if (len > maxH) refine++;
  const double minE = (len < 1e100) ? len : 1e100, maxE = (-1e100 < len) ? len : -1e100, sumE = 0.0 + len;
  if (maxE / minE > prm.edgeRatio) coarse++;
  const double innerRadius = 2 * len / sumE;
  const double jx = b.x - a.x, jy = b.y - a.y;
  const double c1x = a.x + 1.0 * jx, c1y = a.y + 1.0 * jy;
  const double ccx = 0.5 * (a.x + c1x), ccy = 0.5 * (a.y + c1y);
  const double ox = ccx - a.x, oy = ccy - a.y;
  double o2 = 0.0; o2 += ox * ox; o2 += oy * oy;
  const double radiusRatio = sqrt(o2) / (innerRadius * 2);
  if (radiusRatio > prm.radiusRatio) coarse++;
  mark[s] = (int8_t)(coarse > 0 ? -1 : (refine > 0 ? 1 : 0));
}
// RatioIndicator::init reductions (ratioindicator.hh:62-91): min/max are order independent => integer atomics
__global__ void __launch_bounds__(256)
k_edge_minmax(const double2* __restrict__ xy, const int32_t* __restrict__ ends, int stride_ints, int64_t n,
              unsigned long long* __restrict__ mn, unsigned long long* __restrict__ mx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long lo = ~0ull, hi = 0ull;
  if (i < n) {
    const double h = edge_len2d(ldg2(xy + ends[stride_ints * i]), ldg2(xy + ends[stride_ints * i + 1]));
    lo = hi = (unsigned long long)__double_as_longlong(h);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = a < lo ? a : lo; hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0) { if (lo != ~0ull) atomicMin(mn, lo); atomicMax(mx, hi); }
}

// ordered compaction of byte flags equal to `match` (index order = the order MMesh::adapt() visits elements,
// mmesh.hh:832-864): per-block counts by ballot/popc, a one-block scan, then a ballot-ranked scatter.
constexpr int COMPACT_TILE = 2048;   // flags per block (256 threads x 8)
__global__ void __launch_bounds__(256)
k_compact_count(const uint8_t* __restrict__ flags, int64_t n, uint8_t match, uint32_t* __restrict__ block_count) {
  const int64_t base = (int64_t)blockIdx.x * COMPACT_TILE;
  unsigned c = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int64_t i = base + r * 256 + threadIdx.x;
    c += (i < n && flags[i] == match);
  }
  __shared__ unsigned sm[8];
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { unsigned t = 0; for (int w = 0; w < 8; ++w) t += sm[w]; block_count[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024)
k_compact_scan(uint32_t* __restrict__ block_count, int nb, unsigned long long* __restrict__ total) {
  __shared__ unsigned sm[1024];
  unsigned carry = 0;
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned v = i < nb ? block_count[i] : 0u;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const unsigned t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0u;
      __syncthreads();
      sm[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) block_count[i] = carry + sm[threadIdx.x] - v;   // exclusive
    const unsigned tot = sm[1023];
    __syncthreads();
    carry += tot;
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256)
k_compact_scatter(const uint8_t* __restrict__ flags, int64_t n, uint8_t match, const uint32_t* __restrict__ block_off,
                  int32_t* __restrict__ out, int64_t cap) {
  const int64_t base = (int64_t)blockIdx.x * COMPACT_TILE;
  __shared__ unsigned wcount[8];
  unsigned off = block_off[blockIdx.x];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < 8; ++r) {
    const int64_t i = base + r * 256 + threadIdx.x;
    const bool p = (i < n && flags[i] == match);
    const unsigned m = __ballot_sync(0xffffffffu, p);
    if (lane == 0) wcount[w] = __popc(m);
    __syncthreads();
    unsigned before = 0, tot = 0;
    for (int k = 0; k < 8; ++k) { const unsigned t = wcount[k]; if (k < w) before += t; tot += t; }
    if (p) { const int64_t pos = (int64_t)off + before + __popc(m & ((1u << lane) - 1u)); if (pos < cap) out[pos] = (int32_t)i; }
    off += tot;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------
// K5  conservative projection old -> new.  Driver loop mmesh.hh:1143-1161; CutSetTriangulation
//     dune/mmesh/grid/cutsettriangulation.hh:31-62; Sutherland-Hodgman + lineIntersectionPoint
//     dune/mmesh/grid/polygoncutting.hh:17-114.  One thread walks all old children of one new cell in the
//     reference order, so accumulation order (and therefore every bit of the P0 result) is the reference's.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// (line_isect / clip_stage: clip.cuh, shared with the CPU unit test of the clip arithmetic)

// Conservative projection old -> new.  LPC lanes cooperate on one new cell (LPC = 1: one thread walks the pairs
// sequentially, so every bit of the P0 accumulation follows the reference order; LPC = 4: lane q clips pairs
// q, q+4, ... and the per-pair partial sums are added in pair order).  The polygons of the three clip stages
// (<= 4, <= 6, <= 9 vertices) live in shared memory in [slot][thread] layout (bank = thread => conflict free).
constexpr int PROJ_THREADS = 128;
template <int NCOMP, int LPC, int MINB, int UNR = 0>
__global__ void __launch_bounds__(PROJ_THREADS, MINB)
k_project(const double2* __restrict__ xy, const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1,
          const int32_t* __restrict__ cv2, const uint8_t* __restrict__ perm, const long long* __restrict__ new_off,
          const int32_t* __restrict__ new_cell, int64_t i_begin, int64_t nnew, const double* __restrict__ old_xy,
          const int32_t* __restrict__ old_idx, const double* __restrict__ dofs_old, double* __restrict__ dofs_new,
          DevParams prm, double* __restrict__ part_mass, double* __restrict__ part_cov, Counters* __restrict__ cnt) {
  __shared__ double2 sP[6][PROJ_THREADS];      // subject triangle, then output of clip edge 1 (<= 6)
  __shared__ double2 sQ[(UNR == 2 || UNR == 3) ? 6 : 9][PROJ_THREADS];      // outputs of clip edges 0 (<= 4) and 2 (<= 9; UNR >= 2: <= 6, else generic path)
  // UNR < 2: the new cell's TDS-order corners e0, e1, e2, e0 (clip edge st = (sE[st], sE[st+1]));
  // UNR >= 2: per clip edge st the constants dQx, dQy, q2xq1 (rows 3 st ..), hoisted out of the pair loop
  __shared__ double sCell[(UNR == 2 || UNR == 3) ? 9 : 8][PROJ_THREADS];
  double2 (*sE)[PROJ_THREADS] = reinterpret_cast<double2 (*)[PROJ_THREADS]>(&sCell[0][0]);
  double (*sEC)[PROJ_THREADS] = sCell;
  __shared__ double2 sG[3][PROJ_THREADS];      // the new cell's inverse affine map (origin, J^-T rows), used by the fan only
  const int tid = threadIdx.x;
  const int64_t gid = (int64_t)blockIdx.x * PROJ_THREADS + tid;
  const int64_t i = i_begin + gid / LPC;           // new cells [i_begin, nnew) of the pair CSR
  const int q = (int)(gid % LPC);
  const int lane = tid & 31, lead = lane - q;
  const bool live = i < nnew;
  double mass = 0.0, cov = 0.0;
  unsigned npc = 0;
  int K = 0; long long p0 = 0, p1 = 0;
  double ex[3] = { 0, 0, 0 }, ey[3] = { 0, 0, 0 }, kx[3], ky[3], orgx = 0.0, orgy = 0.0, volK = 1.0;
  if (live) {
    K = new_cell[i]; p0 = new_off[i]; p1 = new_off[i + 1];
    const int v[3] = { __ldg(cv0 + K), __ldg(cv1 + K), __ldg(cv2 + K) };
    const unsigned pm = __ldg(perm + K);
#pragma unroll
    for (int t = 0; t < 3; ++t) { const double2 c = ldg2(xy + v[t]); ex[t] = c.x; ey[t] = c.y; }
#pragma unroll
    for (int t = 0; t < 3; ++t) { const int s = (pm >> (2 * t)) & 3; kx[t] = s == 0 ? ex[0] : (s == 1 ? ex[1] : ex[2]); ky[t] = s == 0 ? ey[0] : (s == 1 ? ey[1] : ey[2]); }
    volK = aff2_make(kx[0], ky[0], kx[1], ky[1], kx[2], ky[2]).vol;      // element.geometry().volume()
    if (prm.clip_translate) {
      orgx = ex[0]; orgy = ey[0];
#pragma unroll
      for (int t = 0; t < 3; ++t) { ex[t] -= orgx; ey[t] -= orgy; }
    }
    if (NCOMP == 3) {
      const Aff2 gK = aff2_make(kx[0] - orgx, ky[0] - orgy, kx[1] - orgx, ky[1] - orgy, kx[2] - orgx, ky[2] - orgy);
      sG[0][tid] = make_double2(gK.ox, gK.oy); sG[1][tid] = make_double2(gK.i00, gK.i01); sG[2][tid] = make_double2(gK.i10, gK.i11);
    }
  }
  if (UNR == 2 || UNR == 3) {
#pragma unroll
    for (int st = 0; st < 3; ++st) {
      const int sn = st == 2 ? 0 : st + 1;
      double dQx, dQy, q2xq1;
      clip_edge_constants(ex[st], ey[st], ex[sn], ey[sn], dQx, dQy, q2xq1);
      sEC[3 * st][tid] = dQx; sEC[3 * st + 1][tid] = dQy; sEC[3 * st + 2][tid] = q2xq1;
    }
  } else {
    sE[0][tid] = make_double2(ex[0], ey[0]); sE[1][tid] = make_double2(ex[1], ey[1]);
    sE[2][tid] = make_double2(ex[2], ey[2]); sE[3][tid] = make_double2(ex[0], ey[0]);
  }
  double acc0 = 0.0; bool initialize = true;     // P0 running value (reference order)
  double R1 = 0.0, R2 = 0.0, U = 0.0;            // P1 totals over pairs
  long long maxr = live ? (p1 - p0 + LPC - 1) / LPC : 0;
  if (LPC > 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const long long t = __shfl_xor_sync(0xffffffffu, maxr, o); maxr = t > maxr ? t : maxr; }
  }
  for (long long r = 0; r < maxr; ++r) {
    const long long pr = p0 + LPC * r + q;
    double r1 = 0.0, r2 = 0.0, u = 0.0;          // P1 partial sums of this pair
    if (live && pr < p1) {
      // old corners: cgeo.corner(i) of the cached vertex_ (cutsettriangulation.hh:40)
      const double2* o6 = reinterpret_cast<const double2*>(old_xy + 6 * pr);
      const double2 o0 = __ldg(o6), o1 = __ldg(o6 + 1), o2 = __ldg(o6 + 2);
      const double j00 = o1.x - o0.x, j01 = o1.y - o0.y, j10 = o2.x - o0.x, j11 = o2.y - o0.y;
      double cx[3], cy[3];
      cx[0] = o0.x; cy[0] = o0.y;
      cx[1] = (o0.x + 1.0 * j00) + 0.0 * j10; cy[1] = (o0.y + 1.0 * j01) + 0.0 * j11;
      cx[2] = (o0.x + 0.0 * j00) + 1.0 * j10; cy[2] = (o0.y + 0.0 * j01) + 1.0 * j11;
      if (prm.clip_translate) {
#pragma unroll
        for (int t = 0; t < 3; ++t) { cx[t] -= orgx; cy[t] -= orgy; }
      }
      // cutsettriangulation.hh:46-49: orientation value truncated to int
      const double od = (cy[1] - cy[0]) * (cx[2] - cx[1]) - (cx[1] - cx[0]) * (cy[2] - cy[1]);
      const int o = (fabs(od) < 2147483648.0) ? (int)od : 0;
      const bool sw = o > 0;
      int nC = 3;
      sP[0][tid] = make_double2(cx[0], cy[0]);
      sP[1][tid] = sw ? make_double2(cx[2], cy[2]) : make_double2(cx[1], cy[1]);
      sP[2][tid] = sw ? make_double2(cx[1], cy[1]) : make_double2(cx[2], cy[2]);
      if (UNR == 2 || UNR == 3) {                                       // clip-edge constants hoisted per new cell (sEC)
        auto ldP = [&](int k, double& x, double& y) { const double2 v = sP[k][tid]; x = v.x; y = v.y; };
        auto ldQ = [&](int k, double& x, double& y) { const double2 v = sQ[k][tid]; x = v.x; y = v.y; };
        auto stP = [&](int k, double x, double y) { sP[k][tid] = make_double2(x, y); };
        auto stQ = [&](int k, double x, double y) { sQ[k][tid] = make_double2(x, y); };
        nC = clip_stage_core<3, 4>(ldP, nC, sEC[0][tid], sEC[1][tid], sEC[2][tid], prm.clip_eps, stQ);
        nC = clip_stage_core<4, 6>(ldQ, nC, sEC[3][tid], sEC[4][tid], sEC[5][tid], prm.clip_eps, stP);
        nC = clip_stage_core<6, 6>(ldP, nC, sEC[6][tid], sEC[7][tid], sEC[8][tid], prm.clip_eps, stQ);
      } else if (UNR == 1) {                                // one instance per clip edge, sized for its input (3, <= 4, <= 6)
        auto ldP = [&](int k, double& x, double& y) { const double2 v = sP[k][tid]; x = v.x; y = v.y; };
        auto ldQ = [&](int k, double& x, double& y) { const double2 v = sQ[k][tid]; x = v.x; y = v.y; };
        auto stP = [&](int k, double x, double y) { sP[k][tid] = make_double2(x, y); };
        auto stQ = [&](int k, double x, double y) { sQ[k][tid] = make_double2(x, y); };
        { const double2 q1 = sE[0][tid], q2 = sE[1][tid]; nC = clip_stage<3>(ldP, nC, q1.x, q1.y, q2.x, q2.y, prm.clip_eps, stQ); }
        { const double2 q1 = sE[1][tid], q2 = sE[2][tid]; nC = clip_stage<4>(ldQ, nC, q1.x, q1.y, q2.x, q2.y, prm.clip_eps, stP); }
        { const double2 q1 = sE[2][tid], q2 = sE[3][tid]; nC = clip_stage<6>(ldP, nC, q1.x, q1.y, q2.x, q2.y, prm.clip_eps, stQ); }
      } else
#pragma unroll 1
      for (int st = 0; st < 3; ++st) {                      // one code instance for the three clip edges
        const double2* in = (st == 1) ? &sQ[0][tid] : &sP[0][tid];
        double2* out = (st == 1) ? &sP[0][tid] : &sQ[0][tid];
        const double2 q1 = sE[st][tid], q2 = sE[st + 1][tid];
        const double q1x = q1.x, q1y = q1.y, q2x = q2.x, q2y = q2.y;
        nC = clip_stage<6>([&](int k, double& x, double& y) { const double2 v = in[k * PROJ_THREADS]; x = v.x; y = v.y; },
                           nC, q1x, q1y, q2x, q2y, prm.clip_eps,
                           [&](int k, double x, double y) { out[k * PROJ_THREADS] = make_double2(x, y); });
      }
      auto& sC = sQ;
      if (nC >= 3 || nC < 0) {                              // cutsettriangulation.hh:55 `if (points.size() < 3) return;`
      // old DOFs and the old cell's geometry are needed only now (about every second neighbour pair is empty)
      double u0 = 0.0, u1 = 0.0, u2 = 0.0, gx = 0.0, gy = 0.0;
      Aff2 gO, gK;
      {
        const double* uo = dofs_old + (int64_t)NCOMP * __ldg(old_idx + pr);
        u0 = __ldg(uo);
        if (NCOMP == 3) {
          u1 = __ldg(uo + 1); u2 = __ldg(uo + 2);
          gO = aff2_make(cx[0], cy[0], cx[1], cy[1], cx[2], cy[2]);
          const double2 g0 = sG[0][tid], g1 = sG[1][tid], g2 = sG[2][tid];
          gK.ox = g0.x; gK.oy = g0.y; gK.i00 = g1.x; gK.i01 = g1.y; gK.i10 = g2.x; gK.i11 = g2.y; gK.vol = 0.0;
          const double du1 = u1 - u0, du2 = u2 - u0;
          gx = du1 * gO.i00 + du2 * gO.i01; gy = du1 * gO.i10 + du2 * gO.i11;
        }
      }
      // fan (points[0], points[j-1], points[j]), cutsettriangulation.hh:57-61
      // (the first two points are peeled: the loop body then has no position tests)
      auto eval = [&](const double2 P, double& uj, double& f1, double& f2) {
        uj = 0.0; f1 = 0.0; f2 = 0.0;
        if (NCOMP == 3) {
          const double d0 = P.x - gO.ox, d1 = P.y - gO.oy;
          const double e0 = P.x - gK.ox, e1 = P.y - gK.oy;
          uj = u0 + (d0 * gx + d1 * gy);                       // affine old function: u0 + grad . (x - o)
          f1 = e0 * gK.i00 + e1 * gK.i10;                      // barycentric coordinates in the new cell
          f2 = e0 * gK.i01 + e1 * gK.i11;
        }
      };
      if (NCOMP == 3 && (UNR == 2 || UNR == 3)) {
        unsigned np = 0;
        if (nC >= 3)
          proj_fan<UNR == 3>([&](int k) { return sC[k][tid]; }, nC, u0, gx, gy, gO.ox, gO.oy, make_double2(gK.ox, gK.oy),
                             make_double2(gK.i00, gK.i01), make_double2(gK.i10, gK.i11), prm.drop_area, r1, r2, u, cov, np);
        else {                                              // the polygon outgrew its slots: degenerate input only
          ProjArgs pa; pa.xy = xy; pa.cv0 = cv0; pa.cv1 = cv1; pa.cv2 = cv2; pa.perm = perm; pa.old_xy = old_xy; pa.old_idx = old_idx;
          pa.dofs_old = dofs_old; pa.prm = prm;
          const ProjPartial g = proj_pair_generic<UNR == 3>(pa, K, pr);
          r1 = g.r1; r2 = g.r2; u = g.u; cov += g.cov; np = g.npc;
        }
        npc += np;
      } else {
      const double2 F = sC[0][tid];
      double2 Pp = sC[1][tid];
      double fu, ff1, ff2, pu, pf1, pf2;
      eval(F, fu, ff1, ff2);
      eval(Pp, pu, pf1, pf2);
      const double fx = F.x, fy = F.y;
      for (int j = 2; j < nC; ++j) {
        const double2 P = sC[j][tid];
        double uj, f1, f2;
        eval(P, uj, f1, f2);
        {
          const double a00 = Pp.x - fx, a01 = Pp.y - fy, a10 = P.x - fx, a11 = P.y - fy;
          const double aT = fabs(a00 * a11 - a10 * a01) * 0.5;
          if (aT > prm.drop_area) {                                     // cutsettriangulation.hh:60
            cov += aT; npc++;
            if (NCOMP == 1) {
              const double w = aT / volK;                               // FV restrictLocal weight |son|/|father|
              if (initialize) acc0 = w * u0; else acc0 += w * u0;
              initialize = false;
            } else {
              const double su = fu + pu + uj;
              const double s1 = ff1 + pf1 + f1, s2 = ff2 + pf2 + f2;
              const double w12 = aT * (1.0 / 12.0);
              const double d1 = fu * ff1 + pu * pf1 + uj * f1;
              const double d2 = fu * ff2 + pu * pf2 + uj * f2;
              r1 += w12 * (su * s1 + d1);
              r2 += w12 * (su * s2 + d2);
              u += (aT * (1.0 / 3.0)) * su;
            }
          }
        }
        Pp = P; pu = uj; pf1 = f1; pf2 = f2;
      }
      }
      }
    }
    if (NCOMP == 3) {
      if (LPC == 1) { R1 += r1; R2 += r2; U += u; }
      else {
#pragma unroll
        for (int l = 0; l < LPC; ++l) { R1 += shfl_d(r1, lead + l); R2 += shfl_d(r2, lead + l); U += shfl_d(u, lead + l); }
      }
    }
  }
  if (live && q == 0) {
    if (NCOMP == 1) {
      dofs_new[K] = acc0;
      mass = volK * acc0;
    } else {
      const double R0 = U - R1 - R2;
      const double f = 3.0 / volK;
      const double r0 = f * (4.0 * R0 - U), r1 = f * (4.0 * R1 - U), r2 = f * (4.0 * R2 - U);
      dofs_new[3 * (int64_t)K] = r0; dofs_new[3 * (int64_t)K + 1] = r1; dofs_new[3 * (int64_t)K + 2] = r2;
      mass = volK * ((r0 + r1 + r2) / 3.0);
    }
  }
  // deterministic two-level sums: fixed-shape shuffle tree per block, per-block partials reduced by k_reduce_partials
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mass += __shfl_xor_sync(0xffffffffu, mass, o); cov += __shfl_xor_sync(0xffffffffu, cov, o); }
  npc = warp_sum(npc);
  __shared__ double sm[4], sc[4]; __shared__ unsigned sn[4];
  if (lane == 0) { sm[tid >> 5] = mass; sc[tid >> 5] = cov; sn[tid >> 5] = npc; }
  __syncthreads();
  if (tid == 0) {
    part_mass[blockIdx.x] = (sm[0] + sm[1]) + (sm[2] + sm[3]);
    part_cov[blockIdx.x] = (sc[0] + sc[1]) + (sc[2] + sc[3]);
    atomicAdd(&cnt->n_pieces, (unsigned long long)(sn[0] + sn[1] + sn[2] + sn[3]));
  }
}
// DG-P1 projection, tile form (project.cuh): NT pair slots and NT/4 new cells per CTA; phases A-E separated by barriers.
template <int NT, int MINB, bool FMA>
__global__ void __launch_bounds__(NT, MINB)
k_project_tile(const ProjArgs a, double* __restrict__ part_mass, double* __restrict__ part_cov, Counters* __restrict__ cnt) {
  extern __shared__ __align__(16) unsigned char proj_smem[];
  ProjTile<NT>& T = *reinterpret_cast<ProjTile<NT>*>(proj_smem);
  constexpr int NCELL = NT / 4;
  const int t = threadIdx.x, lane = t & 31;
  const int64_t i0 = a.i_begin + (int64_t)blockIdx.x * NCELL;
  const int ncell = (int)((a.nnew - i0) < NCELL ? (a.nnew - i0) : NCELL);
  const long long p_lo = a.new_off[i0];
  const int np_tile = (int)(a.new_off[i0 + ncell] - p_lo);
  bool four = true;
  if (t < ncell) four = proj_cell_setup<NT>(T, a, t, i0 + t, p_lo) == 4;
  if (t == 0) T.ndense = 0;
  const bool uniform = __syncthreads_and(four) != 0;               // every cell of the tile has four pairs: transposed columns
  for (int cl = 0; cl < np_tile; cl += NT) {
    const int np = (np_tile - cl) < NT ? (np_tile - cl) : NT;
    if (!uniform) { if (t < ncell) proj_paint<NT>(T, t, cl); __syncthreads(); }
    if (t < np) proj_load_pair<NT>(T, a, t, p_lo + cl + t, uniform);    // coalesced: thread t reads pair cl + t
    __syncthreads();
    int j = t; bool live = t < np;
    if (uniform) { const int q = t / NCELL, c = t - q * NCELL; j = 4 * c + q; live = c < ncell; }
    bool keep = false;
    if (live) keep = proj_clip_column<NT>(T, a, proj_column<NT>(uniform, j), uniform ? (j >> 2) : T.owner[j]);
    const unsigned m = __ballot_sync(0xffffffffu, keep);           // C: warp-aggregated append to the dense list
    if (m) {
      const int leader = __ffs(m) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&T.ndense, __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (keep) T.dense[base + __popc(m & ((1u << lane) - 1u))] = (uint8_t)j;
    }
    __syncthreads();
    const int nd = T.ndense;
    if (t < nd) { const int jd = T.dense[t]; proj_integrate<NT, FMA>(T, a, jd, p_lo + cl + jd, uniform); }
    __syncthreads();
    if (t == 0) T.ndense = 0;
    if (t < ncell) proj_gather<NT>(T, t, cl, uniform);
    if (cl + NT < np_tile) __syncthreads();                        // the next chunk overwrites the columns
  }
  double mass = 0.0, cov = 0.0; unsigned npc = 0;
  if (t < ncell) { mass = proj_finish<NT>(T, a, t); cov = T.acc[3][t]; npc = (unsigned)T.npc[t]; }
  // deterministic two-level sums: fixed-shape shuffle tree per warp, fixed order over the warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mass += __shfl_xor_sync(0xffffffffu, mass, o); cov += __shfl_xor_sync(0xffffffffu, cov, o); }
  npc = warp_sum(npc);
  __shared__ double sm[NT / 32], sc[NT / 32]; __shared__ unsigned sn[NT / 32];
  if (lane == 0) { sm[t >> 5] = mass; sc[t >> 5] = cov; sn[t >> 5] = npc; }
  __syncthreads();
  if (t == 0) {
    double M = sm[0], C = sc[0]; unsigned N = sn[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) { M += sm[w]; C += sc[w]; N += sn[w]; }
    part_mass[blockIdx.x] = M; part_cov[blockIdx.x] = C;
    if (N) atomicAdd(&cnt->n_pieces, (unsigned long long)N);
  }
}
// DG-P1 projection as two passes without phase barriers (development variant):
//   k_project_pairs  one thread per (old, new) pair: clip in the thread's own shared-memory column, fan + quadrature, the
//                    pair's partial sums (r1, r2, u, covered area) and piece count go to global memory;
//   k_project_cells  one thread per new cell: the partial sums of its pairs added in CSR order, the local solve.
// The only barrier is the one after the per-cell set-up (clip-edge constants, inverse affine map) that the NT/4 cell
// threads of a CTA share with its NT pair threads.
template <int NT, int MINB, bool FMA>
__global__ void __launch_bounds__(NT, MINB)
k_project_pairs(const ProjArgs a, double4* __restrict__ partial, uint8_t* __restrict__ pieces) {
  extern __shared__ __align__(16) unsigned char proj_smem[];
  ProjTile<NT>& T = *reinterpret_cast<ProjTile<NT>*>(proj_smem);
  constexpr int NCELL = NT / 4;
  const int t = threadIdx.x;
  const int64_t i0 = a.i_begin + (int64_t)blockIdx.x * NCELL;
  const int ncell = (int)((a.nnew - i0) < NCELL ? (a.nnew - i0) : NCELL);
  const long long p_lo = a.new_off[i0];
  const int np_tile = (int)(a.new_off[i0 + ncell] - p_lo);
  bool four = true;
  if (t < ncell) four = proj_cell_setup<NT>(T, a, t, i0 + t, p_lo) == 4;
  const bool uniform = __syncthreads_and(four) != 0;
  for (int cl = 0; cl < np_tile; cl += NT) {
    const int np = (np_tile - cl) < NT ? (np_tile - cl) : NT;
    if (!uniform) { __syncthreads(); if (t < ncell) proj_paint<NT>(T, t, cl); __syncthreads(); }
    int j = t; bool live = t < np;
    if (uniform) { const int q = t / NCELL, c = t - q * NCELL; j = 4 * c + q; live = c < ncell; }
    if (!live) continue;
    const long long p = p_lo + cl + j;
    const int c = uniform ? (j >> 2) : T.owner[j];
    const int col = t;                                               // the thread's own column: no other thread touches it
    double cx[3], cy[3];
    proj_old_corners(a.old_xy, p, T.org[c], cx, cy);
    {
      const double od = (cy[1] - cy[0]) * (cx[2] - cx[1]) - (cx[1] - cx[0]) * (cy[2] - cy[1]);
      const int o = (fabs(od) < 2147483648.0) ? (int)od : 0;
      const bool sw = o > 0;
      T.P[0][col] = make_double2(cx[0], cy[0]);
      T.P[1][col] = sw ? make_double2(cx[2], cy[2]) : make_double2(cx[1], cy[1]);
      T.P[2][col] = sw ? make_double2(cx[1], cy[1]) : make_double2(cx[2], cy[2]);
    }
    double r1 = 0.0, r2 = 0.0, u = 0.0, cov = 0.0; unsigned npc = 0;
    if (proj_clip_column<NT>(T, a, col, c)) {
      const int nC = T.npts[col];
      if (nC != PROJ_HARD) {
        const Aff2 gO = aff2_make(cx[0], cy[0], cx[1], cy[1], cx[2], cy[2]);
        const double* uo = a.dofs_old + 3 * (int64_t)a.old_idx[p];
        const double u0 = uo[0], u1 = uo[1], u2 = uo[2];
        const double du1 = u1 - u0, du2 = u2 - u0;
        const double gx = du1 * gO.i00 + du2 * gO.i01, gy = du1 * gO.i10 + du2 * gO.i11;
        proj_fan<FMA>([&](int k) { return T.Q[k][col]; }, nC, u0, gx, gy, gO.ox, gO.oy, T.G[0][c], T.G[1][c], T.G[2][c], a.prm.drop_area,
                      r1, r2, u, cov, npc);
      } else {
        const ProjPartial g = proj_pair_generic<FMA>(a, T.K[c], p);
        r1 = g.r1; r2 = g.r2; u = g.u; cov = g.cov; npc = g.npc;
      }
    }
    partial[p] = make_double4(r1, r2, u, cov);
    pieces[p] = (uint8_t)npc;
  }
}
__global__ void __launch_bounds__(128)
k_project_cells(const ProjArgs a, const double4* __restrict__ partial, const uint8_t* __restrict__ pieces, double* __restrict__ part_mass,
                double* __restrict__ part_cov, Counters* __restrict__ cnt) {
  const int64_t i = a.i_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double mass = 0.0, cov = 0.0; unsigned npc = 0;
  if (i < a.nnew) {
    const int K = a.new_cell[i];
    const int v[3] = { a.cv0[K], a.cv1[K], a.cv2[K] };
    const unsigned pm = a.perm[K];
    double ex[3], ey[3], kx[3], ky[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) { const double2 q = ldg2(a.xy + v[t]); ex[t] = q.x; ey[t] = q.y; }
#pragma unroll
    for (int t = 0; t < 3; ++t) { const int s = (pm >> (2 * t)) & 3; kx[t] = s == 0 ? ex[0] : (s == 1 ? ex[1] : ex[2]); ky[t] = s == 0 ? ey[0] : (s == 1 ? ey[1] : ey[2]); }
    const double volK = aff2_make(kx[0], ky[0], kx[1], ky[1], kx[2], ky[2]).vol;
    double R1 = 0.0, R2 = 0.0, U = 0.0;
    for (long long p = a.new_off[i]; p < a.new_off[i + 1]; ++p) {     // pairs are added in CSR order
      const double4 r = partial[p];
      R1 += r.x; R2 += r.y; U += r.z; cov += r.w; npc += pieces[p];
    }
    const double R0 = U - R1 - R2;
    const double f = 3.0 / volK;
    const double r0 = f * (4.0 * R0 - U), r1 = f * (4.0 * R1 - U), r2 = f * (4.0 * R2 - U);
    double* o = a.dofs_new + 3 * (int64_t)K;
    o[0] = r0; o[1] = r1; o[2] = r2;
    mass = volK * ((r0 + r1 + r2) / 3.0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mass += __shfl_xor_sync(0xffffffffu, mass, o); cov += __shfl_xor_sync(0xffffffffu, cov, o); }
  npc = warp_sum(npc);
  __shared__ double sm[4], sc[4]; __shared__ unsigned sn[4];
  const int lane = threadIdx.x & 31;
  if (lane == 0) { sm[threadIdx.x >> 5] = mass; sc[threadIdx.x >> 5] = cov; sn[threadIdx.x >> 5] = npc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    part_mass[blockIdx.x] = (sm[0] + sm[1]) + (sm[2] + sm[3]);
    part_cov[blockIdx.x] = (sc[0] + sc[1]) + (sc[2] + sc[3]);
    const unsigned N = sn[0] + sn[1] + sn[2] + sn[3];
    if (N) atomicAdd(&cnt->n_pieces, (unsigned long long)N);
  }
}

// Fixed-shape reduction of the per-block partial sums (deterministic: the shape depends only on n and the grid).  Every
// block reduces one contiguous slice; the block that finishes last (ticket) adds the slice sums in slice order.
constexpr int REDUCE_MAX_BLOCKS = 256;
__global__ void __launch_bounds__(256)
k_reduce_partials(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ slice_a,
                  double* __restrict__ slice_b, unsigned* __restrict__ ticket, Counters* __restrict__ cnt) {
  __shared__ double sa[256], sb[256];
  __shared__ bool last;
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(n, lo + per);
  double xa = 0.0, xb = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += 256) { xa += a[i]; xb += b[i]; }
  sa[threadIdx.x] = xa; sb[threadIdx.x] = xb;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    slice_a[blockIdx.x] = sa[0]; slice_b[blockIdx.x] = sb[0];
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    double ma = 0.0, mb = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) { ma += __ldcg(slice_a + k); mb += __ldcg(slice_b + k); }
    cnt->mass_new = ma; cnt->covered_area = mb;
    *ticket = 0;
  }
}
// MMeshCachingEntity::intersectionVolume (dune/mmesh/grid/cachingentity.hh:154-167): |shoelace(SH(vertex_, new))|
__global__ void __launch_bounds__(128)
k_isect_volume(const double2* __restrict__ xy, const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1,
               const int32_t* __restrict__ cv2, const long long* __restrict__ new_off, const int32_t* __restrict__ new_cell,
               int64_t nnew, const double* __restrict__ old_xy, double* __restrict__ vol) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnew) return;
  const int K = new_cell[i];
  const int v[3] = { __ldg(cv0 + K), __ldg(cv1 + K), __ldg(cv2 + K) };
  double ex[3], ey[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) { const double2 q = ldg2(xy + v[t]); ex[t] = q.x; ey[t] = q.y; }
  for (long long pr = new_off[i]; pr < new_off[i + 1]; ++pr) {
    const double* o = old_xy + 6 * pr;
    const double cx[3] = { o[0], o[2], o[4] }, cy[3] = { o[1], o[3], o[5] };
    Poly A, B; Poly* poly;
    sh_clip(cx, cy, ex, ey, 1e-14, A, B, poly);
    double area = 0.0;
    if (poly->n >= 3) {
      for (int a = 0; a < poly->n; ++a) {
        const int b = a + 1 == poly->n ? 0 : a + 1;
        area += poly->x[a] * poly->y[b];
        area -= poly->x[b] * poly->y[a];
      }
    }
    vol[pr] = fabs(0.5 * area);
  }
}

// ------------------------------------------------------------------------------------------------------
// 3-D: validity (mmesh.hh:1174-1178 -> CGAL/Cartesian/function_objects.h:1162-1169), marks (6 edges; coarsening
// disabled, ratioindicator.hh:151-156), distance estimate (distance.hh:157-173), P0 trace / skeleton transfer.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load3(const double* __restrict__ x, const double* __restrict__ s, int v, bool shift,
                                      double* o) {
  o[0] = __ldg(x + 3 * (int64_t)v); o[1] = __ldg(x + 3 * (int64_t)v + 1); o[2] = __ldg(x + 3 * (int64_t)v + 2);
  if (shift) {
    o[0] = o[0] + __ldg(s + 3 * (int64_t)v); o[1] = o[1] + __ldg(s + 3 * (int64_t)v + 1);
    o[2] = o[2] + __ldg(s + 3 * (int64_t)v + 2);
  }
}
template <bool SHIFT>
__global__ void __launch_bounds__(256)
k_validate3d(const double* __restrict__ x, const double* __restrict__ sh, const int32_t* __restrict__ cv0,
             const int32_t* __restrict__ cv1, const int32_t* __restrict__ cv2, const int32_t* __restrict__ cv3, int64_t nc,
             uint8_t* __restrict__ valid, int8_t* __restrict__ orient, int32_t* __restrict__ wl, Counters* __restrict__ cnt) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned ninv = 0, nnonpos = 0;
  bool need = false;
  if (c < nc) {
    double p[4][3];
    load3(x, sh, __ldg(cv0 + c), SHIFT, p[0]); load3(x, sh, __ldg(cv1 + c), SHIFT, p[1]);
    load3(x, sh, __ldg(cv2 + c), SHIFT, p[2]); load3(x, sh, __ldg(cv3 + c), SHIFT, p[3]);
    double det;
    int s = orient3d_filter(p[0], p[1], p[2], p[3], det);     // same differences and det3 as Compute_volume_3
    const double vol = det / 6;
    const bool ok = !(vol <= 0.0);
    need = (s == SIGN_UNKNOWN);
    valid[c] = ok; orient[c] = (int8_t)(need ? 0 : s);
    ninv = !ok; nnonpos = (!need && s <= 0);
  }
  wl_push(need, (int)c, wl, &cnt->wl_orient);
  block_count_add<256>(ninv, nnonpos, &cnt->n_invalid, &cnt->n_orient_nonpos);
}
template <bool SHIFT>
__global__ void __launch_bounds__(128)
k_exact_orient3d(const double* __restrict__ x, const double* __restrict__ sh, const int32_t* __restrict__ cv0,
                 const int32_t* __restrict__ cv1, const int32_t* __restrict__ cv2, const int32_t* __restrict__ cv3,
                 const int32_t* __restrict__ wl, int8_t* __restrict__ orient, Counters* __restrict__ cnt, int publish) {
  if (publish && blockIdx.x == 0 && threadIdx.x == 0) cnt->n_invalid_global = cnt->n_invalid;
  const unsigned long long n = cnt->wl_orient;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned nnonpos = 0, nexact = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = wl[i];
    double p[4][3];
    load3(x, sh, cv0[c], SHIFT, p[0]); load3(x, sh, cv1[c], SHIFT, p[1]);
    load3(x, sh, cv2[c], SHIFT, p[2]); load3(x, sh, cv3[c], SHIFT, p[3]);
    int s = orient3d_interval(p[0], p[1], p[2], p[3]);
    if (s == SIGN_UNKNOWN) { s = orient3d_exact(p[0], p[1], p[2], p[3]); nexact++; }
    if (s == SIGN_DOMAIN) { s = 0; atomicAdd(&cnt->pad_[1], 1ull); }
    orient[c] = (int8_t)s;
    nnonpos += (s <= 0);
  }
  block_count_add<128>(nnonpos, nexact, &cnt->n_orient_nonpos, &cnt->n_exact_orient);
}
__global__ void __launch_bounds__(256)
k_mark3d(const double* __restrict__ x, const double* __restrict__ dist, const int32_t* __restrict__ cv0,
         const int32_t* __restrict__ cv1, const int32_t* __restrict__ cv2, const int32_t* __restrict__ cv3,
         const uint8_t* __restrict__ perm, int64_t nc, DevParams prm, const Counters* __restrict__ cin,
         int8_t* __restrict__ mark, uint8_t* __restrict__ longest, Counters* __restrict__ cnt) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned nref = 0;
  if (c < nc) {
    const double maxDist = prm.distProportion * __longlong_as_double((long long)cin->max_dist_bits);
    const int v[4] = { __ldg(cv0 + c), __ldg(cv1 + c), __ldg(cv2 + c), __ldg(cv3 + c) };
    const unsigned p = __ldg(perm + c);
    double cr[4][3]; double dsum = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int vv = v[(p >> (2 * k)) & 3]; load3(x, nullptr, vv, false, cr[k]); dsum += __ldg(dist + vv); }
    dsum /= 4;
    const double d = (dsum < maxDist) ? dsum : maxDist;
    const double l = d / maxDist;
    const double maxH = (1. - l) * prm.maxH + l * prm.factor * prm.maxH;
    const int EV[6][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 }, { 0, 3 }, { 1, 3 }, { 2, 3 } };
    int refine = 0, li = -1; double ll = -1e100;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double j = cr[EV[i][1]][k] - cr[EV[i][0]][k]; s += j * j; }
      const double len = sqrt(s);
      if (len > maxH) refine++;
      if (len > ll) { ll = len; li = i; }
    }
    const int m = refine > 0 ? 1 : 0;
    mark[c] = (int8_t)m; longest[c] = (uint8_t)li;
    nref = (m > 0);
  }
  block_count_add<256>(nref, 0u, &cnt->n_refine, &cnt->n_coarsen);
}
// RatioIndicator::init in 3-D (ratioindicator.hh:62-91): min / max over the edges of the listed simplices (k = 3: interface
// triangles, k = 4: tetrahedra); every edge is seen by all its simplices, which min / max do not notice
__global__ void __launch_bounds__(256)
k_edge_minmax3d(const double* __restrict__ x, const int32_t* __restrict__ simplices, int k, int64_t n,
                unsigned long long* __restrict__ mn, unsigned long long* __restrict__ mx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long lo = ~0ull, hi = 0ull;
  if (i < n) {
    for (int a = 0; a < k; ++a)
      for (int b = a + 1; b < k; ++b) {
        const int va = simplices[k * i + a], vb = simplices[k * i + b];
        double q = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) { const double j = __ldg(x + 3 * (int64_t)vb + c) - __ldg(x + 3 * (int64_t)va + c); q += j * j; }
        const unsigned long long h = (unsigned long long)__double_as_longlong(sqrt(q));
        lo = h < lo ? h : lo; hi = h > hi ? h : hi;
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = a < lo ? a : lo; hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0) { if (lo != ~0ull) atomicMin(mn, lo); atomicMax(mx, hi); }
}
// interface triangles of a 3-D grid through RatioIndicator::operator() (mmesh.hh:745-748): distance 0 => l = 0; coarsening
// cannot be returned because dim == 3 (ratioindicator.hh:151-153) => +1 iff an edge is longer than maxH
__global__ void __launch_bounds__(256)
k_mark_interface3d(const double* __restrict__ x, const int32_t* __restrict__ tris, int ns, DevParams prm,
                   const Counters* __restrict__ cin, int8_t* __restrict__ mark) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  const double maxDist = prm.distProportion * __longlong_as_double((long long)cin->max_dist_bits);
  const double dist = (0.0 < maxDist) ? 0.0 : maxDist;
  const double l = dist / maxDist;
  const double maxH = (1. - l) * prm.maxH + l * prm.factor * prm.maxH;
  int refine = 0;
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const int va = tris[3 * s + e], vb = tris[3 * s + (e + 1) % 3];
    double q = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) { const double j = __ldg(x + 3 * (int64_t)vb + c) - __ldg(x + 3 * (int64_t)va + c); q += j * j; }
    if (sqrt(q) > maxH) refine++;
  }
  mark[s] = (int8_t)(refine > 0 ? 1 : 0);
}
// 3-D facet leaves: 7 sample points (3 corners, 3 edge midpoints, centre), distance.hh:157-173.  Same tree machinery as
// in 2-D: fp32 boxes rounded outward (two float4 per node), one-launch refit, hint-seeded leaf->root walk.
struct TriLeaf { double pt[7][3]; };
struct Box3 { float4 lo, hi; };
__device__ __forceinline__ Box3 box3_merge(const Box3& a, const Box3& b) {
  Box3 m;
  m.lo = make_float4(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z), 0.f);
  m.hi = make_float4(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z), 0.f);
  return m;
}
__device__ __forceinline__ Box3 ld_box3(const Box3* p) {
  Box3 b; b.lo = __ldg(&p->lo); b.hi = __ldg(&p->hi); return b;
}
__device__ __forceinline__ Box3 ld_box3_cg(const Box3* p) {
  Box3 b; b.lo = __ldcg(&p->lo); b.hi = __ldcg(&p->hi); return b;
}
__global__ void __launch_bounds__(256)
k_bvh_build3d(const double* __restrict__ x, const int32_t* __restrict__ tris_sorted, int ns, int L,
              TriLeaf* __restrict__ leaf, Box3* __restrict__ box, unsigned* __restrict__ ticket) {
  __shared__ Box3 sb[512];
  __shared__ bool last;
  const int t = threadIdx.x;
  const int W = L < 256 ? L : 256;
  const int i = blockIdx.x * W + t;
  Box3 b; b.lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f); b.hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
  if (t < W && i < ns) {
    const double* a = x + 3 * (int64_t)tris_sorted[3 * i];
    const double* bb = x + 3 * (int64_t)tris_sorted[3 * i + 1];
    const double* cc = x + 3 * (int64_t)tris_sorted[3 * i + 2];
    double c[3][3], ctr[3];
    const double third = 1.0 / 3.0;
    for (int k = 0; k < 3; ++k) {
      const double j0 = bb[k] - a[k], j1 = cc[k] - a[k];
      c[0][k] = a[k];
      c[1][k] = (a[k] + 1.0 * j0) + 0.0 * j1;            // AffineGeometry::corner
      c[2][k] = (a[k] + 0.0 * j0) + 1.0 * j1;
      ctr[k] = (a[k] + third * j0) + third * j1;         // center() = global(1/3, 1/3)
    }
    TriLeaf f;
    for (int tt = 0; tt < 3; ++tt) {
      const int t2 = (tt + 1) % 3;
      for (int k = 0; k < 3; ++k) { f.pt[2 * tt][k] = c[tt][k]; f.pt[2 * tt + 1][k] = 0.5 * (c[tt][k] + c[t2][k]); }
    }
    for (int k = 0; k < 3; ++k) f.pt[6][k] = ctr[k];
    leaf[i] = f;
    double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int tt = 0; tt < 7; ++tt) for (int k = 0; k < 3; ++k) { lo[k] = fmin(lo[k], f.pt[tt][k]); hi[k] = fmax(hi[k], f.pt[tt][k]); }
    b.lo = make_float4(__double2float_rd(lo[0]), __double2float_rd(lo[1]), __double2float_rd(lo[2]), 0.f);
    b.hi = make_float4(__double2float_ru(hi[0]), __double2float_ru(hi[1]), __double2float_ru(hi[2]), 0.f);
  }
  if (t < W) { sb[W + t] = b; box[L + i] = b; }
  const int g0 = L / W + blockIdx.x;
  for (int w = W >> 1; w >= 1; w >>= 1) {
    __syncthreads();
    if (t < w) {
      const Box3 m = box3_merge(sb[2 * (w + t)], sb[2 * (w + t) + 1]);
      sb[w + t] = m;
      int lev = 0; for (int xx = w; xx > 1; xx >>= 1) ++lev;
      box[((size_t)g0 << lev) + t] = m;
    }
  }
  if (L <= W) return;
  __threadfence();
  __syncthreads();
  if (t == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  const int nb = L / W;
  for (int w = nb >> 1; w >= 1; w >>= 1) {
    for (int j = w + t; j < 2 * w; j += 256) box[j] = box3_merge(ld_box3_cg(box + 2 * j), ld_box3_cg(box + 2 * j + 1));
    __threadfence();
    __syncthreads();
  }
  if (t == 0) *ticket = 0;
}
struct DistQuery3 {
  double p[3], best2;
  float plo[3], phi[3], T2f;
  int arg;
  // squared quantities are compared exactly like distance.hh:160-171; a node is skipped only if its fp32 lower bound
  // exceeds best2 (1 + 1e-12): computed squared norms carry a relative error of a few ulps
  __device__ __forceinline__ void set_threshold() { T2f = __double2float_ru(best2 * (1.0 + 1e-12)); }
  __device__ __forceinline__ float lb2(const Box3& b) const {
    const float dx = fmaxf(fmaxf(__fsub_rd(b.lo.x, phi[0]), __fsub_rd(plo[0], b.hi.x)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rd(b.lo.y, phi[1]), __fsub_rd(plo[1], b.hi.y)), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rd(b.lo.z, phi[2]), __fsub_rd(plo[2], b.hi.z)), 0.f);
    return __fadd_rd(__fadd_rd(__fmul_rd(dx, dx), __fmul_rd(dy, dy)), __fmul_rd(dz, dz));
  }
  __device__ __forceinline__ void eval_leaf(const TriLeaf* __restrict__ leaf, int s) {
    const TriLeaf& f = leaf[s];
    bool better = false;
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      double n2 = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double d = p[k] - __ldg(&f.pt[t][k]); n2 += d * d; }
      if (n2 < best2) { best2 = n2; better = true; }
    }
    if (better) { arg = s; set_threshold(); }
  }
  __device__ __forceinline__ void dfs(const TriLeaf* __restrict__ leaf, const Box3* __restrict__ box, int ns, int L, int root) {
    int stack[32]; float sd2[32]; int sp = 0;
    int node = root;
    for (;;) {
      if (node >= L) {
        if (node - L < ns) eval_leaf(leaf, node - L);
        node = 0;
      } else {
        const float dl = lb2(ld_box3(box + 2 * node)), dr = lb2(ld_box3(box + 2 * node + 1));
        const bool lf = dl <= dr;
        const float dn = lf ? dl : dr, df = lf ? dr : dl;
        const int nn = 2 * node + (lf ? 0 : 1), nf = 2 * node + (lf ? 1 : 0);
        if (dn > T2f) node = 0;
        else { if (!(df > T2f)) { stack[sp] = nf; sd2[sp] = df; ++sp; } node = nn; }
      }
      if (node == 0) {
        while (sp > 0) { --sp; if (!(sd2[sp] > T2f)) { node = stack[sp]; break; } }
        if (node == 0) break;
      }
    }
  }
};
__global__ void __launch_bounds__(256)
k_distance3d(const double* __restrict__ x, int64_t nv, const TriLeaf* __restrict__ leaf, const Box3* __restrict__ box,
             int ns, int L, double* __restrict__ dist, int32_t* __restrict__ hint, Counters* __restrict__ cnt) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = v < nv && ns > 0;
  DistQuery3 q;
  q.best2 = 1e100; q.arg = -1;                             // distance.hh:160 initial value
  if (live) {
    load3(x, nullptr, (int)v, false, q.p);
#pragma unroll
    for (int k = 0; k < 3; ++k) { q.plo[k] = __double2float_rd(q.p[k]); q.phi[k] = __double2float_ru(q.p[k]); }
    q.set_threshold();
  }
  int h = live ? hint[v] : 0;
  if (h >= ns) h = -1;
  const unsigned cold = __ballot_sync(0xffffffffu, live && h < 0);
  if (cold) {
    const int leader = __ffs(cold) - 1;
    if ((threadIdx.x & 31) == leader) q.dfs(leaf, box, ns, L, 1);
    const int seed = __shfl_sync(0xffffffffu, q.arg, leader);
    if (live && h < 0) h = seed;
  }
  double best = 1e100;
  if (live) {
    q.eval_leaf(leaf, h);
    for (int node = L + h; node > 1; node >>= 1) {
      const int sib = node ^ 1;
      if (!(q.lb2(ld_box3(box + sib)) > q.T2f)) q.dfs(leaf, box, ns, L, sib);
    }
    hint[v] = q.arg;
    // sqrt is monotone: sqrt(min over facets of the squared minimum) == min over facets of sqrt(...) bit for bit
    best = sqrt(q.best2);
  }
  if (v < nv) dist[v] = best;
  unsigned long long bits = (v < nv) ? (unsigned long long)__double_as_longlong(best) : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, bits, o); bits = t > bits ? t : bits; }
  if ((threadIdx.x & 31) == 0) atomicMax(&cnt->max_dist_bits, bits);
}
// dune/python/mmesh/pyskeletontrace.hh:42-64,133-160 (P0 data)
__global__ void __launch_bounds__(256)
k_trace_skeleton(int64_t nf, const int32_t* __restrict__ inside, const int32_t* __restrict__ outside,
                 const double* __restrict__ u, const double* __restrict__ ug, double* __restrict__ tin,
                 double* __restrict__ tout, double* __restrict__ skel) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  tin[f] = __ldg(u + inside[f]);
  tout[f] = outside[f] >= 0 ? __ldg(u + outside[f]) : 0.0;
  skel[f] = __ldg(ug + f);
}

// MMeshInterfaceGrid::adapt(handle) with FV (P0) data, dune/mmesh/interface/grid.hh:388-415: the longer of (old, new)
// is the father; old longer => prolongLocal copies the father value, else restrictLocal adds (|old|/|new|) * old.
__global__ void __launch_bounds__(256)
k_project_interface_p0(const double* __restrict__ new_len, const long long* __restrict__ new_off, int64_t nnew,
                       const double* __restrict__ old_len, const int32_t* __restrict__ old_idx,
                       const double* __restrict__ dofs_old, double* __restrict__ dofs_new) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnew) return;
  bool initialize = true; double acc = 0.0;
  const double ln = new_len[i];
  for (long long p = new_off[i]; p < new_off[i + 1]; ++p) {
    const double u = __ldg(dofs_old + old_idx[p]);
    if (old_len[p] > ln) acc = u;
    else { const double w = old_len[p] / ln; if (initialize) acc = w * u; else acc += w * u; }
    initialize = false;
  }
  dofs_new[i] = acc;
}
// MMeshInterfaceGrid::adapt(handle) with DG-P1 data on the segments (2 nodal DOFs on the id-ordered corners),
// dune/mmesh/interface/grid.hh:388-415 + geometryInFather (interface/entity.hh:491-507).  The longer of (old, new) is the
// father: old longer => the father's function evaluated at the son's corners (prolongLocal, overwrites); else the local L2
// projection of the son on the father's P1 space, M_f^-1 int_son u phi_i, set or added (restrictLocal).  The arithmetic is
// the oracle's definition (dune-fem un-vendored): local(x) = ((x - c0) . d) / (d . d), vertex rule len/6 (2 1; 1 2).
__device__ __forceinline__ double seg_len4(const double4 c) {
  const double dx = c.z - c.x, dy = c.w - c.y; double q = 0.0; q += dx * dx; q += dy * dy; return sqrt(q);
}
__device__ __forceinline__ double seg_local4(const double4 c, double px, double py) {
  const double dx = c.z - c.x, dy = c.w - c.y;
  double s = 0.0; s += (px - c.x) * dx; s += (py - c.y) * dy;
  double n2 = 0.0; n2 += dx * dx; n2 += dy * dy;
  return s / n2;
}
__global__ void __launch_bounds__(256)
k_project_interface_p1(const double4* __restrict__ new_xy, const long long* __restrict__ new_off, int64_t nnew,
                       const double4* __restrict__ old_xy, const int32_t* __restrict__ old_idx, const double2* __restrict__ dofs_old,
                       double2* __restrict__ dofs_new) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnew) return;
  const double4 nw = ldg4(new_xy + i);
  const double ln = seg_len4(nw);
  double a0 = 0.0, a1 = 0.0; bool initialize = true;
  for (long long p = new_off[i]; p < new_off[i + 1]; ++p) {
    const double4 od = ldg4(old_xy + p);
    const double2 u = ldg2(dofs_old + old_idx[p]);
    const double lo = seg_len4(od);
    if (lo > ln) {
      const double t0 = seg_local4(od, nw.x, nw.y), t1 = seg_local4(od, nw.z, nw.w);
      a0 = (1.0 - t0) * u.x + t0 * u.y;
      a1 = (1.0 - t1) * u.x + t1 * u.y;
    } else {
      const double ta = seg_local4(nw, od.x, od.y), tb = seg_local4(nw, od.z, od.w);
      const double pa0 = 1.0 - ta, pb0 = 1.0 - tb, pa1 = ta, pb1 = tb;
      const double w = lo * (1.0 / 6.0);
      const double r0 = w * (((2.0 * u.x) * pa0 + u.x * pb0) + (u.y * pa0 + (2.0 * u.y) * pb0));
      const double r1 = w * (((2.0 * u.x) * pa1 + u.x * pb1) + (u.y * pa1 + (2.0 * u.y) * pb1));
      const double f = 2.0 / ln;
      const double c0 = f * (2.0 * r0 - r1), c1 = f * (2.0 * r1 - r0);
      if (initialize) { a0 = c0; a1 = c1; } else { a0 += c0; a1 += c1; }
    }
    initialize = false;
  }
  dofs_new[i] = make_double2(a0, a1);
}
// LongestEdgeRefinement::refinement (longestedgerefinement.hh:41-57): centre of the longest edge of the listed cells,
// edge.geometry().center() = origin + 0.5 * (second - first) on the id-ordered corners (DUNE edge numbering)
__global__ void __launch_bounds__(256)
k_refinement_points(const double2* __restrict__ xy, const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1,
                    const int32_t* __restrict__ cv2, const uint8_t* __restrict__ perm, const uint8_t* __restrict__ longest,
                    const int32_t* __restrict__ cells, int64_t n, double2* __restrict__ out, int2* __restrict__ edge_vertices) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cells[i];
  const int v[3] = { cv0[c], cv1[c], cv2[c] };
  const unsigned p = perm[c];
  const int k[3] = { v[p & 3], v[(p >> 2) & 3], v[(p >> 4) & 3] };
  const int e = longest[c];
  const int ia = e == 2 ? 1 : 0, ib = e == 0 ? 1 : 2;            // edges 0:(0,1) 1:(0,2) 2:(1,2)
  const double2 a = ldg2(xy + k[ia]), b = ldg2(xy + k[ib]);
  out[i] = make_double2(a.x + 0.5 * (b.x - a.x), a.y + 0.5 * (b.y - a.y));
  if (edge_vertices) edge_vertices[i] = make_int2(k[ia], k[ib]);
}

// Refinement candidates of MMesh::adapt() (mmesh.hh:828-851): for every element with mark == 1, in index order, the
// longest edge is inserted unless it is an interface edge or was already inserted by an earlier element.  An edge has two
// cells, so "already inserted" <=> the face neighbour across that edge has a smaller index, is marked for refinement and
// picked the same edge.  flag[c] = 1 for exactly the cells whose insertion point enters insert_.
__device__ __forceinline__ int pick3(int i, int x0, int x1, int x2) { return i == 0 ? x0 : (i == 1 ? x1 : x2); }
struct IdCorners { int k0, k1, k2, t0, t1, t2; };     // id-sorted vertex ids and their TDS positions
__device__ __forceinline__ IdCorners id_corners(const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1,
                                                const int32_t* __restrict__ cv2, const uint8_t* __restrict__ perm, int64_t c) {
  const int v0 = cv0[c], v1 = cv1[c], v2 = cv2[c];
  const unsigned p = perm[c];
  IdCorners r;
  r.t0 = p & 3; r.t1 = (p >> 2) & 3; r.t2 = (p >> 4) & 3;
  r.k0 = pick3(r.t0, v0, v1, v2); r.k1 = pick3(r.t1, v0, v1, v2); r.k2 = pick3(r.t2, v0, v1, v2);
  return r;
}
__global__ void __launch_bounds__(256)
k_refine_candidates(const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1, const int32_t* __restrict__ cv2,
                    const uint8_t* __restrict__ perm, const int8_t* __restrict__ mark, const uint8_t* __restrict__ longest,
                    const int32_t* __restrict__ nb /*[nc][3]*/, const uint8_t* __restrict__ v_if, const unsigned long long* __restrict__ facet_keys,
                    int ns, int64_t nc, uint8_t* __restrict__ flag) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  uint8_t f = 0;
  if (mark[c] == 1) {
    const IdCorners K = id_corners(cv0, cv1, cv2, perm, c);
    const int e = longest[c];
    const int ia = e == 2 ? 1 : 0, ib = e == 0 ? 1 : 2, io = 3 - ia - ib;      // DUNE edges 0:(0,1) 1:(0,2) 2:(1,2)
    const int a = pick3(ia, K.k0, K.k1, K.k2), b = pick3(ib, K.k0, K.k1, K.k2);
    bool iface = false;                                                         // MMesh::isInterface(edge), mmesh.hh:607-621
    if (v_if[a] && v_if[b] && ns > 0) {
      const unsigned long long key = ((unsigned long long)(unsigned)min(a, b) << 32) | (unsigned)max(a, b);
      int lo = 0, hi = ns - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const unsigned long long km = facet_keys[mid];
        if (km == key) { iface = true; break; }
        if (km < key) lo = mid + 1; else hi = mid - 1;
      }
    }
    if (!iface) {
      f = 1;
      const int n = nb[3 * c + pick3(io, K.t0, K.t1, K.t2)];                    // neighbour opposite the remaining corner
      if (n >= 0 && n < c && mark[n] == 1) {
        const IdCorners N = id_corners(cv0, cv1, cv2, perm, n);
        const int en = longest[n];
        const int ja = en == 2 ? 1 : 0, jb = en == 0 ? 1 : 2;
        const int na = pick3(ja, N.k0, N.k1, N.k2), nbv = pick3(jb, N.k0, N.k1, N.k2);
        if ((na == a && nbv == b) || (na == b && nbv == a)) f = 0;             // inserted_ already holds this edge id
      }
    }
  }
  flag[c] = f;
}
// Coarsening candidates of MMesh::adapt() (mmesh.hh:853-865).  LongestEdgeRefinement::coarsening
// (longestedgerefinement.hh:64-112) sums, per id-ordered corner, the lengths of its two element edges (edge order 0,1,2)
// and std::sorts the three (vertex, length) pairs with a comparator that is NOT a strict weak order; libstdc++ sorts
// <= 16 elements by insertion sort, where element i reaches the front iff comp(*i, *first), so the returned vertex is
// the fold   y = comp(x1, x0) ? x1 : x0;   result = comp(x2, y) ? x2 : y.
__device__ __forceinline__ bool coarsen_less(int la, double sa, bool ca, int lb, double sb, bool cb) {
  if (la < lb) return true;                         // :95-97
  if (sa < sb - 1e-14) return true;                 // :100
  if (!ca && cb) return true;                       // :103-106
  return false;
}
// pass 1: chosen[c] = vertex picked by a cell with mark -1 (else -1); first_claim[v] = lowest such cell index
__global__ void __launch_bounds__(256)
k_coarsen_choice(const double2* __restrict__ xy, const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1,
                 const int32_t* __restrict__ cv2, const uint8_t* __restrict__ perm, const int8_t* __restrict__ mark,
                 const uint8_t* __restrict__ v_if, const uint8_t* __restrict__ v_bnd, const int32_t* __restrict__ v_level,
                 int64_t nc, int32_t* __restrict__ chosen, int32_t* __restrict__ first_claim) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int v = -1;
  if (mark[c] == -1) {
    const IdCorners K = id_corners(cv0, cv1, cv2, perm, c);
    const double2 p0 = ldg2(xy + K.k0), p1 = ldg2(xy + K.k1), p2 = ldg2(xy + K.k2);
    const double e0 = edge_len2d(p0, p1), e1 = edge_len2d(p0, p2), e2 = edge_len2d(p1, p2);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    s0 += e0; s1 += e0;  s0 += e1; s2 += e1;  s1 += e2; s2 += e2;      // :80-90, edges 0:(0,1) 1:(0,2) 2:(1,2)
    const int l0 = v_level[K.k0], l1 = v_level[K.k1], l2 = v_level[K.k2];
    const bool c0 = v_if[K.k0] | v_bnd[K.k0], c1 = v_if[K.k1] | v_bnd[K.k1], c2 = v_if[K.k2] | v_bnd[K.k2];
    const bool f1 = coarsen_less(l1, s1, c1, l0, s0, c0);
    const int ly = f1 ? l1 : l0; const double sy = f1 ? s1 : s0; const bool cy = f1 ? c1 : c0; const int ky = f1 ? K.k1 : K.k0;
    v = coarsen_less(l2, s2, c2, ly, sy, cy) ? K.k2 : ky;
    atomicMin(first_claim + v, (int)c);
  }
  chosen[c] = v;
}
// pass 2: removed_.insert(id(vertex)).second  <=>  this cell is the first (lowest index) one that picked the vertex
__global__ void __launch_bounds__(256)
k_coarsen_flag(const int32_t* __restrict__ chosen, const int32_t* __restrict__ first_claim, int64_t nc, uint8_t* __restrict__ flag) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int v = chosen[c];
  flag[c] = (v >= 0 && first_claim[v] == (int)c) ? 1 : 0;
}
__global__ void __launch_bounds__(256)
k_gather_i32(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int64_t n, int32_t* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
// interface loop (mmesh.hh:869-908): mark -1 -> first vertex of the element with exactly two incident interface
// elements (isRemoveable, longestedgerefinement.hh:160-168) unless the bulk loop already put it into removed_
__global__ void __launch_bounds__(256)
k_if_coarsen_choice(const int2* __restrict__ segs, const int8_t* __restrict__ imark, const int32_t* __restrict__ v_ifcount,
                    const int32_t* __restrict__ first_claim, int ns, int32_t* __restrict__ chosen, int32_t* __restrict__ first_seg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  int v = -1;
  if (imark[s] == -1) {
    const int2 sv = segs[s];
    v = v_ifcount[sv.x] == 2 ? sv.x : (v_ifcount[sv.y] == 2 ? sv.y : -1);
    if (v >= 0 && first_claim[v] != 0x7f7f7f7f) v = -1;
    if (v >= 0) atomicMin(first_seg + v, s);
  }
  chosen[s] = v;
}
__global__ void __launch_bounds__(256)
k_flag_boundary_vertices(const int4* __restrict__ edges, int64_t ne, uint8_t* __restrict__ v_bnd) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int4 q = edges[e];
  if (q.w < 0) { v_bnd[q.x] = 1; v_bnd[q.y] = 1; }          // hull edge: both ends are incident to the infinite vertex
}
__global__ void __launch_bounds__(256)
k_count_interface_elements(const int32_t* __restrict__ facets, int64_t n_entries, int32_t* __restrict__ v_ifcount) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_entries) atomicAdd(v_ifcount + facets[i], 1);
}
// ---- 3-D refinement candidates: an edge has many cells, so inserted_ becomes an open-addressing table keyed by the
// (min,max) vertex pair whose value is the lowest claiming cell index (atomicMin => independent of the thread order)
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
__device__ __forceinline__ int pick4(int i, int x0, int x1, int x2, int x3) { return i == 0 ? x0 : (i == 1 ? x1 : (i == 2 ? x2 : x3)); }
// the two vertices (id order) of DUNE tetrahedron edge e: 0:(0,1) 1:(0,2) 2:(1,2) 3:(0,3) 4:(1,3) 5:(2,3)
__device__ __forceinline__ int2 tet_edge_vertices(const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1,
                                                  const int32_t* __restrict__ cv2, const int32_t* __restrict__ cv3,
                                                  const uint8_t* __restrict__ perm, int64_t c, int e) {
  const int v0 = cv0[c], v1 = cv1[c], v2 = cv2[c], v3 = cv3[c];
  const unsigned p = perm[c];
  const int ia = (e == 2 || e == 4) ? 1 : (e == 5 ? 2 : 0);
  const int ib = e == 0 ? 1 : ((e == 1 || e == 2) ? 2 : 3);
  const int ta = (p >> (2 * ia)) & 3, tb = (p >> (2 * ib)) & 3;
  return make_int2(pick4(ta, v0, v1, v2, v3), pick4(tb, v0, v1, v2, v3));
}
__global__ void __launch_bounds__(256)
k_refine3d_claim(const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1, const int32_t* __restrict__ cv2,
                 const int32_t* __restrict__ cv3, const uint8_t* __restrict__ perm, const int8_t* __restrict__ mark,
                 const uint8_t* __restrict__ longest, const uint8_t* __restrict__ v_if,
                 const unsigned long long* __restrict__ iface_keys, int n_keys, int64_t nc,
                 unsigned long long* __restrict__ tab_key, int32_t* __restrict__ tab_val, unsigned mask, int32_t* __restrict__ slot) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int sl = -1;
  if (mark[c] == 1) {
    const int2 ab = tet_edge_vertices(cv0, cv1, cv2, cv3, perm, c, longest[c]);
    const unsigned long long key = ((unsigned long long)(unsigned)min(ab.x, ab.y) << 32) | (unsigned)max(ab.x, ab.y);
    bool iface = false;                                     // isInterface(edge) in 3-D: an edge of an interface triangle
    if (v_if[ab.x] && v_if[ab.y] && n_keys > 0) {
      int lo = 0, hi = n_keys - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const unsigned long long km = iface_keys[mid];
        if (km == key) { iface = true; break; }
        if (km < key) lo = mid + 1; else hi = mid - 1;
      }
    }
    if (!iface) {
      unsigned h = (unsigned)mix64(key) & mask;
      for (;;) {
        const unsigned long long prev = atomicCAS(tab_key + h, ~0ull, key);
        if (prev == ~0ull || prev == key) { atomicMin(tab_val + h, (int)c); sl = (int)h; break; }
        h = (h + 1) & mask;
      }
    }
  }
  slot[c] = sl;
}
__global__ void __launch_bounds__(256)
k_refine3d_flag(const int32_t* __restrict__ slot, const int32_t* __restrict__ tab_val, int64_t nc, uint8_t* __restrict__ flag) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int sl = slot[c];
  flag[c] = (sl >= 0 && tab_val[sl] == (int)c) ? 1 : 0;
}
__global__ void __launch_bounds__(256)
k_refinement_points3d(const double* __restrict__ x, const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1,
                      const int32_t* __restrict__ cv2, const int32_t* __restrict__ cv3, const uint8_t* __restrict__ perm,
                      const uint8_t* __restrict__ longest, const int32_t* __restrict__ cells, int64_t n,
                      double* __restrict__ out, int2* __restrict__ edge_vertices) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cells[i];
  const int2 ab = tet_edge_vertices(cv0, cv1, cv2, cv3, perm, c, longest[c]);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double a = x[3 * (int64_t)ab.x + k], b = x[3 * (int64_t)ab.y + k];
    out[3 * i + k] = a + 0.5 * (b - a);                     // edge.geometry().center() = origin + J^T (1/2)
  }
  if (edge_vertices) edge_vertices[i] = ab;
}
__global__ void __launch_bounds__(256)
k_flag_interface_vertices(const int32_t* __restrict__ facets, int64_t n_entries, uint8_t* __restrict__ v_if) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_entries) v_if[facets[i]] = 1;
}

// ---- component numbers of adapt_() (mmesh.hh:1194-1218 with getComponentNumber_ :1919-1941) -----------------------------
// The reference visits the zones (conflict zone of every insertion, star of every removal) in order, gives a zone the
// smallest number found on its cells or a fresh one, writes it to all its cells, and repeats the whole pass while a zone saw
// two different numbers.  Its fix point in closed form: a zone is FRESH iff none of its cells belongs to an earlier zone;
// fresh zones are numbered consecutively in zone order; every zone / cell ends with the smallest number of its connected
// group (zones connected through shared cells).  => first-zone-per-cell by atomicMin, an exclusive scan over the fresh
// flags, then min-label propagation zones <-> cells until nothing changes.
__global__ void __launch_bounds__(256)
k_cc_first(const long long* __restrict__ zone_off, const int32_t* __restrict__ zone_cells, int64_t nz, int32_t* __restrict__ cell_first) {
  const int64_t z = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= nz) return;
  for (long long e = zone_off[z]; e < zone_off[z + 1]; ++e) atomicMin(cell_first + zone_cells[e], (int)z);
}
__global__ void __launch_bounds__(256)
k_cc_fresh(const long long* __restrict__ zone_off, const int32_t* __restrict__ zone_cells, int64_t nz,
           const int32_t* __restrict__ cell_first, uint32_t* __restrict__ fresh, uint8_t* __restrict__ fresh_flag) {
  const int64_t z = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= nz) return;
  bool f = true;
  for (long long e = zone_off[z]; e < zone_off[z + 1]; ++e) f = f && (cell_first[zone_cells[e]] == (int)z);
  fresh[z] = f ? 1u : 0u; fresh_flag[z] = f ? 1 : 0;
}
__global__ void __launch_bounds__(256)
k_cc_init(const uint32_t* __restrict__ prefix, const uint8_t* __restrict__ fresh_flag, int64_t nz, int base, int32_t* __restrict__ label) {
  const int64_t z = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (z < nz) label[z] = fresh_flag[z] ? base + (int)prefix[z] : 0x7fffffff;
}
__global__ void __launch_bounds__(256)
k_cc_to_cells(const long long* __restrict__ zone_off, const int32_t* __restrict__ zone_cells, int64_t nz,
              const int32_t* __restrict__ label, int32_t* __restrict__ cell_label) {
  const int64_t z = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= nz) return;
  const int l = label[z];
  for (long long e = zone_off[z]; e < zone_off[z + 1]; ++e) atomicMin(cell_label + zone_cells[e], l);
}
__global__ void __launch_bounds__(256)
k_cc_to_zones(const long long* __restrict__ zone_off, const int32_t* __restrict__ zone_cells, int64_t nz,
              const int32_t* __restrict__ cell_label, int32_t* __restrict__ label, int* __restrict__ changed) {
  const int64_t z = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= nz) return;
  int m = label[z];
  for (long long e = zone_off[z]; e < zone_off[z + 1]; ++e) m = min(m, cell_label[zone_cells[e]]);
  if (m < label[z]) { label[z] = m; *changed = 1; }
}
__global__ void __launch_bounds__(256)
k_cc_finish(int32_t* __restrict__ cell_label, int64_t nc) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nc && cell_label[c] >= 0x7f7f7f7f) cell_label[c] = 0;              // not in any zone (still the memset pattern): componentNumber stays 0
}

// halo exchange helpers: owned values -> contiguous send buffer, received values -> halo slots
__global__ void __launch_bounds__(256)
k_gather2(const double2* __restrict__ src, const int32_t* __restrict__ idx, int64_t n, double2* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
__global__ void __launch_bounds__(256)
k_scatter2(double2* __restrict__ dst, const int32_t* __restrict__ idx, int64_t n, const double2* __restrict__ src) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[i];
}
// scatter of packed DOF blocks (ncomp doubles each) to their cell slots: the few old DOFs a projection chunk needs from
// outside its own index range are uploaded first through a packed staging buffer
__global__ void __launch_bounds__(256)
k_scatter_dofs(double* __restrict__ dst, const int32_t* __restrict__ idx, int64_t n, int ncomp, const double* __restrict__ src) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ncomp) return;
  const int64_t e = i / ncomp; const int c = (int)(i - e * ncomp);
  dst[(int64_t)idx[e] * ncomp + c] = src[i];
}
// the additive scalars of a step as one double vector (counts < 2^53 are exact), ready for a single all-reduce
__global__ void k_reduce_vector(const Counters* __restrict__ c, double* __restrict__ v) {
  if (threadIdx.x == 0) {
    v[0] = (double)c->n_invalid; v[1] = (double)c->n_orient_nonpos; v[2] = (double)c->n_illegal;
    v[3] = (double)c->n_refine; v[4] = (double)c->n_coarsen; v[5] = (double)c->n_exact_orient;
    v[6] = (double)c->n_exact_incircle; v[7] = (double)c->n_pieces; v[8] = c->mass_new; v[9] = c->covered_area;
  }
}


// mmb_patch_mesh: scatter of the changed vertices / cells / edges into the resident tables
__global__ void __launch_bounds__(256)
k_patch_vertices(double* __restrict__ x, int32_t* __restrict__ hint, int dim, const int32_t* __restrict__ idx, const double* __restrict__ val, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = idx[i];
  for (int c = 0; c < dim; ++c) x[v * dim + c] = val[i * dim + c];
  hint[v] = -1;                                              // the vertex moved or is new: no usable hint
}
__global__ void __launch_bounds__(256)
k_patch_cells(int32_t* __restrict__ aos, int32_t* p0, int32_t* p1, int32_t* p2, int32_t* p3, uint8_t* __restrict__ perm, int k,
              const int32_t* __restrict__ idx, const int32_t* __restrict__ val, const uint8_t* __restrict__ pval, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = idx[i];
  for (int j = 0; j < k; ++j) aos[k * c + j] = val[k * i + j];
  p0[c] = val[k * i]; p1[c] = val[k * i + 1]; p2[c] = val[k * i + 2];
  if (k == 4) p3[c] = val[k * i + 3];
  perm[c] = pval[i];
}
__global__ void __launch_bounds__(256)
k_patch_edges(int4* __restrict__ edges, const int32_t* __restrict__ idx, const int4* __restrict__ val, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) edges[idx[i]] = val[i];
}

// AoS -> SoA transposition of the uploaded cell table
__global__ void __launch_bounds__(256)
k_cells_to_planes(const int32_t* __restrict__ aos, int k, int64_t nc, int64_t nc_pad, int32_t* p0, int32_t* p1,
                  int32_t* p2, int32_t* p3) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc_pad) return;
  const bool live = c < nc;
  p0[c] = live ? aos[k * c] : 0; p1[c] = live ? aos[k * c + 1] : 0; p2[c] = live ? aos[k * c + 2] : 0;
  if (k == 4) p3[c] = live ? aos[k * c + 3] : 0;
}

// ------------------------------------------------------------------------------------------------------
// F4  Delaunay restoration by edge flips (flip.cuh: mark / select / apply) and the derived edge list.
//     CGAL/Delaunay_triangulation_2.h:936-1022 (which edges flip), CGAL/Triangulation_data_structure_2.h:885-915
//     (the flip).  Off the timed step: runs between steps when the legality pass reports illegal edges.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_flip_mark(const double* __restrict__ xy, const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1, const int32_t* __restrict__ cv2,
            const int32_t* __restrict__ nb, const uint8_t* __restrict__ v_if, const unsigned long long* __restrict__ keys, int nkeys,
            int64_t nc, const uint8_t* __restrict__ active, unsigned* __restrict__ claim, uint8_t* __restrict__ ill, FlipCounts* __restrict__ fc) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  // only cells whose own or neighbouring triangles changed in the last round (or whose illegal edge was deferred) can own an
  // illegal edge: every other cell was legal when it was last tested and nothing it depends on has changed since
  if (!active[c]) { ill[c] = 0; return; }
  unsigned cnt[5] = { 0, 0, 0, 0, 0 };
  ill[c] = (uint8_t)flip_mark_cell(c, xy, cv0, cv1, cv2, nb, v_if, keys, nkeys,
                                   [&](int64_t cell, unsigned id) { atomicMin(claim + cell, id); }, [&](int which) { cnt[which]++; });
  if (cnt[0]) atomicAdd(&fc->illegal, (unsigned long long)cnt[0]);
  if (cnt[1]) atomicAdd(&fc->constrained, (unsigned long long)cnt[1]);
  if (cnt[2]) atomicAdd(&fc->inverted, (unsigned long long)cnt[2]);
  if (cnt[3]) atomicAdd(&fc->broken, (unsigned long long)cnt[3]);
  if (cnt[4]) atomicAdd(&fc->domain, (unsigned long long)cnt[4]);
}
__global__ void __launch_bounds__(256)
k_flip_select(const int32_t* __restrict__ nb, const unsigned* __restrict__ claim, const uint8_t* __restrict__ ill, int64_t nc,
              uint8_t* __restrict__ sel) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nc) sel[c] = (uint8_t)flip_select_cell(c, ill[c], nb, claim);
}
__global__ void __launch_bounds__(256)
k_flip_apply(int32_t* cv0, int32_t* cv1, int32_t* cv2, int32_t* nb, int32_t* __restrict__ aos, const int64_t* __restrict__ vid,
             uint8_t* perm, const uint8_t* __restrict__ sel, const uint8_t* __restrict__ ill, uint8_t* __restrict__ active_next, int64_t nc,
             FlipCounts* __restrict__ fc, uint32_t* __restrict__ log, unsigned long long log_cap) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k1 = c < nc ? sel[c] : 0;
  if (k1) {
    const int32_t n = nb[3 * c + (k1 - 1)];
    flip_apply_cell(c, k1 - 1, cv0, cv1, cv2, nb, vid, perm);
    aos[3 * c] = cv0[c]; aos[3 * c + 1] = cv1[c]; aos[3 * c + 2] = cv2[c];
    aos[3 * (int64_t)n] = cv0[n]; aos[3 * (int64_t)n + 1] = cv1[n]; aos[3 * (int64_t)n + 2] = cv2[n];
    // next round: the two new cells and everything around them (an edge is tested by its lower cell, which may be a neighbour)
    active_next[c] = 1; active_next[n] = 1;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int32_t a = nb[3 * c + j], b = nb[3 * (int64_t)n + j];
      if (a >= 0) active_next[a] = 1;
      if (b >= 0) active_next[b] = 1;
    }
  } else if (c < nc && ill[c]) {
    active_next[c] = 1;                                   // an illegal edge that had to wait for a neighbouring flip
  }
  // flip counter + log of (cell, slot) = the arguments of TDS::flip(f, i), one atomic per warp
  const unsigned m = __ballot_sync(0xffffffffu, k1 != 0);
  unsigned long long base = 0;
  if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(&fc->flips, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (k1) {
    const unsigned long long pos = base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
    if (pos < log_cap) log[pos] = 3u * (uint32_t)c + (uint32_t)(k1 - 1);
  }
}
// edge list from cells + neighbours: per-block counts (COMPACT_TILE cells per block), k_compact_scan, emission
__global__ void __launch_bounds__(256)
k_edges_count(const int32_t* __restrict__ nb, int64_t nc, uint32_t* __restrict__ block_count) {
  const int64_t base = (int64_t)blockIdx.x * COMPACT_TILE;
  unsigned c = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int64_t i = base + r * 256 + threadIdx.x;
    if (i < nc) c += (unsigned)flip_edge_count(i, nb);
  }
  __shared__ unsigned sm[8];
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { unsigned t = 0; for (int w = 0; w < 8; ++w) t += sm[w]; block_count[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(256)
k_edges_emit(const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1, const int32_t* __restrict__ cv2, const int32_t* __restrict__ nb,
             int64_t nc, const uint32_t* __restrict__ block_off, int4* __restrict__ edges, int64_t cap) {
  const int64_t base = (int64_t)blockIdx.x * COMPACT_TILE;
  __shared__ unsigned wcount[8];
  unsigned off = block_off[blockIdx.x];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < 8; ++r) {
    const int64_t i = base + r * 256 + threadIdx.x;
    int32_t q[12];
    const unsigned m = i < nc ? (unsigned)flip_edge_emit(i, cv0, cv1, cv2, nb, q) : 0u;
    unsigned incl = m;                                     // inclusive warp scan of the per-cell counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wcount[w] = incl;
    __syncthreads();
    unsigned before = 0, tot = 0;
    for (int k = 0; k < 8; ++k) { const unsigned t = wcount[k]; if (k < w) before += t; tot += t; }
    const int64_t pos = (int64_t)off + before + incl - m;
    for (unsigned e = 0; e < m; ++e)
      if (pos + e < cap) edges[pos + e] = make_int4(q[4 * e], q[4 * e + 1], q[4 * e + 2], q[4 * e + 3]);
    off += tot;
    __syncthreads();
  }
}

}  // namespace mmb
