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

namespace mmb {

struct Counters {
  unsigned long long n_invalid, n_orient_nonpos, n_illegal, n_refine, n_coarsen;
  unsigned long long wl_orient, wl_incircle;      // worklist cursors
  unsigned long long n_exact_orient, n_exact_incircle;
  unsigned long long n_pieces;
  unsigned long long max_dist_bits;               // non-negative double compared as integer
  unsigned long long minmax_bits[4];              // indicator init: iface min, iface max, bulk min, bulk max
  double mass_new, covered_area;
  unsigned long long pad_[2];
};

struct DevParams {
  double minH, maxH, factor, distProportion, edgeRatio, radiusRatio, clip_eps, drop_area;
  int clip_translate;
};

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
__global__ void __launch_bounds__(256) k_move(double2* __restrict__ x, const double2* __restrict__ s, int64_t n2) {
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
// K1  trial move + validity, 2-D.  mmesh.hh:1094-1134 (ensureVertexMovement): signedVolume_ <= 0.0 on
//     x (+) s in TDS vertex order (mmesh.hh:1169-1172 -> CGAL/Cartesian/function_objects.h:918-926), plus the
//     exact CGAL orientation sign (static filter here, worklist for the fallback kernel).
// ------------------------------------------------------------------------------------------------------
template <bool SHIFT>
__global__ void __launch_bounds__(256)
k_validate2d(const double2* __restrict__ xy, const double2* __restrict__ sh, const int4* __restrict__ cv0,
             const int4* __restrict__ cv1, const int4* __restrict__ cv2, int64_t nc4, int64_t nc,
             uint32_t* __restrict__ valid4, uint32_t* __restrict__ orient4, int32_t* __restrict__ wl,
             Counters* __restrict__ cnt) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned ninv = 0, nnonpos = 0;
  bool need[4] = { false, false, false, false };
  if (g < nc4) {
    const int4 a = __ldg(cv0 + g), b = __ldg(cv1 + g), c = __ldg(cv2 + g);
    const int va[4] = { a.x, a.y, a.z, a.w }, vb[4] = { b.x, b.y, b.z, b.w }, vc[4] = { c.x, c.y, c.z, c.w };
    double2 P[4], Q[4], R[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { P[j] = ldg2(xy + va[j]); Q[j] = ldg2(xy + vb[j]); R[j] = ldg2(xy + vc[j]); }
    if (SHIFT) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double2 sp = ldg2(sh + va[j]), sq = ldg2(sh + vb[j]), sr = ldg2(sh + vc[j]);
        P[j].x = P[j].x + sp.x; P[j].y = P[j].y + sp.y;     // mmesh.hh:1427 point = center() + shift
        Q[j].x = Q[j].x + sq.x; Q[j].y = Q[j].y + sq.y;
        R[j].x = R[j].x + sr.x; R[j].y = R[j].y + sr.y;
      }
    }
    uint32_t vbits = 0, obits = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double det;
      int s = orient2d_filter(P[j].x, P[j].y, Q[j].x, Q[j].y, R[j].x, R[j].y, det);
      const double area = det / 2;                           // Compute_area_2: determinant(...)/2
      const bool ok = !(area <= 0.0);                        // mmesh.hh:1102
      const bool live = 4 * g + j < nc;
      need[j] = live && (s == SIGN_UNKNOWN);
      if (s == SIGN_UNKNOWN) s = 0;
      vbits |= (uint32_t)(ok ? 1u : 0u) << (8 * j);
      obits |= ((uint32_t)(uint8_t)(int8_t)s) << (8 * j);
      ninv += (live && !ok);
      nnonpos += (live && !need[j] && s <= 0);
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
                 int8_t* __restrict__ orient, Counters* __restrict__ cnt) {
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
    orient[c] = (int8_t)s;
    nnonpos += (s <= 0);
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
           uint8_t* __restrict__ flag, int32_t* __restrict__ wl, Counters* __restrict__ cnt) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned nill = 0;
  bool need = false;
  if (e < ne) {
    const int4 q = __ldg(edges + e);
    uint8_t f = 2;
    if (q.w >= 0) {
      double2 A = ldg2(xy + q.x), B = ldg2(xy + q.y), C = ldg2(xy + q.z), D = ldg2(xy + q.w);
      if (SHIFT) {
        const double2 sa = ldg2(sh + q.x), sb = ldg2(sh + q.y), sc = ldg2(sh + q.z), sd = ldg2(sh + q.w);
        A.x = A.x + sa.x; A.y = A.y + sa.y; B.x = B.x + sb.x; B.y = B.y + sb.y;
        C.x = C.x + sc.x; C.y = C.y + sc.y; D.x = D.x + sd.x; D.y = D.y + sd.y;
      }
      const int s = incircle_filter(A.x, A.y, B.x, B.y, C.x, C.y, D.x, D.y);
      need = (s == SIGN_UNKNOWN);
      f = (s > 0 && !need) ? 0 : 1;
      nill += (f == 0);
    }
    flag[e] = f;
  }
  wl_push(need, (int)e, wl, &cnt->wl_incircle);
  block_count_add<256>(nill, 0u, &cnt->n_illegal, &cnt->n_illegal);
}
template <bool SHIFT>
__global__ void __launch_bounds__(128)
k_exact_incircle(const double2* __restrict__ xy, const double2* __restrict__ sh, const int4* __restrict__ edges,
                 const int32_t* __restrict__ wl, uint8_t* __restrict__ flag, Counters* __restrict__ cnt) {
  const unsigned long long n = cnt->wl_incircle;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned nill = 0, nexact = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int e = wl[i];
    const int4 q = __ldg(edges + e);
    double2 A = ldg2(xy + q.x), B = ldg2(xy + q.y), C = ldg2(xy + q.z), D = ldg2(xy + q.w);
    if (SHIFT) {
      const double2 sa = ldg2(sh + q.x), sb = ldg2(sh + q.y), sc = ldg2(sh + q.z), sd = ldg2(sh + q.w);
      A.x = A.x + sa.x; A.y = A.y + sa.y; B.x = B.x + sb.x; B.y = B.y + sb.y;
      C.x = C.x + sc.x; C.y = C.y + sc.y; D.x = D.x + sd.x; D.y = D.y + sd.y;
    }
    int s = incircle_interval(A.x, A.y, B.x, B.y, C.x, C.y, D.x, D.y);
    if (s == SIGN_UNKNOWN) { s = incircle_exact(A.x, A.y, B.x, B.y, C.x, C.y, D.x, D.y); nexact++; }
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
    sign[i] = (int8_t)s;
  }
  block_count_add<128>(nexact, 0u, &cnt->n_exact_orient, &cnt->n_exact_orient);
}

// ------------------------------------------------------------------------------------------------------
// K3  interface distance field, 2-D.  Per-pair arithmetic = dune/mmesh/remeshing/distance.hh:143-156 and
//     Plane (dune/mmesh/cgal/triangulationwrapper.hh:14-40); min over ALL facets (distance.hh:42-59,133-140) is
//     obtained with an AABB tree whose culling is conservative, so the minimum has the brute-force bits.
// ------------------------------------------------------------------------------------------------------
struct SegLeaf { double c0x, c0y, c1x, c1y, n0, n1; };   // facet corners (AffineGeometry::corner) + Plane normal

// leaves in Morton order of the (upload-time) midpoints; boxes: node i in [1, 2L), children 2i, 2i+1
__global__ void __launch_bounds__(256)
k_seg_prepare(const double2* __restrict__ xy, const int2* __restrict__ segs_sorted, int ns, int L,
              SegLeaf* __restrict__ leaf, double4* __restrict__ box) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  double4 b = make_double4(INFINITY, INFINITY, -INFINITY, -INFINITY);
  if (i < ns) {
    const int2 s = segs_sorted[i];
    const double2 v0 = ldg2(xy + s.x), v1 = ldg2(xy + s.y);
    SegLeaf f;
    f.c0x = v0.x; f.c0y = v0.y;
    f.c1x = v0.x + 1.0 * (v1.x - v0.x);                    // corner(1) = origin + J^T e_1
    f.c1y = v0.y + 1.0 * (v1.y - v0.y);
    double n0 = f.c0x - f.c1x, n1 = f.c0y - f.c1y;         // n_ = a - b
    { const double t = n0; n0 = n1; n1 = t; }              // swap
    n0 *= -1.0;
    double nn = 0.0; nn += n0 * n0; nn += n1 * n1; nn = sqrt(nn);
    f.n0 = n0 / nn; f.n1 = n1 / nn;
    leaf[i] = f;
    b = make_double4(fmin(f.c0x, f.c1x), fmin(f.c0y, f.c1y), fmax(f.c0x, f.c1x), fmax(f.c0y, f.c1y));
  }
  box[L + i] = b;
}
__global__ void __launch_bounds__(1024) k_bvh_upper(int L, double4* __restrict__ box) {
  for (int lvl = L >> 1; lvl >= 1; lvl >>= 1) {
    for (int i = lvl + threadIdx.x; i < 2 * lvl; i += blockDim.x) {
      const double4 a = box[2 * i], b = box[2 * i + 1];
      box[i] = make_double4(fmin(a.x, b.x), fmin(a.y, b.y), fmax(a.z, b.z), fmax(a.w, b.w));
    }
    __syncthreads();
  }
}
__device__ __forceinline__ double seg_distance(const SegLeaf& f, double px, double py) {
  {  // i = 0: vi = c0, vj = c1
    const double ax = f.c0x - f.c1x, ay = f.c0y - f.c1y, bx = px - f.c0x, by = py - f.c0y;
    double sgn = 0.0; sgn += ax * bx; sgn += ay * by;
    if (sgn >= 0) { double n2 = 0.0; n2 += bx * bx; n2 += by * by; return sqrt(n2); }
  }
  {  // i = 1: vi = c1, vj = c0
    const double ax = f.c1x - f.c0x, ay = f.c1y - f.c0y, bx = px - f.c1x, by = py - f.c1y;
    double sgn = 0.0; sgn += ax * bx; sgn += ay * by;
    if (sgn >= 0) { double n2 = 0.0; n2 += bx * bx; n2 += by * by; return sqrt(n2); }
  }
  const double dx = px - f.c0x, dy = py - f.c0y;
  double sd = 0.0; sd += f.n0 * dx; sd += f.n1 * dy;
  return fabs(sd);
}
__device__ __forceinline__ double box_dist2(const double4 b, double px, double py) {
  const double dx = fmax(fmax(b.x - px, px - b.z), 0.0), dy = fmax(fmax(b.y - py, py - b.w), 0.0);
  return dx * dx + dy * dy;
}
__global__ void __launch_bounds__(256)
k_distance2d(const double2* __restrict__ xy, int64_t nv, const SegLeaf* __restrict__ leaf,
             const double4* __restrict__ box, int ns, int L, double* __restrict__ dist, Counters* __restrict__ cnt) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double best = 1e100;
  if (v < nv && ns > 0) {
    const double2 p = ldg2(xy + v);
    const double4 root = ldg4(box + 1);
    const double diag = (root.z - root.x) + (root.w - root.y);   // >= any facet length
    // conservative cull threshold: a node whose box distance exceeds T cannot hold a facet whose *computed*
    // distance is below `best` (computed vs true distance differ by O(1e-15 (d + len)))
    double T2;
    { const double T = best + 1e-12 * (best + diag); T2 = T * T * (1.0 + 1e-12); }
    int stack[40]; double sd2[40]; int sp = 0;
    int node = 1;
    for (;;) {
      if (node >= L) {
        const int s = node - L;
        if (s < ns) {
          const SegLeaf f = leaf[s];
          const double d = seg_distance(f, p.x, p.y);
          if (d < best) {                                           // distance.hh:137-138
            best = d;
            const double T = best + 1e-12 * (best + diag); T2 = T * T * (1.0 + 1e-12);
          }
        }
        node = 0;
      } else {
        const double dl = box_dist2(ldg4(box + 2 * node), p.x, p.y);
        const double dr = box_dist2(ldg4(box + 2 * node + 1), p.x, p.y);
        const bool lf = dl <= dr;
        const double dn = lf ? dl : dr, df = lf ? dr : dl;
        const int nn = 2 * node + (lf ? 0 : 1), nf = 2 * node + (lf ? 1 : 0);
        if (dn > T2) node = 0;
        else {
          if (!(df > T2)) { stack[sp] = nf; sd2[sp] = df; ++sp; }
          node = nn;
        }
      }
      if (node == 0) {
        while (sp > 0) { --sp; if (!(sd2[sp] > T2)) { node = stack[sp]; break; } }
        if (node == 0) break;
      }
    }
  }
  if (v < nv) dist[v] = best;
  // Distance::maximum (distance.hh:118-123): max starting from 0.0; non-negative doubles order like integers
  unsigned long long bits = (v < nv) ? (unsigned long long)__double_as_longlong(best) : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, bits, o); bits = t > bits ? t : bits; }
  __shared__ unsigned long long sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = bits;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long m = sm[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = sm[w] > m ? sm[w] : m;
    atomicMax(&cnt->max_dist_bits, m);
  }
}

// ------------------------------------------------------------------------------------------------------
// K4  RatioIndicator::operator() + LongestEdgeRefinement::refinement argmax, 2-D bulk cells.
//     dune/mmesh/remeshing/ratioindicator.hh:106-160, distance.hh:92-101, longestedgerefinement.hh:41-57,
//     circumcenter CGAL/constructions/kernel_ftC2.h:40-74, geometry dune/mmesh/grid/geometry.hh:76-104
//     (+ AffineGeometry of dune-geometry).  One evaluation per cell (the reference evaluates twice,
//     mmesh.hh:741-742).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double edge_len2d(double2 a, double2 b) {
  const double jx = b.x - a.x, jy = b.y - a.y;
  double s = 0.0; s += jx * jx; s += jy * jy;
  return sqrt(s);
}
__device__ __forceinline__ int indicator2d(const double2 c0, const double2 c1, const double2 c2, double d0, double d1,
                                           double d2, const DevParams& prm, double maxDist, int& longest) {
  int refine = 0, coarse = 0;
  double minE = 1e100, maxE = -1e100, sumE = 0.0;
  double dist = 0.0; dist += d0; dist += d1; dist += d2; dist /= 3;
  dist = (dist < maxDist) ? dist : maxDist;
  const double l = dist / maxDist;
  const double minH = (1. - l) * prm.minH + l * prm.factor * prm.minH;
  const double maxH = (1. - l) * prm.maxH + l * prm.factor * prm.maxH;
  const double len[3] = { edge_len2d(c0, c1), edge_len2d(c0, c2), edge_len2d(c1, c2) };   // DUNE edge order
  double longestLen = -1e100; longest = -1;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (len[i] < minH) coarse++;
    if (len[i] > maxH) refine++;
    minE = (len[i] < minE) ? len[i] : minE;
    maxE = (maxE < len[i]) ? len[i] : maxE;
    sumE += len[i];
    if (len[i] > longestLen) { longestLen = len[i]; longest = i; }
  }
  const double edgeRatio = maxE / minE;
  if (edgeRatio > prm.edgeRatio) coarse++;
  const double j00 = c1.x - c0.x, j01 = c1.y - c0.y, j10 = c2.x - c0.x, j11 = c2.y - c0.y;
  const double vol = fabs(j00 * j11 - j10 * j01) * 0.5;
  const double g1x = (c0.x + 1.0 * j00) + 0.0 * j10, g1y = (c0.y + 1.0 * j01) + 0.0 * j11;   // corner(1)
  const double g2x = (c0.x + 0.0 * j00) + 1.0 * j10, g2y = (c0.y + 0.0 * j01) + 1.0 * j11;   // corner(2)
  const double innerRadius = 2 * vol / sumE;
  const double dqx = g1x - c0.x, dqy = g1y - c0.y, drx = g2x - c0.x, dry = g2y - c0.y;
  const double r2 = drx * drx + dry * dry;
  const double q2 = dqx * dqx + dqy * dqy;
  const double den = 2 * det2(dqx, dqy, drx, dry);
  const double dcx = det2(dry, dqy, r2, q2) / den;
  const double dcy = -det2(drx, dqx, r2, q2) / den;
  const double ccx = dcx + c0.x, ccy = dcy + c0.y;
  const double ox = ccx - c0.x, oy = ccy - c0.y;
  double o2 = 0.0; o2 += ox * ox; o2 += oy * oy;
  const double outerRadius = sqrt(o2);
  const double radiusRatio = outerRadius / (innerRadius * 2);
  if (radiusRatio > prm.radiusRatio) coarse++;
  if (coarse > 0) return -1;
  if (refine > 0) return 1;
  return 0;
}
__global__ void __launch_bounds__(256)
k_mark2d(const double2* __restrict__ xy, const double* __restrict__ dist, const int32_t* __restrict__ cv0,
         const int32_t* __restrict__ cv1, const int32_t* __restrict__ cv2, const uint8_t* __restrict__ perm, int64_t nc,
         DevParams prm, const Counters* __restrict__ cin, int8_t* __restrict__ mark, uint8_t* __restrict__ longest,
         Counters* __restrict__ cnt) {
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
    nref = (m > 0); ncoa = (m < 0);
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
constexpr int MAXPOLY = 12;   // rigorous bound for triangle x triangle with the epsilon cases is 9
struct Poly { double x[MAXPOLY], y[MAXPOLY]; int n; };

__device__ __forceinline__ void line_isect(double p1x, double p1y, double p2x, double p2y, double q1x, double q1y,
                                           double q2x, double q2y, double& ox, double& oy) {
  double dPx = p1x - p2x, dPy = p1y - p2y, dQx = q1x - q2x, dQy = q1y - q2y;
  const double scale = dPx * dQy - dQx * dPy;
  const double fq = q1x * q2y - q1y * q2x;
  const double fp = p1x * p2y - p1y * p2x;
  dPx *= fq; dPy *= fq; dQx *= fp; dQy *= fp;
  ox = (dQx - dPx) / scale; oy = (dQy - dPy) / scale;
}
__device__ __forceinline__ void sh_clip(const double cx[3], const double cy[3], const double ex[3], const double ey[3],
                                        double eps, Poly& out) {
  Poly in;
  out.n = 3;
#pragma unroll
  for (int i = 0; i < 3; ++i) { out.x[i] = cx[i]; out.y[i] = cy[i]; }
#pragma unroll 1
  for (int ci = 0; ci < 3; ++ci) {
    const int cn = ci == 2 ? 0 : ci + 1;
    const double q1x = ex[ci], q1y = ey[ci], q2x = ex[cn], q2y = ey[cn];
    const double dQx = q2x - q1x, dQy = q2y - q1y;
    const double q2xq1 = q1y * q2x - q1x * q2y;
    in = out; out.n = 0;
    for (int ii = 0; ii < in.n; ++ii) {
      const int i2 = ii + 1 == in.n ? 0 : ii + 1;
      const double first = -dQx * in.y[ii] + dQy * in.x[ii] + q2xq1;
      const double second = -dQx * in.y[i2] + dQy * in.x[i2] + q2xq1;
      if (first > eps && second < -eps) {
        line_isect(in.x[ii], in.y[ii], in.x[i2], in.y[i2], q1x, q1y, q2x, q2y, out.x[out.n], out.y[out.n]); out.n++;
        out.x[out.n] = in.x[i2]; out.y[out.n] = in.y[i2]; out.n++;
      } else if (first < -eps && second > eps) {
        line_isect(in.x[ii], in.y[ii], in.x[i2], in.y[i2], q1x, q1y, q2x, q2y, out.x[out.n], out.y[out.n]); out.n++;
      } else if (first >= -eps && second > eps) {
      } else {
        out.x[out.n] = in.x[i2]; out.y[out.n] = in.y[i2]; out.n++;
      }
      if (out.n > MAXPOLY - 2) { out.n = 0; return; }
    }
  }
}
struct Aff2 { double ox, oy, i00, i01, i10, i11, vol; };
__device__ __forceinline__ Aff2 aff2_make(double c0x, double c0y, double c1x, double c1y, double c2x, double c2y) {
  const double a00 = c1x - c0x, a01 = c1y - c0y, a10 = c2x - c0x, a11 = c2y - c0y;
  const double det = a00 * a11 - a10 * a01;
  const double detInv = 1.0 / det;
  Aff2 g;
  g.ox = c0x; g.oy = c0y;
  g.i00 = a11 * detInv; g.i10 = -a10 * detInv; g.i01 = -a01 * detInv; g.i11 = a00 * detInv;
  g.vol = fabs(det) * 0.5;
  return g;
}
__device__ __forceinline__ void aff2_local(const Aff2& g, double x, double y, double& xi, double& eta) {
  const double d0 = x - g.ox, d1 = y - g.oy;
  double l0 = 0.0, l1 = 0.0;
  l0 += d0 * g.i00; l1 += d0 * g.i01;
  l0 += d1 * g.i10; l1 += d1 * g.i11;
  xi = l0; eta = l1;
}

// cut one (old, new) pair; old6 = CachingEntity::vertex_; e = new cell TDS-order corners.
// On return poly holds the clipped polygon (possibly translated by -org when prm.clip_translate).
__device__ __forceinline__ void cut_pair(const double* __restrict__ old6, const double ex_[3], const double ey_[3],
                                         const DevParams& prm, Poly& poly, double cux[3], double cuy[3], double& orgx,
                                         double& orgy) {
  double cx[3], cy[3], ex[3], ey[3];
  const double o0x = old6[0], o0y = old6[1];
  const double j00 = old6[2] - o0x, j01 = old6[3] - o0y, j10 = old6[4] - o0x, j11 = old6[5] - o0y;
  cx[0] = o0x; cy[0] = o0y;                                           // cgeo.corner(i), cutsettriangulation.hh:40
  cx[1] = (o0x + 1.0 * j00) + 0.0 * j10; cy[1] = (o0y + 1.0 * j01) + 0.0 * j11;
  cx[2] = (o0x + 0.0 * j00) + 1.0 * j10; cy[2] = (o0y + 0.0 * j01) + 1.0 * j11;
#pragma unroll
  for (int i = 0; i < 3; ++i) { ex[i] = ex_[i]; ey[i] = ey_[i]; }
  orgx = 0.0; orgy = 0.0;
  if (prm.clip_translate) {
    orgx = ex[0]; orgy = ey[0];
#pragma unroll
    for (int i = 0; i < 3; ++i) { cx[i] -= orgx; cy[i] -= orgy; ex[i] -= orgx; ey[i] -= orgy; }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) { cux[i] = cx[i]; cuy[i] = cy[i]; }
  // cutsettriangulation.hh:46-49: orientation value truncated to int
  const double od = (cy[1] - cy[0]) * (cx[2] - cx[1]) - (cx[1] - cx[0]) * (cy[2] - cy[1]);
  const int o = (fabs(od) < 2147483648.0) ? (int)od : 0;
  if (o > 0) { double t = cx[1]; cx[1] = cx[2]; cx[2] = t; t = cy[1]; cy[1] = cy[2]; cy[2] = t; }
  sh_clip(cx, cy, ex, ey, prm.clip_eps, poly);
}

template <int NCOMP>
__global__ void __launch_bounds__(128)
k_project(const double2* __restrict__ xy, const int32_t* __restrict__ cv0, const int32_t* __restrict__ cv1,
          const int32_t* __restrict__ cv2, const uint8_t* __restrict__ perm, const long long* __restrict__ new_off,
          const int32_t* __restrict__ new_cell, int64_t nnew, const double* __restrict__ old_xy,
          const int32_t* __restrict__ old_idx, const double* __restrict__ dofs_old, double* __restrict__ dofs_new,
          DevParams prm, double* __restrict__ part_mass, double* __restrict__ part_cov, Counters* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double mass = 0.0, cov = 0.0;
  unsigned npc = 0;
  if (i < nnew) {
    const int K = new_cell[i];
    const int v[3] = { __ldg(cv0 + K), __ldg(cv1 + K), __ldg(cv2 + K) };
    const unsigned p = __ldg(perm + K);
    double ex[3], ey[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) { const double2 q = ldg2(xy + v[t]); ex[t] = q.x; ey[t] = q.y; }
    double kx[3], ky[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) { const int s = (p >> (2 * t)) & 3; kx[t] = ex[s]; ky[t] = ey[s]; }
    const double volK = aff2_make(kx[0], ky[0], kx[1], ky[1], kx[2], ky[2]).vol;   // element.geometry().volume()
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    bool initialize = true;
    for (long long pr = new_off[i]; pr < new_off[i + 1]; ++pr) {
      Poly poly; double cux[3], cuy[3], orgx, orgy;
      cut_pair(old_xy + 6 * pr, ex, ey, prm, poly, cux, cuy, orgx, orgy);
      if (poly.n < 3) continue;
      const double* uo = dofs_old + (int64_t)NCOMP * old_idx[pr];
      Aff2 gK, gO;
      double u0 = 0, u1 = 0, u2 = 0;
      if (NCOMP == 3) {
        gK = aff2_make(kx[0] - orgx, ky[0] - orgy, kx[1] - orgx, ky[1] - orgy, kx[2] - orgx, ky[2] - orgy);
        gO = aff2_make(cux[0], cuy[0], cux[1], cuy[1], cux[2], cuy[2]);
        u0 = uo[0]; u1 = uo[1]; u2 = uo[2];
      } else u0 = uo[0];
      for (int t = 1; t < poly.n - 1; ++t) {
        const double T[6] = { poly.x[0], poly.y[0], poly.x[t], poly.y[t], poly.x[t + 1], poly.y[t + 1] };
        const double a00 = T[2] - T[0], a01 = T[3] - T[1], a10 = T[4] - T[0], a11 = T[5] - T[1];
        const double aT = fabs(a00 * a11 - a10 * a01) * 0.5;
        if (!(aT > prm.drop_area)) continue;                          // cutsettriangulation.hh:60
        cov += aT; npc++;
        if (NCOMP == 1) {
          const double w = aT / volK;                                 // FV restrictLocal weight |son|/|father|
          if (initialize) acc0 = w * u0; else acc0 += w * u0;
        } else {
          const double wq = aT / 3;
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            const int m2 = m == 2 ? 0 : m + 1;
            const double x = 0.5 * (T[2 * m] + T[2 * m2]), y = 0.5 * (T[2 * m + 1] + T[2 * m2 + 1]);
            double xo, eo, xn, en;
            aff2_local(gO, x, y, xo, eo);
            aff2_local(gK, x, y, xn, en);
            const double uval = u0 * (1.0 - xo - eo) + u1 * xo + u2 * eo;
            const double f = wq * uval;
            acc0 += f * (1.0 - xn - en); acc1 += f * xn; acc2 += f * en;
          }
        }
        initialize = false;
      }
    }
    if (NCOMP == 1) {
      dofs_new[K] = acc0;
      mass = volK * acc0;
    } else {
      const double S = acc0 + acc1 + acc2;
      const double f = 3.0 / volK;
      const double r0 = f * (4.0 * acc0 - S), r1 = f * (4.0 * acc1 - S), r2 = f * (4.0 * acc2 - S);
      dofs_new[3 * (int64_t)K] = r0; dofs_new[3 * (int64_t)K + 1] = r1; dofs_new[3 * (int64_t)K + 2] = r2;
      mass = volK * ((r0 + r1 + r2) / 3.0);
    }
  }
  // deterministic two-level sums: fixed-shape shuffle tree per block, per-block partials reduced by k_reduce_partials
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mass += __shfl_xor_sync(0xffffffffu, mass, o); cov += __shfl_xor_sync(0xffffffffu, cov, o); }
  npc = warp_sum(npc);
  __shared__ double sm[4], sc[4]; __shared__ unsigned sn[4];
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = mass; sc[threadIdx.x >> 5] = cov; sn[threadIdx.x >> 5] = npc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    part_mass[blockIdx.x] = (sm[0] + sm[1]) + (sm[2] + sm[3]);
    part_cov[blockIdx.x] = (sc[0] + sc[1]) + (sc[2] + sc[3]);
    atomicAdd(&cnt->n_pieces, (unsigned long long)(sn[0] + sn[1] + sn[2] + sn[3]));
  }
}
// pairwise (fixed shape) reduction of the per-block partial sums; one block
__global__ void __launch_bounds__(1024)
k_reduce_partials(const double* __restrict__ a, const double* __restrict__ b, int n, Counters* __restrict__ cnt) {
  __shared__ double sa[1024], sb[1024];
  double xa = 0.0, xb = 0.0;
  // each thread owns a contiguous chunk (fixed for a given n) and sums it left to right
  const int chunk = (n + 1023) / 1024;
  const int lo = threadIdx.x * chunk, hi = min(n, lo + chunk);
  for (int i = lo; i < hi; ++i) { xa += a[i]; xb += b[i]; }
  sa[threadIdx.x] = xa; sb[threadIdx.x] = xb;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { cnt->mass_new = sa[0]; cnt->covered_area = sb[0]; }
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
    Poly poly;
    sh_clip(cx, cy, ex, ey, 1e-14, poly);
    double area = 0.0;
    if (poly.n >= 3) {
      for (int a = 0; a < poly.n; ++a) {
        const int b = a + 1 == poly.n ? 0 : a + 1;
        area += poly.x[a] * poly.y[b];
        area -= poly.x[b] * poly.y[a];
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
                 const int32_t* __restrict__ wl, int8_t* __restrict__ orient, Counters* __restrict__ cnt) {
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
// 3-D facet leaves: 7 sample points (3 corners, 3 edge midpoints, centre), distance.hh:157-173
struct TriLeaf { double pt[7][3]; };
__global__ void __launch_bounds__(256)
k_tri_prepare(const double* __restrict__ x, const int32_t* __restrict__ tris_sorted, int ns, int L,
              TriLeaf* __restrict__ leaf, double* __restrict__ box /* [2L][6] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  double b[6] = { INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY };
  if (i < ns) {
    const double* a = x + 3 * (int64_t)tris_sorted[3 * i];
    const double* bb = x + 3 * (int64_t)tris_sorted[3 * i + 1];
    const double* cc = x + 3 * (int64_t)tris_sorted[3 * i + 2];
    double c[3][3], ctr[3];
    const double third = 1.0 / 3.0;
    for (int k = 0; k < 3; ++k) {
      const double j0 = bb[k] - a[k], j1 = cc[k] - a[k];
      c[0][k] = a[k];
      c[1][k] = (a[k] + 1.0 * j0) + 0.0 * j1;
      c[2][k] = (a[k] + 0.0 * j0) + 1.0 * j1;
      ctr[k] = (a[k] + third * j0) + third * j1;
    }
    TriLeaf f;
    for (int t = 0; t < 3; ++t) {
      const int t2 = (t + 1) % 3;
      for (int k = 0; k < 3; ++k) { f.pt[2 * t][k] = c[t][k]; f.pt[2 * t + 1][k] = 0.5 * (c[t][k] + c[t2][k]); }
    }
    for (int k = 0; k < 3; ++k) f.pt[6][k] = ctr[k];
    leaf[i] = f;
    for (int t = 0; t < 7; ++t) for (int k = 0; k < 3; ++k) { b[k] = fmin(b[k], f.pt[t][k]); b[3 + k] = fmax(b[3 + k], f.pt[t][k]); }
  }
  for (int k = 0; k < 6; ++k) box[6 * (int64_t)(L + i) + k] = b[k];
}
__global__ void __launch_bounds__(1024) k_bvh_upper3(int L, double* __restrict__ box) {
  for (int lvl = L >> 1; lvl >= 1; lvl >>= 1) {
    for (int i = lvl + threadIdx.x; i < 2 * lvl; i += blockDim.x) {
      const double* a = box + 6 * (int64_t)(2 * i); const double* b = a + 6;
      for (int k = 0; k < 3; ++k) { box[6 * (int64_t)i + k] = fmin(a[k], b[k]); box[6 * (int64_t)i + 3 + k] = fmax(a[3 + k], b[3 + k]); }
    }
    __syncthreads();
  }
}
__device__ __forceinline__ double box_dist2_3(const double* __restrict__ b, const double* p) {
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) { const double d = fmax(fmax(__ldg(b + k) - p[k], p[k] - __ldg(b + 3 + k)), 0.0); s += d * d; }
  return s;
}
__global__ void __launch_bounds__(256)
k_distance3d(const double* __restrict__ x, int64_t nv, const TriLeaf* __restrict__ leaf, const double* __restrict__ box,
             int ns, int L, double* __restrict__ dist, Counters* __restrict__ cnt) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double best = 1e100;
  if (v < nv && ns > 0) {
    double p[3]; load3(x, nullptr, (int)v, false, p);
    // the reference takes sqrt(min of squared norms) per facet, then the min over facets: sqrt is monotone, so
    // minimise the squared value over all facets and take one sqrt -- identical bits.
    double best2 = 1e100;       // distance.hh:160 initial value (squared quantities are compared against it too)
    double T2 = best2;
    int stack[40]; double sd2[40]; int sp = 0; int node = 1;
    for (;;) {
      if (node >= L) {
        const int s = node - L;
        if (s < ns) {
          const TriLeaf& f = leaf[s];
#pragma unroll
          for (int t = 0; t < 7; ++t) {
            double n2 = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) { const double d = p[k] - f.pt[t][k]; n2 += d * d; }
            if (n2 < best2) best2 = n2;
          }
          T2 = best2 * (1.0 + 1e-12);
        }
        node = 0;
      } else {
        const double dl = box_dist2_3(box + 6 * (int64_t)(2 * node), p), dr = box_dist2_3(box + 6 * (int64_t)(2 * node + 1), p);
        const bool lf = dl <= dr;
        const double dn = lf ? dl : dr, df = lf ? dr : dl;
        const int nn = 2 * node + (lf ? 0 : 1), nf = 2 * node + (lf ? 1 : 0);
        if (dn > T2) node = 0;
        else { if (!(df > T2)) { stack[sp] = nf; sd2[sp] = df; ++sp; } node = nn; }
      }
      if (node == 0) {
        while (sp > 0) { --sp; if (!(sd2[sp] > T2)) { node = stack[sp]; break; } }
        if (node == 0) break;
      }
    }
    best = sqrt(best2);
    if (best > 1e100) best = 1e100;
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

}  // namespace mmb
