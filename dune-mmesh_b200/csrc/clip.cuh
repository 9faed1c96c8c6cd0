// clip.cuh -- one clip edge of the reference's Sutherland-Hodgman (dune/mmesh/grid/polygoncutting.hh:17-114) in the form
// the projection kernel uses it.  Host-compilable (tests/native/host_clip.cc runs it against the oracle without a GPU);
// compile with -fmad=false (device) / -ffp-contract=off (host): the arithmetic must not be contracted.
#pragma once
#include <math.h>
#include "predicates.cuh"      // MMB_HD

#if !defined(__CUDACC__)
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

namespace mmb {

MMB_HD int mmb_ffs(unsigned v) {
#if defined(__CUDA_ARCH__)
  return (int)__ffs((int)v);
#else
  return __builtin_ffs((int)v);
#endif
}

// ---- shared-memory Sutherland-Hodgman, organised in convergent passes ------------------------------------------
// lineIntersectionPoint, polygoncutting.hh:100-114 (iQ = q1 - q2, fq = q1 x q2 hoisted per clip edge)
MMB_HD double2 line_isect(double ax, double ay, double bx, double by, double iQx, double iQy, double fq) {
  double dPx = ax - bx, dPy = ay - by;
  const double scale = dPx * iQy - iQx * dPy;
  const double fp = ax * by - ay * bx;
  dPx *= fq; dPy *= fq;
  return make_double2((iQx * fp - dPx) / scale, (iQy * fp - dPy) / scale);
}
// One clip edge (q1 -> q2) of polygoncutting.hh:17-96 applied to the polygon load(0..n-1), n <= NMAX; the result
// goes to store(0..m-1), m <= n + n/2 is returned.  Same case analysis, operand order and IEEE operations as the
// reference, but split into three passes so that the lanes of a warp stay together:
//   1. side value of every vertex, reduced to the three comparison bits the case analysis needs,
//   2. case masks of the edges by bit operations, the crossing points (a convex polygon has at most one edge of each
//      crossing kind),
//   3. emission in edge order, fully unrolled and predicated (pure data movement); several crossings of one kind -- only
//      possible through the epsilon band -- take a generic loop that computes the further ones in place.
// The clip edge enters through three numbers: dQ = q2 - q1 and q2xq1 = q1y*q2x - q1x*q2y (clip_edge_constants).  The other
// quantities of the reference are exact negations of these (IEEE subtraction is anti-commutative, the products are
// rounded separately): q1 - q2 = -dQ and q1x*q2y - q1y*q2x = -q2xq1, so they need not be stored per clip edge.
// CAP = capacity of the output polygon: a result with more than CAP points writes nothing and returns -1 (the caller
// sends that pair through the generic path); CAP >= NMAX + NMAX/2 can never trigger.
template <int NMAX, int CAP, class Load, class Store>
MMB_HD int clip_stage_core(Load load, int n, double dQx, double dQy, double q2xq1, double eps, Store store) {
  if (n <= 0) return 0;
  const double ndQx = -dQx;
  const double iQx = -dQx, iQy = -dQy;
  const double fq = -q2xq1;
  // pass 1: side value of every vertex, kept only as three comparison bits (f > eps, f < -eps, f >= -eps: the three
  // tests of polygoncutting.hh:63-89; NaN makes all of them false exactly as in the reference's if-chain)
  unsigned Pm = 0, Nm = 0, Gm = 0;
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    if (i < n) {
      double x, y; load(i, x, y);
      const double f = ndQx * y + dQy * x + q2xq1;
      Pm |= (f > eps ? 1u : 0u) << i; Nm |= (f < -eps ? 1u : 0u) << i; Gm |= (f >= -eps ? 1u : 0u) << i;
    }
  }
  // pass 2: case of every edge (i, i+1 mod n) from the bits of its end points, with the priority of the reference's
  // if / else-if chain (the cases are disjoint anyway for eps >= 0):
  //   c1: first > eps  && second < -eps     c2: first < -eps && second > eps     c3: first >= -eps && second > eps
  const unsigned Pj = (Pm >> 1) | ((Pm & 1u) << (n - 1)), Nj = (Nm >> 1) | ((Nm & 1u) << (n - 1));
  const unsigned c1 = Pm & Nj, c2 = Nm & Pj & ~c1, c3 = Gm & Pj & ~(c1 | c2);
  const int e1 = mmb_ffs(c1) - 1, e2 = mmb_ffs(c2) - 1;      // a convex polygon has at most one edge of each kind
  double2 X1 = make_double2(0.0, 0.0), X2 = X1;
  if (e1 >= 0) { double ax, ay, bx, by; load(e1, ax, ay); load(e1 + 1 == n ? 0 : e1 + 1, bx, by); X1 = line_isect(ax, ay, bx, by, iQx, iQy, fq); }
  if (e2 >= 0) { double ax, ay, bx, by; load(e2, ax, ay); load(e2 + 1 == n ? 0 : e2 + 1, bx, by); X2 = line_isect(ax, ay, bx, by, iQx, iQy, fq); }
  // pass 3: emission in edge order (pure data movement): c1 -> X, second; c2 -> X; c3 -> nothing; otherwise -> second.
  const unsigned keep = ~(c2 | c3);
  if (CAP < NMAX + NMAX / 2) {
    const unsigned live = (1u << n) - 1u;
#if defined(__CUDA_ARCH__)
    const int total = __popc(c1 & live) + __popc(c2 & live) + __popc(keep & live);
#else
    const int total = __builtin_popcount(c1 & live) + __builtin_popcount(c2 & live) + __builtin_popcount(keep & live);
#endif
    if (total > CAP) return -1;
  }
  int m = 0;
  if (((c1 & (c1 - 1u)) | (c2 & (c2 - 1u))) == 0u) {
    // at most one crossing of each kind (always, for a convex polygon outside the epsilon band): unrolled, predicated
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
      if (i < n) {
        const unsigned bit = 1u << i;
        if (c1 & bit) { store(m, X1.x, X1.y); ++m; }
        if (c2 & bit) { store(m, X2.x, X2.y); ++m; }
        if (keep & bit) { double bx, by; load(i + 1 == n ? 0 : i + 1, bx, by); store(m, bx, by); ++m; }
      }
    }
  } else {
    // several crossings of one kind (only through the epsilon band / degenerate input): the others are computed in place
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      const unsigned bit = 1u << i;
      const int j = i + 1 == n ? 0 : i + 1;
      if ((c1 | c2) & bit) {
        const bool k1 = (c1 & bit) != 0u;
        double2 X = k1 ? X1 : X2;
        if (i != (k1 ? e1 : e2)) { double ax, ay, bx, by; load(i, ax, ay); load(j, bx, by); X = line_isect(ax, ay, bx, by, iQx, iQy, fq); }
        store(m, X.x, X.y); ++m;
      }
      if (keep & bit) { double bx, by; load(j, bx, by); store(m, bx, by); ++m; }
    }
  }
  return m;
}

MMB_HD void clip_edge_constants(double q1x, double q1y, double q2x, double q2y, double& dQx, double& dQy, double& q2xq1) {
  dQx = q2x - q1x; dQy = q2y - q1y;
  q2xq1 = q1y * q2x - q1x * q2y;
}
// the same stage from the two end points of the clip edge (polygoncutting.hh:26-29)
template <int NMAX, class Load, class Store>
MMB_HD int clip_stage(Load load, int n, double q1x, double q1y, double q2x, double q2y, double eps, Store store) {
  double dQx, dQy, q2xq1;
  clip_edge_constants(q1x, q1y, q2x, q2y, dQx, dQy, q2xq1);
  return clip_stage_core<NMAX, NMAX + NMAX / 2>(load, n, dQx, dQy, q2xq1, eps, store);
}

}  // namespace mmb
