// predicates.cuh -- filtered exact geometric predicates for sm_100a (and the host, for CPU unit tests).
//
// Chain = the one CGAL::Epick uses (CGAL/Filtered_predicate.h:86-106 with the Static_filters layer):
//   1. semi-static filter, constants verbatim from CGAL/internal/Static_filters/{Orientation_2,Orientation_3,
//      Side_of_oriented_circle_2}.h                                               (in the streaming kernels)
//   2. interval arithmetic on the same formula, directed rounding (__dadd_rd/__dmul_ru ...)   (fallback kernel)
//   3. exact evaluation with Shewchuk-style floating-point expansions (grow / scale with zero elimination)
// The result is the true sign of the determinant, i.e. bit-exact to CGAL's exact fallback by construction.
//
// Everything is written against plain IEEE double ops; the translation unit must be compiled with
// -fmad=false (device) and -ffp-contract=off (host) so that no product-sum is contracted.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MMB_HD __host__ __device__ __forceinline__
#define MMB_HDN __host__ __device__ __noinline__
#else
#define MMB_HD inline
#define MMB_HDN
#endif

namespace mmb {

// ------------------------------------------------------------------------------------------------------
// inexact determinants, op order of CGAL/determinant.h:25-52
// ------------------------------------------------------------------------------------------------------
MMB_HD double det2(double a00, double a01, double a10, double a11) { return a00 * a11 - a10 * a01; }
MMB_HD double det3(double a00, double a01, double a02, double a10, double a11, double a12, double a20, double a21,
                   double a22) {
  const double m01 = a00 * a11 - a10 * a01;
  const double m02 = a00 * a21 - a20 * a01;
  const double m12 = a10 * a21 - a20 * a11;
  return m01 * a22 - m02 * a12 + m12 * a02;
}

constexpr int SIGN_UNKNOWN = 2;
constexpr int SIGN_DOMAIN = 3;     // exact stage: the input's exponent range exceeds what double expansions can hold (see pred_rescale)

// CGAL/internal/Static_filters/Orientation_2.h:54-91.  Also returns the inexact determinant (== 2*area of
// CGAL/Cartesian/function_objects.h:918-926, same operations) so validity and the filter share the arithmetic.
MMB_HD int orient2d_filter(double px, double py, double qx, double qy, double rx, double ry, double& det) {
  const double pqx = qx - px, pqy = qy - py, prx = rx - px, pry = ry - py;
  det = det2(pqx, pqy, prx, pry);
  double maxx = fabs(pqx), maxy = fabs(pqy);
  const double aprx = fabs(prx), apry = fabs(pry);
  if (maxx < aprx) maxx = aprx;
  if (maxy < apry) maxy = apry;
  if (maxx > maxy) { const double t = maxx; maxx = maxy; maxy = t; }
  if (maxx < 1e-146) {
    if (maxx == 0) return 0;
  } else if (maxy < 1e153) {
    const double eps = 8.8872057372592798e-16 * maxx * maxy;
    if (det > eps) return 1;
    if (det < -eps) return -1;
  }
  return SIGN_UNKNOWN;
}

// CGAL/internal/Static_filters/Orientation_3.h:60-138
MMB_HD int orient3d_filter(const double* p, const double* q, const double* r, const double* s, double& det) {
  const double pqx = q[0] - p[0], pqy = q[1] - p[1], pqz = q[2] - p[2];
  const double prx = r[0] - p[0], pry = r[1] - p[1], prz = r[2] - p[2];
  const double psx = s[0] - p[0], psy = s[1] - p[1], psz = s[2] - p[2];
  double maxx = fabs(pqx), maxy = fabs(pqy), maxz = fabs(pqz);
  const double aprx = fabs(prx), apsx = fabs(psx), apry = fabs(pry), apsy = fabs(psy), aprz = fabs(prz),
               apsz = fabs(psz);
  if (maxx < aprx) maxx = aprx;
  if (maxx < apsx) maxx = apsx;
  if (maxy < apry) maxy = apry;
  if (maxy < apsy) maxy = apsy;
  if (maxz < aprz) maxz = aprz;
  if (maxz < apsz) maxz = apsz;
  det = det3(pqx, pqy, pqz, prx, pry, prz, psx, psy, psz);
  const double eps = 5.1107127829973299e-15 * maxx * maxy * maxz;
  double t;
  if (maxx > maxz) { t = maxx; maxx = maxz; maxz = t; }
  if (maxy > maxz) { t = maxy; maxy = maxz; maxz = t; }
  else if (maxy < maxx) { t = maxx; maxx = maxy; maxy = t; }
  if (maxx < 1e-97) {
    if (maxx == 0) return 0;
  } else if (maxz < 1e102) {
    if (det > eps) return 1;
    if (det < -eps) return -1;
  }
  return SIGN_UNKNOWN;
}

// CGAL/internal/Static_filters/Side_of_oriented_circle_2.h:45-102
MMB_HD int incircle_filter(double px, double py, double qx, double qy, double rx, double ry, double tx, double ty) {
  const double qpx = qx - px, qpy = qy - py, rpx = rx - px, rpy = ry - py;
  const double tpx = tx - px, tpy = ty - py, tqx = tx - qx, tqy = ty - qy;
  const double rqx = rx - qx, rqy = ry - qy;
  const double det = det2(qpx * tpy - qpy * tpx, tpx * tqx + tpy * tqy, qpx * rpy - qpy * rpx, rpx * rqx + rpy * rqy);
  double maxx = fabs(qpx), maxy = fabs(qpy);
  const double arpx = fabs(rpx), arpy = fabs(rpy), atqx = fabs(tqx), atqy = fabs(tqy);
  const double atpx = fabs(tpx), atpy = fabs(tpy), arqx = fabs(rqx), arqy = fabs(rqy);
  if (maxx < arpx) maxx = arpx;
  if (maxx < atpx) maxx = atpx;
  if (maxx < atqx) maxx = atqx;
  if (maxx < arqx) maxx = arqx;
  if (maxy < arpy) maxy = arpy;
  if (maxy < atpy) maxy = atpy;
  if (maxy < atqy) maxy = atqy;
  if (maxy < arqy) maxy = arqy;
  if (maxx > maxy) { const double t = maxx; maxx = maxy; maxy = t; }
  if (maxx < 1e-73) {
    if (maxx == 0) return 0;
  } else if (maxy < 1e76) {
    const double eps = 8.8878565762001373e-15 * maxx * maxy * (maxy * maxy);
    if (det > eps) return 1;
    if (det < -eps) return -1;
  }
  return SIGN_UNKNOWN;
}

// ------------------------------------------------------------------------------------------------------
// stage 2: interval arithmetic with directed rounding
// ------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
MMB_HD double add_rd(double a, double b) { return __dadd_rd(a, b); }
MMB_HD double add_ru(double a, double b) { return __dadd_ru(a, b); }
MMB_HD double mul_rd(double a, double b) { return __dmul_rd(a, b); }
MMB_HD double mul_ru(double a, double b) { return __dmul_ru(a, b); }
#else
// host build (unit tests only): widen the round-to-nearest result by one ulp -- still a valid enclosure
MMB_HD double add_rd(double a, double b) { return nextafter(a + b, -INFINITY); }
MMB_HD double add_ru(double a, double b) { return nextafter(a + b, INFINITY); }
MMB_HD double mul_rd(double a, double b) { return nextafter(a * b, -INFINITY); }
MMB_HD double mul_ru(double a, double b) { return nextafter(a * b, INFINITY); }
#endif

struct Itv { double lo, hi; };
MMB_HD Itv iv_diff(double a, double b) { return Itv{ add_rd(a, -b), add_ru(a, -b) }; }
MMB_HD Itv iv_add(Itv a, Itv b) { return Itv{ add_rd(a.lo, b.lo), add_ru(a.hi, b.hi) }; }
MMB_HD Itv iv_sub(Itv a, Itv b) { return Itv{ add_rd(a.lo, -b.hi), add_ru(a.hi, -b.lo) }; }
MMB_HD Itv iv_mul(Itv a, Itv b) {
  const double l1 = mul_rd(a.lo, b.lo), l2 = mul_rd(a.lo, b.hi), l3 = mul_rd(a.hi, b.lo), l4 = mul_rd(a.hi, b.hi);
  const double h1 = mul_ru(a.lo, b.lo), h2 = mul_ru(a.lo, b.hi), h3 = mul_ru(a.hi, b.lo), h4 = mul_ru(a.hi, b.hi);
  return Itv{ fmin(fmin(l1, l2), fmin(l3, l4)), fmax(fmax(h1, h2), fmax(h3, h4)) };
}
MMB_HD int iv_sign(Itv a) {
  if (a.lo > 0) return 1;
  if (a.hi < 0) return -1;
  if (a.lo == 0 && a.hi == 0) return 0;
  return SIGN_UNKNOWN;  // also taken for NaN / inf enclosures
}
MMB_HD int orient2d_interval(double px, double py, double qx, double qy, double rx, double ry) {
  const Itv a = iv_diff(qx, px), b = iv_diff(qy, py), c = iv_diff(rx, px), d = iv_diff(ry, py);
  return iv_sign(iv_sub(iv_mul(a, d), iv_mul(c, b)));
}
MMB_HD int incircle_interval(double px, double py, double qx, double qy, double rx, double ry, double tx, double ty) {
  const Itv qpx = iv_diff(qx, px), qpy = iv_diff(qy, py), rpx = iv_diff(rx, px), rpy = iv_diff(ry, py);
  const Itv tpx = iv_diff(tx, px), tpy = iv_diff(ty, py), tqx = iv_diff(tx, qx), tqy = iv_diff(ty, qy);
  const Itv rqx = iv_diff(rx, qx), rqy = iv_diff(ry, qy);
  const Itv a00 = iv_sub(iv_mul(qpx, tpy), iv_mul(qpy, tpx));
  const Itv a01 = iv_add(iv_mul(tpx, tqx), iv_mul(tpy, tqy));
  const Itv a10 = iv_sub(iv_mul(qpx, rpy), iv_mul(qpy, rpx));
  const Itv a11 = iv_add(iv_mul(rpx, rqx), iv_mul(rpy, rqy));
  return iv_sign(iv_sub(iv_mul(a00, a11), iv_mul(a10, a01)));
}
MMB_HD int orient3d_interval(const double* p, const double* q, const double* r, const double* s) {
  Itv a[3][3];
  for (int j = 0; j < 3; ++j) { a[0][j] = iv_diff(q[j], p[j]); a[1][j] = iv_diff(r[j], p[j]); a[2][j] = iv_diff(s[j], p[j]); }
  const Itv m01 = iv_sub(iv_mul(a[0][0], a[1][1]), iv_mul(a[1][0], a[0][1]));
  const Itv m02 = iv_sub(iv_mul(a[0][0], a[2][1]), iv_mul(a[2][0], a[0][1]));
  const Itv m12 = iv_sub(iv_mul(a[1][0], a[2][1]), iv_mul(a[2][0], a[1][1]));
  return iv_sign(iv_add(iv_sub(iv_mul(m01, a[2][2]), iv_mul(m02, a[1][2])), iv_mul(m12, a[0][2])));
}

// ------------------------------------------------------------------------------------------------------
// stage 3: exact sign with floating-point expansions (non-overlapping, increasing magnitude, zero-eliminated)
// ------------------------------------------------------------------------------------------------------
MMB_HD double fma_exact(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}
MMB_HD void two_sum(double a, double b, double& x, double& y) {
  x = a + b;
  const double bv = x - a;
  const double av = x - bv;
  const double br = b - bv;
  const double ar = a - av;
  y = ar + br;
}
MMB_HD void two_prod(double a, double b, double& x, double& y) {
  x = a * b;
  y = fma_exact(a, b, -x);
}
// e <- e + b, in place; returns the new length
MMB_HD int ex_grow(int n, double* e, double b) {
  double Q = b, Qn, h;
  int k = 0;
  for (int i = 0; i < n; ++i) {
    two_sum(Q, e[i], Qn, h);
    Q = Qn;
    if (h != 0.0) e[k++] = h;
  }
  if (Q != 0.0 || k == 0) e[k++] = Q;
  return k;
}
// e = a - b exactly (1 or 2 components)
MMB_HD int ex_diff(double a, double b, double* e) {
  double x, y;
  two_sum(a, -b, x, y);
  int n = 0;
  if (y != 0.0) e[n++] = y;
  if (x != 0.0 || n == 0) e[n++] = x;
  return n;
}
// h = e * b (h has room for 2n doubles); returns length
MMB_HD int ex_scale(int n, const double* e, double b, double* h) {
  double Q, hh, p1, p0, sum;
  int k = 0;
  two_prod(e[0], b, Q, hh);
  if (hh != 0.0) h[k++] = hh;
  for (int i = 1; i < n; ++i) {
    two_prod(e[i], b, p1, p0);
    two_sum(Q, p0, sum, hh);
    if (hh != 0.0) h[k++] = hh;
    two_sum(p1, sum, Q, hh);
    if (hh != 0.0) h[k++] = hh;
  }
  if (Q != 0.0 || k == 0) h[k++] = Q;
  return k;
}
// acc <- acc + sgn * (a * b); tmp has room for 2*na doubles; returns new length of acc
MMB_HD int ex_muladd(int nacc, double* acc, int na, const double* a, int nb, const double* b, double sgn, double* tmp) {
  for (int j = 0; j < nb; ++j) {
    const int nt = ex_scale(na, a, sgn * b[j], tmp);
    for (int i = 0; i < nt; ++i) nacc = ex_grow(nacc, acc, tmp[i]);
  }
  return nacc;
}
MMB_HD int ex_sign(int n, const double* e) {
  const double top = e[n - 1];
  return (top > 0.0) - (top < 0.0);
}

// Exponent range of the exact stage.  CGAL's exact fallback computes with rationals and has no range limit; expansions of
// doubles are exact only while no product over- or underflows.  The predicates are homogeneous polynomials of degree `deg`
// in the coordinates, so multiplying every coordinate by one power of two (exact) leaves the sign alone: the inputs are
// rescaled so that the largest magnitude lies in [1/2, 1) -- no overflow for any absolute scale.  Every intermediate
// value is then an integer multiple of beta^k, beta = (smallest nonzero |coordinate|) * 2^-52 (its last bit), k <= deg;
// all of them, and the error terms of two_prod, are representable iff beta^deg >= 2^-1074.  That bounds the admissible
// RATIO largest / smallest nonzero magnitude inside one predicate: 2^480 (orient2d), 2^300 (orient3d), 2^212 (in-circle).
// Inputs beyond it, or non-finite ones, yield SIGN_DOMAIN: the caller counts them and the C-ABI reports
// MMB_ERR_CAPACITY instead of returning a sign that might be wrong.
MMB_HD bool pred_rescale(double* c, int n, int max_ratio_log2) {
  double M = 0.0, m = INFINITY;
  for (int i = 0; i < n; ++i) {
    const double a = fabs(c[i]);
    if (!(a <= 1.79769313486231570e308)) return false;      // inf / NaN
    if (a > 0.0) { M = a > M ? a : M; m = a < m ? a : m; }
  }
  if (M == 0.0) return true;
  int eM, em;
  frexp(M, &eM); frexp(m, &em);
  if (eM - em > max_ratio_log2) return false;
  for (int i = 0; i < n; ++i) c[i] = ldexp(c[i], -eM);      // exact: the smallest magnitude stays far above the denormals
  return true;
}

// exact orientationC2 (CGAL/predicates/kernel_ftC2.h:401-406)
MMB_HDN int orient2d_exact(double px, double py, double qx, double qy, double rx, double ry) {
  {
    double c[6] = { px, py, qx, qy, rx, ry };
    if (!pred_rescale(c, 6, 480)) return SIGN_DOMAIN;
    px = c[0]; py = c[1]; qx = c[2]; qy = c[3]; rx = c[4]; ry = c[5];
  }
  double a[2], b[2], c[2], d[2], tmp[4], acc[16];
  const int na = ex_diff(qx, px, a), nb = ex_diff(qy, py, b), nc = ex_diff(rx, px, c), nd = ex_diff(ry, py, d);
  int n = 1; acc[0] = 0.0;
  n = ex_muladd(n, acc, na, a, nd, d, 1.0, tmp);
  n = ex_muladd(n, acc, nc, c, nb, b, -1.0, tmp);
  return ex_sign(n, acc);
}

// exact side_of_oriented_circleC2 (CGAL/predicates/kernel_ftC2.h:477-499, 2x2 reduced form)
MMB_HDN int incircle_exact(double px, double py, double qx, double qy, double rx, double ry, double tx, double ty) {
  {
    double c[8] = { px, py, qx, qy, rx, ry, tx, ty };
    if (!pred_rescale(c, 8, 212)) return SIGN_DOMAIN;
    px = c[0]; py = c[1]; qx = c[2]; qy = c[3]; rx = c[4]; ry = c[5]; tx = c[6]; ty = c[7];
  }
  double qpx[2], qpy[2], rpx[2], rpy[2], tpx[2], tpy[2], tqx[2], tqy[2], rqx[2], rqy[2];
  const int nqpx = ex_diff(qx, px, qpx), nqpy = ex_diff(qy, py, qpy), nrpx = ex_diff(rx, px, rpx),
            nrpy = ex_diff(ry, py, rpy), ntpx = ex_diff(tx, px, tpx), ntpy = ex_diff(ty, py, tpy),
            ntqx = ex_diff(tx, qx, tqx), ntqy = ex_diff(ty, qy, tqy), nrqx = ex_diff(rx, qx, rqx),
            nrqy = ex_diff(ry, qy, rqy);
  double A[16], B[16], C[16], D[16], tmp[32];
  int nA = 1, nB = 1, nC = 1, nD = 1;
  A[0] = B[0] = C[0] = D[0] = 0.0;
  nA = ex_muladd(nA, A, nqpx, qpx, ntpy, tpy, 1.0, tmp);   // a00 = qpx*tpy - qpy*tpx
  nA = ex_muladd(nA, A, nqpy, qpy, ntpx, tpx, -1.0, tmp);
  nB = ex_muladd(nB, B, ntpx, tpx, ntqx, tqx, 1.0, tmp);   // a01 = tpx*tqx + tpy*tqy
  nB = ex_muladd(nB, B, ntpy, tpy, ntqy, tqy, 1.0, tmp);
  nC = ex_muladd(nC, C, nqpx, qpx, nrpy, rpy, 1.0, tmp);   // a10 = qpx*rpy - qpy*rpx
  nC = ex_muladd(nC, C, nqpy, qpy, nrpx, rpx, -1.0, tmp);
  nD = ex_muladd(nD, D, nrpx, rpx, nrqx, rqx, 1.0, tmp);   // a11 = rpx*rqx + rpy*rqy
  nD = ex_muladd(nD, D, nrpy, rpy, nrqy, rqy, 1.0, tmp);
  double acc[1024];                                        // 2 * 16 * 16 * 2 components, worst case
  int n = 1; acc[0] = 0.0;
  n = ex_muladd(n, acc, nA, A, nD, D, 1.0, tmp);           // a00*a11 - a10*a01
  n = ex_muladd(n, acc, nC, C, nB, B, -1.0, tmp);
  return ex_sign(n, acc);
}

// exact orientationC3 (CGAL/predicates/kernel_ftC3.h:114-122; determinant.h:37-52)
MMB_HDN int orient3d_exact(const double* p_in, const double* q_in, const double* r_in, const double* s_in) {
  double c12[12];
  for (int j = 0; j < 3; ++j) { c12[j] = p_in[j]; c12[3 + j] = q_in[j]; c12[6 + j] = r_in[j]; c12[9 + j] = s_in[j]; }
  if (!pred_rescale(c12, 12, 300)) return SIGN_DOMAIN;
  const double *p = c12, *q = c12 + 3, *r = c12 + 6, *s = c12 + 9;
  double a[3][3][2]; int na[3][3];
  for (int j = 0; j < 3; ++j) {
    na[0][j] = ex_diff(q[j], p[j], a[0][j]);
    na[1][j] = ex_diff(r[j], p[j], a[1][j]);
    na[2][j] = ex_diff(s[j], p[j], a[2][j]);
  }
  double m01[16], m02[16], m12[16], tmp[32];
  int n01 = 1, n02 = 1, n12 = 1;
  m01[0] = m02[0] = m12[0] = 0.0;
  n01 = ex_muladd(n01, m01, na[0][0], a[0][0], na[1][1], a[1][1], 1.0, tmp);
  n01 = ex_muladd(n01, m01, na[1][0], a[1][0], na[0][1], a[0][1], -1.0, tmp);
  n02 = ex_muladd(n02, m02, na[0][0], a[0][0], na[2][1], a[2][1], 1.0, tmp);
  n02 = ex_muladd(n02, m02, na[2][0], a[2][0], na[0][1], a[0][1], -1.0, tmp);
  n12 = ex_muladd(n12, m12, na[1][0], a[1][0], na[2][1], a[2][1], 1.0, tmp);
  n12 = ex_muladd(n12, m12, na[2][0], a[2][0], na[1][1], a[1][1], -1.0, tmp);
  double acc[192];
  int n = 1; acc[0] = 0.0;
  n = ex_muladd(n, acc, n01, m01, na[2][2], a[2][2], 1.0, tmp);
  n = ex_muladd(n, acc, n02, m02, na[1][2], a[1][2], -1.0, tmp);
  n = ex_muladd(n, acc, n12, m12, na[0][2], a[0][2], 1.0, tmp);
  return ex_sign(n, acc);
}

// full chains for the fallback kernels (stage 1 already failed)
MMB_HD int orient2d_robust(double px, double py, double qx, double qy, double rx, double ry) {
  const int s = orient2d_interval(px, py, qx, qy, rx, ry);
  return s != SIGN_UNKNOWN ? s : orient2d_exact(px, py, qx, qy, rx, ry);
}
MMB_HD int incircle_robust(double px, double py, double qx, double qy, double rx, double ry, double tx, double ty) {
  const int s = incircle_interval(px, py, qx, qy, rx, ry, tx, ty);
  return s != SIGN_UNKNOWN ? s : incircle_exact(px, py, qx, qy, rx, ry, tx, ty);
}
MMB_HD int orient3d_robust(const double* p, const double* q, const double* r, const double* s4) {
  const int s = orient3d_interval(p, q, r, s4);
  return s != SIGN_UNKNOWN ? s : orient3d_exact(p, q, r, s4);
}

}  // namespace mmb
