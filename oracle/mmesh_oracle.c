/*
 * mmesh_oracle.c -- CPU ORACLE (test infrastructure only; see mmesh_oracle.h).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: the reference is a default (non -march=native) x86-64 build, so
 * no FMA contraction anywhere in the inexact arithmetic.  fma() is used only inside two_prod.
 *
 * Reference citations are relative to the reference checkout (dune/..., CGAL/...).
 */
#include "mmesh_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int orc_num_threads(void) { return g_threads; }

/* ======================================================================================= */
/* 1. Exact sign machinery: non-adaptive floating-point expansions (Shewchuk 1997, Thm 10  */
/*    grow-expansion, Thm 19 scale-expansion).  Deliberately simple and always fully exact; */
/*    cross-checked against python fractions.Fraction in tests/test_oracle_predicates.py.   */
/* ======================================================================================= */

#define EXPCAP 4096
typedef struct { int n; double c[EXPCAP]; } expn;

static inline void two_sum(double a, double b, double* x, double* y) {
  double s = a + b;
  double bv = s - a;
  double av = s - bv;
  double br = b - bv;
  double ar = a - av;
  *x = s; *y = ar + br;
}
static inline void two_prod(double a, double b, double* x, double* y) {
  double p = a * b;
  *x = p; *y = fma(a, b, -p);
}
/* h = e + b (zero-eliminating); h may alias e */
static int grow(int elen, const double* e, double b, double* h) {
  double Q = b, Qn, hh; int hi = 0;
  for (int i = 0; i < elen; ++i) {
    two_sum(Q, e[i], &Qn, &hh); Q = Qn;
    if (hh != 0.0) h[hi++] = hh;
  }
  if (Q != 0.0 || hi == 0) h[hi++] = Q;
  return hi;
}
static void ex_set2(expn* e, double a, double b) { /* e = a - b exactly */
  double x, y; two_sum(a, -b, &x, &y);
  e->n = 0;
  if (y != 0.0) e->c[e->n++] = y;
  if (x != 0.0 || e->n == 0) e->c[e->n++] = x;
}
static void ex_add(const expn* e, const expn* f, expn* h) { /* h = e + f ; h distinct from f */
  expn t; t.n = e->n; memcpy(t.c, e->c, sizeof(double) * (size_t)e->n);
  for (int i = 0; i < f->n; ++i) t.n = grow(t.n, t.c, f->c[i], t.c);
  h->n = t.n; memcpy(h->c, t.c, sizeof(double) * (size_t)t.n);
}
static void ex_neg(expn* e) { for (int i = 0; i < e->n; ++i) e->c[i] = -e->c[i]; }
static void ex_scale(const expn* e, double b, expn* h) { /* Shewchuk scale_expansion_zeroelim, Two_Sum everywhere */
  double Q, hh, p1, p0, sum; int hi = 0;
  two_prod(e->c[0], b, &Q, &hh);
  if (hh != 0.0) h->c[hi++] = hh;
  for (int i = 1; i < e->n; ++i) {
    two_prod(e->c[i], b, &p1, &p0);
    two_sum(Q, p0, &sum, &hh);
    if (hh != 0.0) h->c[hi++] = hh;
    two_sum(p1, sum, &Q, &hh);
    if (hh != 0.0) h->c[hi++] = hh;
  }
  if (Q != 0.0 || hi == 0) h->c[hi++] = Q;
  h->n = hi;
}
static void ex_mul(const expn* e, const expn* f, expn* h) { /* h = e * f */
  expn acc, t; acc.n = 1; acc.c[0] = 0.0;
  for (int j = 0; j < f->n; ++j) {
    ex_scale(e, f->c[j], &t);
    for (int i = 0; i < t.n; ++i) acc.n = grow(acc.n, acc.c, t.c[i], acc.c);
  }
  h->n = acc.n; memcpy(h->c, acc.c, sizeof(double) * (size_t)acc.n);
}
static int ex_sign(const expn* e) {
  double top = e->c[e->n - 1];
  return (top > 0.0) - (top < 0.0);
}
/* h = a*d - b*c */
static void ex_det2(const expn* a, const expn* b, const expn* c, const expn* d, expn* h) {
  expn* t1 = (expn*)malloc(sizeof(expn)); expn* t2 = (expn*)malloc(sizeof(expn));
  ex_mul(a, d, t1); ex_mul(b, c, t2); ex_neg(t2); ex_add(t1, t2, h);
  free(t1); free(t2);
}

/* CGAL/predicates/kernel_ftC2.h:401-406 orientationC2 = sign_of_determinant(qx-px, qy-py, rx-px, ry-py),
   evaluated over the reals (Filtered_predicate.h:86-106 falls back to an exact number type). */
/* Exponent range of the expansion arithmetic (the product's predicates.cuh states the same rule): the predicates are
   homogeneous, so all coordinates are multiplied by one power of two (exact) until the largest magnitude is in [1/2, 1);
   beyond a ratio largest / smallest nonzero magnitude of 2^max_ratio_log2, or for non-finite input, the result is 3
   ("outside the exact stage's range") instead of a sign.  CGAL's rational arithmetic has no such limit. */
static int orc_rescale(double* c, int n, int max_ratio_log2) {
  double M = 0.0, m = INFINITY;
  for (int i = 0; i < n; ++i) {
    double a = fabs(c[i]);
    if (!(a <= 1.79769313486231570e308)) return 0;
    if (a > 0.0) { if (a > M) M = a; if (a < m) m = a; }
  }
  if (M == 0.0) return 1;
  int eM, em;
  frexp(M, &eM); frexp(m, &em);
  if (eM - em > max_ratio_log2) return 0;
  for (int i = 0; i < n; ++i) c[i] = ldexp(c[i], -eM);
  return 1;
}
int orc_orient2d(const double* p_, const double* q_, const double* r_) {
  double cc[6] = { p_[0], p_[1], q_[0], q_[1], r_[0], r_[1] };
  if (!orc_rescale(cc, 6, 480)) return 3;
  const double *p = cc, *q = cc + 2, *r = cc + 4;
  expn *a = malloc(sizeof(expn)), *b = malloc(sizeof(expn)), *c = malloc(sizeof(expn)),
       *d = malloc(sizeof(expn)), *h = malloc(sizeof(expn));
  ex_set2(a, q[0], p[0]); ex_set2(b, q[1], p[1]);
  ex_set2(c, r[0], p[0]); ex_set2(d, r[1], p[1]);
  ex_det2(a, b, c, d, h);              /* a00*a11 - a10*a01 (determinant.h:28-35) */
  int s = ex_sign(h);
  free(a); free(b); free(c); free(d); free(h);
  return s;
}

/* CGAL/predicates/kernel_ftC3.h:114-122 orientationC3 (transposed layout, same determinant),
   CGAL/determinant.h:37-52 3x3 formula. */
int orc_orient3d(const double* p_, const double* q_, const double* r_, const double* s_) {
  double cc[12];
  for (int j = 0; j < 3; ++j) { cc[j] = p_[j]; cc[3 + j] = q_[j]; cc[6 + j] = r_[j]; cc[9 + j] = s_[j]; }
  if (!orc_rescale(cc, 12, 300)) return 3;
  const double *p = cc, *q = cc + 3, *r = cc + 6, *s = cc + 9;
  expn* A[3][3]; expn *m01 = malloc(sizeof(expn)), *m02 = malloc(sizeof(expn)), *m12 = malloc(sizeof(expn));
  expn *t = malloc(sizeof(expn)), *u = malloc(sizeof(expn)), *acc = malloc(sizeof(expn));
  const double* rows[3] = { q, r, s };
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    A[i][j] = malloc(sizeof(expn)); ex_set2(A[i][j], rows[i][j], p[j]);
  }
  ex_det2(A[0][0], A[0][1], A[1][0], A[1][1], m01);
  ex_det2(A[0][0], A[0][1], A[2][0], A[2][1], m02);
  ex_det2(A[1][0], A[1][1], A[2][0], A[2][1], m12);
  ex_mul(m01, A[2][2], acc);
  ex_mul(m02, A[1][2], t); ex_neg(t); ex_add(acc, t, u);
  ex_mul(m12, A[0][2], t); ex_add(u, t, acc);
  int sg = ex_sign(acc);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) free(A[i][j]);
  free(m01); free(m02); free(m12); free(t); free(u); free(acc);
  return sg;
}

/* CGAL/predicates/kernel_ftC2.h:477-499 side_of_oriented_circleC2 (2x2 reduced form). */
int orc_incircle(const double* p_, const double* q_, const double* r_, const double* t_) {
  double cc[8] = { p_[0], p_[1], q_[0], q_[1], r_[0], r_[1], t_[0], t_[1] };
  if (!orc_rescale(cc, 8, 212)) return 3;
  const double *p = cc, *q = cc + 2, *r = cc + 4, *t = cc + 6;
  expn* v[10]; for (int i = 0; i < 10; ++i) v[i] = malloc(sizeof(expn));
  expn *a00 = malloc(sizeof(expn)), *a01 = malloc(sizeof(expn)), *a10 = malloc(sizeof(expn)),
       *a11 = malloc(sizeof(expn)), *x = malloc(sizeof(expn)), *y = malloc(sizeof(expn)), *h = malloc(sizeof(expn));
  expn *qpx = v[0], *qpy = v[1], *rpx = v[2], *rpy = v[3], *tpx = v[4], *tpy = v[5],
       *tqx = v[6], *tqy = v[7], *rqx = v[8], *rqy = v[9];
  ex_set2(qpx, q[0], p[0]); ex_set2(qpy, q[1], p[1]);
  ex_set2(rpx, r[0], p[0]); ex_set2(rpy, r[1], p[1]);
  ex_set2(tpx, t[0], p[0]); ex_set2(tpy, t[1], p[1]);
  ex_set2(tqx, t[0], q[0]); ex_set2(tqy, t[1], q[1]);
  ex_set2(rqx, r[0], q[0]); ex_set2(rqy, r[1], q[1]);
  ex_det2(qpx, qpy, tpx, tpy, a00);                       /* qpx*tpy - qpy*tpx */
  ex_mul(tpx, tqx, x); ex_mul(tpy, tqy, y); ex_add(x, y, a01);
  ex_det2(qpx, qpy, rpx, rpy, a10);                       /* qpx*rpy - qpy*rpx */
  ex_mul(rpx, rqx, x); ex_mul(rpy, rqy, y); ex_add(x, y, a11);
  ex_det2(a00, a01, a10, a11, h);                         /* a00*a11 - a10*a01 */
  int s = ex_sign(h);
  for (int i = 0; i < 10; ++i) free(v[i]);
  free(a00); free(a01); free(a10); free(a11); free(x); free(y); free(h);
  return s;
}

/* ---- CGAL semi-static filters (constants verbatim) --------------------------------------- */
static inline double det2(double a00, double a01, double a10, double a11) { /* determinant.h:25-35 */
  return a00 * a11 - a10 * a01;
}
static inline double det3(double a00, double a01, double a02, double a10, double a11, double a12,
                          double a20, double a21, double a22) {                /* determinant.h:37-52 */
  const double m01 = a00 * a11 - a10 * a01;
  const double m02 = a00 * a21 - a20 * a01;
  const double m12 = a10 * a21 - a20 * a11;
  const double m012 = m01 * a22 - m02 * a12 + m12 * a02;
  return m012;
}
/* CGAL/internal/Static_filters/Orientation_2.h:54-91 */
int orc_orient2d_filter(const double* p, const double* q, const double* r) {
  double pqx = q[0] - p[0], pqy = q[1] - p[1], prx = r[0] - p[0], pry = r[1] - p[1];
  double det = det2(pqx, pqy, prx, pry);
  double maxx = fabs(pqx), maxy = fabs(pqy), aprx = fabs(prx), apry = fabs(pry);
  if (maxx < aprx) maxx = aprx;
  if (maxy < apry) maxy = apry;
  if (maxx > maxy) { double t = maxx; maxx = maxy; maxy = t; }
  if (maxx < 1e-146) { if (maxx == 0) return 0; }
  else if (maxy < 1e153) {
    double eps = 8.8872057372592798e-16 * maxx * maxy;
    if (det > eps) return 1;
    if (det < -eps) return -1;
  }
  return 2;
}
/* CGAL/internal/Static_filters/Orientation_3.h:60-138 */
int orc_orient3d_filter(const double* p, const double* q, const double* r, const double* s) {
  double pqx = q[0] - p[0], pqy = q[1] - p[1], pqz = q[2] - p[2];
  double prx = r[0] - p[0], pry = r[1] - p[1], prz = r[2] - p[2];
  double psx = s[0] - p[0], psy = s[1] - p[1], psz = s[2] - p[2];
  double maxx = fabs(pqx), maxy = fabs(pqy), maxz = fabs(pqz);
  double aprx = fabs(prx), apsx = fabs(psx), apry = fabs(pry), apsy = fabs(psy), aprz = fabs(prz), apsz = fabs(psz);
  if (maxx < aprx) maxx = aprx;
  if (maxx < apsx) maxx = apsx;
  if (maxy < apry) maxy = apry;
  if (maxy < apsy) maxy = apsy;
  if (maxz < aprz) maxz = aprz;
  if (maxz < apsz) maxz = apsz;
  double det = det3(pqx, pqy, pqz, prx, pry, prz, psx, psy, psz);
  double eps = 5.1107127829973299e-15 * maxx * maxy * maxz;
  double t;
  if (maxx > maxz) { t = maxx; maxx = maxz; maxz = t; }
  if (maxy > maxz) { t = maxy; maxy = maxz; maxz = t; }
  else if (maxy < maxx) { t = maxx; maxx = maxy; maxy = t; }
  if (maxx < 1e-97) { if (maxx == 0) return 0; }
  else if (maxz < 1e102) {
    if (det > eps) return 1;
    if (det < -eps) return -1;
  }
  return 2;
}
/* CGAL/internal/Static_filters/Side_of_oriented_circle_2.h:45-102 */
int orc_incircle_filter(const double* p, const double* q, const double* r, const double* t) {
  double qpx = q[0] - p[0], qpy = q[1] - p[1], rpx = r[0] - p[0], rpy = r[1] - p[1];
  double tpx = t[0] - p[0], tpy = t[1] - p[1], tqx = t[0] - q[0], tqy = t[1] - q[1];
  double rqx = r[0] - q[0], rqy = r[1] - q[1];
  double det = det2(qpx * tpy - qpy * tpx, tpx * tqx + tpy * tqy,
                    qpx * rpy - qpy * rpx, rpx * rqx + rpy * rqy);
  double maxx = fabs(qpx), maxy = fabs(qpy);
  double arpx = fabs(rpx), arpy = fabs(rpy), atqx = fabs(tqx), atqy = fabs(tqy);
  double atpx = fabs(tpx), atpy = fabs(tpy), arqx = fabs(rqx), arqy = fabs(rqy);
  if (maxx < arpx) maxx = arpx;
  if (maxx < atpx) maxx = atpx;
  if (maxx < atqx) maxx = atqx;
  if (maxx < arqx) maxx = arqx;
  if (maxy < arpy) maxy = arpy;
  if (maxy < atpy) maxy = atpy;
  if (maxy < atqy) maxy = atqy;
  if (maxy < arqy) maxy = arqy;
  if (maxx > maxy) { double s = maxx; maxx = maxy; maxy = s; }
  if (maxx < 1e-73) { if (maxx == 0) return 0; }
  else if (maxy < 1e76) {
    double eps = 8.8878565762001373e-15 * maxx * maxy * (maxy * maxy);
    if (det > eps) return 1;
    if (det < -eps) return -1;
  }
  return 2;
}
/* filtered predicate = filter, else exact (Filtered_predicate.h:86-106; result is the true sign) */
static inline int orient2d_f(const double* p, const double* q, const double* r) {
  int s = orc_orient2d_filter(p, q, r); return s != 2 ? s : orc_orient2d(p, q, r);
}
static inline int orient3d_f(const double* p, const double* q, const double* r, const double* s4) {
  int s = orc_orient3d_filter(p, q, r, s4); return s != 2 ? s : orc_orient3d(p, q, r, s4);
}
static inline int incircle_f(const double* p, const double* q, const double* r, const double* t) {
  int s = orc_incircle_filter(p, q, r, t); return s != 2 ? s : orc_incircle(p, q, r, t);
}
void orc_orient2d_batch(const double* pts, int64_t n, int8_t* sign) {
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 1024)
  for (int64_t i = 0; i < n; ++i) sign[i] = (int8_t)orient2d_f(pts + 6 * i, pts + 6 * i + 2, pts + 6 * i + 4);
}
void orc_orient3d_batch(const double* pts, int64_t n, int8_t* sign) {
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 1024)
  for (int64_t i = 0; i < n; ++i)
    sign[i] = (int8_t)orient3d_f(pts + 12 * i, pts + 12 * i + 3, pts + 12 * i + 6, pts + 12 * i + 9);
}
void orc_incircle_batch(const double* pts, int64_t n, int8_t* sign) {
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 1024)
  for (int64_t i = 0; i < n; ++i)
    sign[i] = (int8_t)incircle_f(pts + 8 * i, pts + 8 * i + 2, pts + 8 * i + 4, pts + 8 * i + 6);
}

/* ======================================================================================= */
/* 2. Move + validity                                                                      */
/* ======================================================================================= */

/* dune/mmesh/grid/mmesh.hh:1421-1430: point = center() + shift (one IEEE add per component) */
void orc_move(double* x, const double* shift, int64_t n) {
#pragma omp parallel for num_threads(g_threads)
  for (int64_t i = 0; i < n; ++i) x[i] = x[i] + shift[i];
}
/* CGAL/Cartesian/function_objects.h:918-926 Compute_area_2 on TDS-order vertices
   (CGAL/Triangulation_2.h:1147-1155), used by mmesh.hh:1169-1172 signedVolume_ */
double orc_signed_area(const double* p, const double* q, const double* r) {
  double v1x = q[0] - p[0], v1y = q[1] - p[1], v2x = r[0] - p[0], v2y = r[1] - p[1];
  return det2(v1x, v1y, v2x, v2y) / 2;
}
/* CGAL/Cartesian/function_objects.h:1162-1169 Compute_volume_3 (mmesh.hh:1174-1178) */
double orc_signed_volume(const double* p0, const double* p1, const double* p2, const double* p3) {
  return det3(p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2],
              p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2],
              p3[0] - p0[0], p3[1] - p0[1], p3[2] - p0[2]) / 6;
}
static inline void moved(const double* x, const double* s, int64_t v, int dim, double* out) {
  for (int k = 0; k < dim; ++k) out[k] = s ? x[dim * v + k] + s[dim * v + k] : x[dim * v + k];
}
/* mmesh.hh:1094-1134 ensureVertexMovement: validity test `signedVolume_ <= 0.0` (:1102) on trial-moved
   coordinates; plus the exact CGAL orientation sign the north-star asks for. */
void orc_trial_validate_2d(const double* xy, const double* shift, int64_t nv, const int32_t* cells, int64_t nc,
                           uint8_t* valid_fp, int8_t* orient_sign, int64_t* n_invalid) {
  (void)nv; int64_t ninv = 0;
#pragma omp parallel for num_threads(g_threads) reduction(+ : ninv) schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    double p[2], q[2], r[2];
    moved(xy, shift, cells[3 * c], 2, p); moved(xy, shift, cells[3 * c + 1], 2, q); moved(xy, shift, cells[3 * c + 2], 2, r);
    double a = orc_signed_area(p, q, r);
    uint8_t ok = !(a <= 0.0);
    if (valid_fp) valid_fp[c] = ok;
    if (orient_sign) orient_sign[c] = (int8_t)orient2d_f(p, q, r);
    ninv += !ok;
  }
  if (n_invalid) *n_invalid = ninv;
}
void orc_trial_validate_3d(const double* xyz, const double* shift, int64_t nv, const int32_t* cells, int64_t nc,
                           uint8_t* valid_fp, int8_t* orient_sign, int64_t* n_invalid) {
  (void)nv; int64_t ninv = 0;
#pragma omp parallel for num_threads(g_threads) reduction(+ : ninv) schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    double p[4][3];
    for (int k = 0; k < 4; ++k) moved(xyz, shift, cells[4 * c + k], 3, p[k]);
    double v = orc_signed_volume(p[0], p[1], p[2], p[3]);
    uint8_t ok = !(v <= 0.0);
    if (valid_fp) valid_fp[c] = ok;
    if (orient_sign) orient_sign[c] = (int8_t)orient3d_f(p[0], p[1], p[2], p[3]);
    ninv += !ok;
  }
  if (n_invalid) *n_invalid = ninv;
}

/* ======================================================================================= */
/* 3. Per-edge Delaunay legality: CGAL/Delaunay_triangulation_2.h:663-681                   */
/*    (ON_POSITIVE_SIDE != side_of_oriented_circle(face, mirror_vertex))                   */
/* ======================================================================================= */
void orc_legality(const double* xy, const double* shift, const int32_t* edges, int64_t ne, uint8_t* flag,
                  int64_t* n_illegal) {
  int64_t nill = 0;
#pragma omp parallel for num_threads(g_threads) reduction(+ : nill) schedule(static)
  for (int64_t e = 0; e < ne; ++e) {
    const int32_t* q = edges + 4 * e;
    if (q[3] < 0) { flag[e] = 2; continue; }
    double a[2], b[2], c[2], d[2];
    moved(xy, shift, q[0], 2, a); moved(xy, shift, q[1], 2, b); moved(xy, shift, q[2], 2, c); moved(xy, shift, q[3], 2, d);
    int s = incircle_f(a, b, c, d);
    flag[e] = (uint8_t)(s > 0 ? 0 : 1);
    nill += (s > 0);
  }
  if (n_illegal) *n_illegal = nill;
}

/* ======================================================================================= */
/* 4. Distance field                                                                       */
/* ======================================================================================= */

/* dune/mmesh/remeshing/distance.hh:143-156 (2-D branch) + Plane, dune/mmesh/cgal/triangulationwrapper.hh:14-40.
   Facet corners are AffineGeometry::corner(i) = origin + J^T e_i (dune-geometry, un-vendored). */
double orc_point_segment_distance(const double* vp, const double* v0, const double* v1) {
  double c[2][2];
  c[0][0] = v0[0]; c[0][1] = v0[1];
  c[1][0] = v0[0] + 1.0 * (v1[0] - v0[0]);
  c[1][1] = v0[1] + 1.0 * (v1[1] - v0[1]);
  for (int i = 0; i < 2; ++i) {
    const double* vi = c[i]; const double* vj = c[1 - i];
    double ax = vi[0] - vj[0], ay = vi[1] - vj[1];
    double bx = vp[0] - vi[0], by = vp[1] - vi[1];
    double sgn = 0.0; sgn += ax * bx; sgn += ay * by;             /* DenseVector operator* */
    if (sgn >= 0) { double n2 = 0.0; n2 += bx * bx; n2 += by * by; return sqrt(n2); }
  }
  double n0 = c[0][0] - c[1][0], n1 = c[0][1] - c[1][1];            /* n_ = a; n_ -= b */
  { double t = n0; n0 = n1; n1 = t; }                               /* std::swap(n_[0], n_[1]) */
  n0 *= -1.0;
  double nn = 0.0; nn += n0 * n0; nn += n1 * n1; nn = sqrt(nn);     /* two_norm */
  n0 /= nn; n1 /= nn;
  double dx = vp[0] - c[0][0], dy = vp[1] - c[0][1];
  double sd = 0.0; sd += n0 * dx; sd += n1 * dy;                    /* n_ * (x - p_) */
  return fabs(sd);
}
/* distance.hh:42-59 update + :133-140 handleFacet: min over all interface facets, init 1e100 */
void orc_distance_2d(const double* xy, int64_t nv, const int32_t* segs, int64_t ns, double* dist) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int64_t v = 0; v < nv; ++v) {
    double cur = 1e100;
    for (int64_t s = 0; s < ns; ++s) {
      double d = orc_point_segment_distance(xy + 2 * v, xy + 2 * (int64_t)segs[2 * s], xy + 2 * (int64_t)segs[2 * s + 1]);
      if (d < cur) cur = d;
    }
    dist[v] = cur;
  }
}
void orc_distance_2d_subset(const double* xy, const int32_t* vidx, int64_t nsub, const int32_t* segs, int64_t ns,
                            double* dist_sub) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int64_t i = 0; i < nsub; ++i) {
    int64_t v = vidx[i];
    double cur = 1e100;
    for (int64_t s = 0; s < ns; ++s) {
      double d = orc_point_segment_distance(xy + 2 * v, xy + 2 * (int64_t)segs[2 * s], xy + 2 * (int64_t)segs[2 * s + 1]);
      if (d < cur) cur = d;
    }
    dist_sub[i] = cur;
  }
}
/* CULLED variant (NOT the reference's loop, which is the brute force above): the same per-pair arithmetic, but the facets are
   bucketed in a uniform grid and a vertex visits rings of grid cells around its own until the ring distance exceeds its
   current minimum.  The minimum is taken over a superset of every facet that can attain it, so the values equal the brute
   force bit for bit; only the cost changes (O(Nv * few) instead of O(Nv * Ns)).  Used as the labelled "culled" CPU baseline. */
void orc_distance_2d_culled_subset(const double* xy, int64_t nv_all, const int32_t* vidx, int64_t nsub, const int32_t* segs,
                                   int64_t ns, double* dist_sub) {
  if (ns == 0) { for (int64_t i = 0; i < nsub; ++i) dist_sub[i] = 1e100; return; }
  double lo[2] = { 1e300, 1e300 }, hi[2] = { -1e300, -1e300 };
  for (int64_t v = 0; v < nv_all; ++v) for (int k = 0; k < 2; ++k) { if (xy[2 * v + k] < lo[k]) lo[k] = xy[2 * v + k]; if (xy[2 * v + k] > hi[k]) hi[k] = xy[2 * v + k]; }
  int R = (int)(2.0 * sqrt((double)ns)); if (R < 1) R = 1; if (R > 4096) R = 4096;
  double ext = hi[0] - lo[0] > hi[1] - lo[1] ? hi[0] - lo[0] : hi[1] - lo[1];
  if (!(ext > 0)) ext = 1.0;
  const double cell = ext / R * (1.0 + 1e-12);
  const int RX = (int)((hi[0] - lo[0]) / cell) + 1, RY = (int)((hi[1] - lo[1]) / cell) + 1;
  int64_t* cnt = (int64_t*)calloc((size_t)RX * RY + 1, sizeof(int64_t));
#define CELLOF(x, k) ((int)(((x) - lo[k]) / cell))
  for (int pass = 0; pass < 2; ++pass) {
    static int32_t* items = NULL;
    if (pass == 1) {
      int64_t tot = 0;
      for (int64_t c = 0; c < (int64_t)RX * RY; ++c) { int64_t t = cnt[c]; cnt[c] = tot; tot += t; }
      cnt[(int64_t)RX * RY] = tot;
      items = (int32_t*)malloc((size_t)(tot + 1) * sizeof(int32_t));
    }
    for (int64_t sgi = 0; sgi < ns; ++sgi) {
      const double* a = xy + 2 * (int64_t)segs[2 * sgi]; const double* b = xy + 2 * (int64_t)segs[2 * sgi + 1];
      int x0 = CELLOF(a[0] < b[0] ? a[0] : b[0], 0), x1 = CELLOF(a[0] < b[0] ? b[0] : a[0], 0);
      int y0 = CELLOF(a[1] < b[1] ? a[1] : b[1], 1), y1 = CELLOF(a[1] < b[1] ? b[1] : a[1], 1);
      for (int y = y0; y <= y1; ++y) for (int x = x0; x <= x1; ++x) { if (pass == 0) cnt[(int64_t)y * RX + x]++; else items[cnt[(int64_t)y * RX + x]++] = (int32_t)sgi; }
    }
    if (pass == 1) {
      /* cnt now holds the END offsets: shift back to starts */
      for (int64_t c = (int64_t)RX * RY; c > 0; --c) cnt[c] = cnt[c - 1];
      cnt[0] = 0;
#pragma omp parallel for num_threads(g_threads) schedule(static)
      for (int64_t i = 0; i < nsub; ++i) {
        const double* vp = xy + 2 * (int64_t)vidx[i];
        const int cx = CELLOF(vp[0], 0), cy = CELLOF(vp[1], 1);
        double cur = 1e100;
        const int rmax = RX > RY ? RX : RY;
        for (int r = 0; r <= rmax; ++r) {
          /* everything outside the (2r-1)^2 block around the vertex's cell is at least (r-1) * cell away */
          if (r >= 1 && cur < (double)(r - 1) * cell * (1.0 - 1e-9)) break;
          for (int y = cy - r; y <= cy + r; ++y) {
            if (y < 0 || y >= RY) continue;
            const int step = (y == cy - r || y == cy + r) ? 1 : 2 * r;
            for (int x = cx - r; x <= cx + r; x += (step > 0 ? step : 1)) {
              if (x < 0 || x >= RX) continue;
              for (int64_t e = cnt[(int64_t)y * RX + x]; e < cnt[(int64_t)y * RX + x + 1]; ++e) {
                const int64_t sg = items[e];
                const double d = orc_point_segment_distance(vp, xy + 2 * (int64_t)segs[2 * sg], xy + 2 * (int64_t)segs[2 * sg + 1]);
                if (d < cur) cur = d;
              }
            }
          }
        }
        dist_sub[i] = cur;
      }
      free(items); items = NULL;
    }
  }
#undef CELLOF
  free(cnt);
}
/* distance.hh:157-173 (3-D estimate: corners, edge midpoints, centroid) */
void orc_distance_3d(const double* xyz, int64_t nv, const int32_t* tris, int64_t ns, double* dist) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int64_t v = 0; v < nv; ++v) {
    const double* vp = xyz + 3 * v;
    double cur = 1e100;
    for (int64_t s = 0; s < ns; ++s) {
      const double* a = xyz + 3 * (int64_t)tris[3 * s];
      const double* b = xyz + 3 * (int64_t)tris[3 * s + 1];
      const double* cc = xyz + 3 * (int64_t)tris[3 * s + 2];
      double c[3][3], ctr[3];
      const double third = 1.0 / 3.0;
      for (int k = 0; k < 3; ++k) {
        double j0 = b[k] - a[k], j1 = cc[k] - a[k];
        c[0][k] = a[k];
        c[1][k] = (a[k] + 1.0 * j0) + 0.0 * j1;   /* origin + J^T (1,0) */
        c[2][k] = (a[k] + 0.0 * j0) + 1.0 * j1;   /* origin + J^T (0,1) */
        ctr[k] = (a[k] + third * j0) + third * j1; /* center() = global(1/3,1/3) */
      }
      double d = 1e100;
      for (int i = 0; i < 3; ++i) {
        const double* vi = c[i]; const double* vj = c[(i + 1) % 3];
        double n2 = 0.0;
        for (int k = 0; k < 3; ++k) { double t = vp[k] - vi[k]; n2 += t * t; }
        if (n2 < d) d = n2;
        double m2 = 0.0;
        for (int k = 0; k < 3; ++k) { double t = vp[k] - 0.5 * (vi[k] + vj[k]); m2 += t * t; }
        if (m2 < d) d = m2;
      }
      double c2 = 0.0;
      for (int k = 0; k < 3; ++k) { double t = vp[k] - ctr[k]; c2 += t * t; }
      if (c2 < d) d = c2;
      d = sqrt(d);
      if (d < cur) cur = d;
    }
    dist[v] = cur;
  }
}
/* distance.hh:118-123 */
double orc_distance_max(const double* dist, int64_t nv) {
  double m = 0.0;
  for (int64_t i = 0; i < nv; ++i) m = (m < dist[i]) ? dist[i] : m; /* std::max(d, maximum) */
  return m;
}

/* ======================================================================================= */
/* 5. RatioIndicator                                                                       */
/* ======================================================================================= */
static inline double edge_len2d(const double* a, const double* b) {
  /* AffineGeometry<ct,1,2>::volume(): sqrt(J J^T) via AAT_L + Cholesky (dune-geometry, un-vendored) */
  double jx = b[0] - a[0], jy = b[1] - a[1];
  double s = 0.0; s += jx * jx; s += jy * jy;
  return sqrt(s);
}
static inline double edge_len3d(const double* a, const double* b) {
  double s = 0.0;
  for (int k = 0; k < 3; ++k) { double j = b[k] - a[k]; s += j * j; }
  return sqrt(s);
}
/* ratioindicator.hh:62-91 */
void orc_indicator_init_2d(const double* xy, const int32_t* edges, int64_t ne, const int32_t* segs, int64_t ns,
                           orc_params* prm) {
  const double K = 2.0, k = K / (2.0 * 4.0);
  double maxH = 0.0, minH = 1e100;
  for (int64_t s = 0; s < ns; ++s) {
    double h = edge_len2d(xy + 2 * (int64_t)segs[2 * s], xy + 2 * (int64_t)segs[2 * s + 1]);
    maxH = (maxH < h) ? h : maxH; minH = (h < minH) ? h : minH;
  }
  maxH = K * maxH; minH = k * minH;
  double maxh = 0.0, minh = 1e100;
  for (int64_t e = 0; e < ne; ++e) {
    double h = edge_len2d(xy + 2 * (int64_t)edges[4 * e], xy + 2 * (int64_t)edges[4 * e + 1]);
    maxh = (maxh < h) ? h : maxh; minh = (h < minh) ? h : minh;
  }
  if (ns == 0) { maxH = K * maxh; minH = k * minh; }
  prm->maxH = maxH; prm->minH = minH; prm->factor = maxh / minh;
}

static inline void corners2d(const double* xy, const int32_t* cell, uint8_t perm, double c[3][2]) {
  for (int k = 0; k < 3; ++k) {
    int t = (perm >> (2 * k)) & 3;
    c[k][0] = xy[2 * (int64_t)cell[t]]; c[k][1] = xy[2 * (int64_t)cell[t] + 1];
  }
}
/* CGAL/constructions/kernel_ftC2.h:40-74 circumcenterC2 */
static inline void circumcenter2d(const double* p, const double* q, const double* r, double* out) {
  double dqx = q[0] - p[0], dqy = q[1] - p[1], drx = r[0] - p[0], dry = r[1] - p[1];
  double r2 = drx * drx + dry * dry;
  double q2 = dqx * dqx + dqy * dqy;
  double den = 2 * det2(dqx, dqy, drx, dry);
  double dcx = det2(dry, dqy, r2, q2) / den;
  double dcy = -det2(drx, dqx, r2, q2) / den;
  out[0] = dcx + p[0]; out[1] = dcy + p[1];
}
/* ratioindicator.hh:106-160 for a bulk triangle; corners c0..c2 in id order (geometry.hh:76-104) */
static inline int indicator2d(const double c[3][2], const double d[3], const orc_params* prm, double maxDist,
                              int* longest) {
  int refine = 0, coarse = 0;
  double minE = 1e100, maxE = -1e100, sumE = 0.0;
  double dist = 0.0; dist += d[0]; dist += d[1]; dist += d[2]; dist /= 3;     /* distance.hh:92-101 */
  dist = (dist < maxDist) ? dist : maxDist;                                  /* std::min(maxDist_, dist) */
  const double l = dist / maxDist;
  const double minH = (1. - l) * prm->minH + l * prm->factor * prm->minH;
  const double maxH = (1. - l) * prm->maxH + l * prm->factor * prm->maxH;
  static const int EV[3][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 } };              /* common.hh:68-78 DUNE edge order */
  double longestLen = -1e100; int longestIdx = -1;
  for (int i = 0; i < 3; ++i) {
    const double len = edge_len2d(c[EV[i][0]], c[EV[i][1]]);
    if (len < minH) coarse++;
    if (len > maxH) refine++;
    minE = (len < minE) ? len : minE;
    maxE = (maxE < len) ? len : maxE;
    sumE += len;
    if (len > longestLen) { longestLen = len; longestIdx = i; }              /* longestedgerefinement.hh:41-57 */
  }
  const double edgeRatio = maxE / minE;
  if (edgeRatio > prm->edgeRatio) coarse++;
  /* AffineGeometry: J rows = c1-c0, c2-c0; volume = |det| * refVolume(1/2); corner(i) = origin + J^T e_i */
  double j00 = c[1][0] - c[0][0], j01 = c[1][1] - c[0][1], j10 = c[2][0] - c[0][0], j11 = c[2][1] - c[0][1];
  double vol = fabs(j00 * j11 - j10 * j01) * 0.5;
  double g0[2] = { c[0][0], c[0][1] };
  double g1[2] = { (c[0][0] + 1.0 * j00) + 0.0 * j10, (c[0][1] + 1.0 * j01) + 0.0 * j11 };
  double g2[2] = { (c[0][0] + 0.0 * j00) + 1.0 * j10, (c[0][1] + 0.0 * j01) + 1.0 * j11 };
  const double innerRadius = 2 * vol / sumE;
  double cc[2]; circumcenter2d(g0, g1, g2, cc);
  double ox = cc[0] - g0[0], oy = cc[1] - g0[1];
  double o2 = 0.0; o2 += ox * ox; o2 += oy * oy;
  const double outerRadius = sqrt(o2);
  const double radiusRatio = outerRadius / (innerRadius * 2);
  if (radiusRatio > prm->radiusRatio) coarse++;
  if (longest) *longest = longestIdx;
  if (coarse > 0) return -1;
  if (refine > 0) return 1;
  return 0;
}
/* mmesh.hh:736-751 markElements (bulk loop; indicator evaluated once, quirk 3) */
void orc_mark_2d(const double* xy, const double* dist, const int32_t* cells, const uint8_t* perm, int64_t nc,
                 const orc_params* prm, double maxDist, int8_t* mark, uint8_t* longest_edge, int64_t counts[2]) {
  int64_t nref = 0, ncoa = 0;
#pragma omp parallel for num_threads(g_threads) reduction(+ : nref, ncoa) schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    double cr[3][2], d[3]; int le;
    corners2d(xy, cells + 3 * c, perm[c], cr);
    for (int k = 0; k < 3; ++k) d[k] = dist[cells[3 * c + ((perm[c] >> (2 * k)) & 3)]];
    int m = indicator2d(cr, d, prm, maxDist, &le);
    mark[c] = (int8_t)m;
    if (longest_edge) longest_edge[c] = (uint8_t)le;
    nref += (m > 0); ncoa += (m < 0);
  }
  if (counts) { counts[0] = nref; counts[1] = ncoa; }
}
/* interface elements through the same operator() (mmesh.hh:745-748; distance_(ielement)=0, distance.hh:104):
   one "edge" = the segment, geo.volume() = len, circumcenter = midpoint (pointfieldvector.hh:95) */
void orc_mark_interface_2d(const double* xy, const int32_t* segs, int64_t ns, const orc_params* prm, double maxDist,
                           int8_t* mark) {
  for (int64_t s = 0; s < ns; ++s) {
    const double* a = xy + 2 * (int64_t)segs[2 * s]; const double* b = xy + 2 * (int64_t)segs[2 * s + 1];
    int refine = 0, coarse = 0;
    double dist = (0.0 < maxDist) ? 0.0 : maxDist;
    const double l = dist / maxDist;
    const double minH = (1. - l) * prm->minH + l * prm->factor * prm->minH;
    const double maxH = (1. - l) * prm->maxH + l * prm->factor * prm->maxH;
    const double len = edge_len2d(a, b);
    if (len < minH) coarse++;
    if (len > maxH) refine++;
    double minE = (len < 1e100) ? len : 1e100, maxE = (-1e100 < len) ? len : -1e100, sumE = 0.0 + len;
    if (maxE / minE > prm->edgeRatio) coarse++;
    const double innerRadius = 2 * len / sumE;
    double jx = b[0] - a[0], jy = b[1] - a[1];
    double c1x = a[0] + 1.0 * jx, c1y = a[1] + 1.0 * jy;
    double ccx = 0.5 * (a[0] + c1x), ccy = 0.5 * (a[1] + c1y);
    double ox = ccx - a[0], oy = ccy - a[1];
    double o2 = 0.0; o2 += ox * ox; o2 += oy * oy;
    const double radiusRatio = sqrt(o2) / (innerRadius * 2);
    if (radiusRatio > prm->radiusRatio) coarse++;
    mark[s] = (int8_t)(coarse > 0 ? -1 : (refine > 0 ? 1 : 0));
  }
}
/* longestedgerefinement.hh:92-112: the comparator handed to std::sort (a, b = (vertex, summed edge length)) */
typedef struct { int idx; int64_t level; double len; int constrained; } coarsen_key;
static inline int coarsen_less(const coarsen_key* a, const coarsen_key* b) {
  if (a->level < b->level) return 1;                    /* :95-97 */
  if (a->len < b->len - 1e-14) return 1;                /* :100 */
  if (!a->constrained && b->constrained) return 1;      /* :103-106 */
  return 0;
}
/* libstdc++ std::sort on <= 16 elements is __insertion_sort (bits/stl_algo.h): element i moves to the front iff
   comp(*i, *first); otherwise __unguarded_linear_insert shifts it left while comp(val, *next). */
static void insertion_sort_keys(coarsen_key* v, int n) {
  for (int i = 1; i < n; ++i) {
    coarsen_key val = v[i];
    if (coarsen_less(&val, &v[0])) {
      for (int j = i; j > 0; --j) v[j] = v[j - 1];
      v[0] = val;
    } else {
      int j = i;
      while (coarsen_less(&val, &v[j - 1])) { v[j] = v[j - 1]; --j; }
      v[j] = val;
    }
  }
}
void orc_coarsening_order(int n, const int64_t* level, const double* len, const uint8_t* constrained, int* order) {
  coarsen_key v[8];
  if (n > 8) n = 8;
  for (int k = 0; k < n; ++k) { v[k].idx = k; v[k].level = level[k]; v[k].len = len[k]; v[k].constrained = constrained[k]; }
  insertion_sort_keys(v, n);
  for (int k = 0; k < n; ++k) order[k] = v[k].idx;
}
int orc_coarsening_choice_2d(const double corners6[6], const int64_t level[3], const uint8_t constrained[3]) {
  static const int EV[3][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 } };
  coarsen_key v[3];
  for (int k = 0; k < 3; ++k) { v[k].idx = k; v[k].level = level[k]; v[k].len = 0.0; v[k].constrained = constrained[k]; }
  for (int i = 0; i < 3; ++i) {                          /* :80-90 */
    const double len = edge_len2d(corners6 + 2 * EV[i][0], corners6 + 2 * EV[i][1]);
    v[EV[i][0]].len += len; v[EV[i][1]].len += len;
  }
  insertion_sort_keys(v, 3);
  return v[0].idx;
}
/* tiny open-addressing set of 64-bit keys (inserted_ / removed_ of mmesh.hh:2022-2023) */
typedef struct { uint64_t* slot; uint64_t mask; } keyset;
static int keyset_init(keyset* s, int64_t n) {
  uint64_t cap = 16; while (cap < (uint64_t)(2 * n + 2)) cap <<= 1;
  s->slot = (uint64_t*)malloc(cap * sizeof(uint64_t)); s->mask = cap - 1;
  if (!s->slot) return 0;
  memset(s->slot, 0xff, cap * sizeof(uint64_t));
  return 1;
}
static int keyset_insert(keyset* s, uint64_t key) {       /* 1 if newly inserted */
  uint64_t h = (key * 0x9E3779B97F4A7C15ull) & s->mask;
  while (s->slot[h] != ~0ull) { if (s->slot[h] == key) return 0; h = (h + 1) & s->mask; }
  s->slot[h] = key; return 1;
}
static int keyset_has(const keyset* s, uint64_t key) {
  uint64_t h = (key * 0x9E3779B97F4A7C15ull) & s->mask;
  while (s->slot[h] != ~0ull) { if (s->slot[h] == key) return 1; h = (h + 1) & s->mask; }
  return 0;
}
static inline uint64_t edge_key(int a, int b) { return a < b ? ((uint64_t)a << 32) | (uint32_t)b : ((uint64_t)b << 32) | (uint32_t)a; }
void orc_adapt_candidates_2d(const double* xy, int64_t nv, const int32_t* cells, const uint8_t* perm, int64_t nc,
                             const int32_t* cell_nb, const int8_t* mark, const uint8_t* longest_edge,
                             const int32_t* segs, int64_t ns, const int8_t* imark, const int32_t* vertex_level,
                             int32_t* ins_cells, int64_t* n_ins, int32_t* rem_cells, int32_t* rem_vertices, int64_t* n_rem,
                             int32_t* ins_segs, int64_t* n_ins_segs, int32_t* rem_segs, int32_t* rem_seg_vertices,
                             int64_t* n_rem_segs) {
  static const int EV[3][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 } };
  uint8_t* v_if = (uint8_t*)calloc((size_t)nv + 1, 1);    /* VertexInfo::isInterface */
  uint8_t* v_bnd = (uint8_t*)calloc((size_t)nv + 1, 1);   /* atBoundary(v): incident to the infinite vertex (:114-120) */
  int32_t* v_cnt = (int32_t*)calloc((size_t)nv + 1, sizeof(int32_t));   /* incident interface elements (:160-168) */
  keyset iface, inserted, removed;
  keyset_init(&iface, ns); keyset_init(&inserted, nc + ns); keyset_init(&removed, nc + ns);
  for (int64_t s = 0; s < ns; ++s) {
    const int a = segs[2 * s], b = segs[2 * s + 1];
    v_if[a] = v_if[b] = 1; v_cnt[a]++; v_cnt[b]++;
    keyset_insert(&iface, edge_key(a, b));
  }
  for (int64_t c = 0; c < nc; ++c)
    for (int k = 0; k < 3; ++k)
      if (cell_nb[3 * c + k] < 0) { v_bnd[cells[3 * c + (k + 1) % 3]] = 1; v_bnd[cells[3 * c + (k + 2) % 3]] = 1; }
  int64_t ni = 0, nr = 0;
  for (int64_t c = 0; c < nc; ++c) {                      /* mmesh.hh:828-866 */
    int k[3];
    for (int t = 0; t < 3; ++t) k[t] = cells[3 * c + ((perm[c] >> (2 * t)) & 3)];
    if (mark[c] == 1) {
      const int e = longest_edge[c];
      const int a = k[EV[e][0]], b = k[EV[e][1]];
      if (!keyset_has(&iface, edge_key(a, b)))            /* :844 isInterface(ip.edge) */
        if (keyset_insert(&inserted, edge_key(a, b))) ins_cells[ni++] = (int32_t)c;
    } else if (mark[c] == -1) {
      double cr[6]; int64_t lev[3]; uint8_t con[3];
      for (int t = 0; t < 3; ++t) {
        cr[2 * t] = xy[2 * (int64_t)k[t]]; cr[2 * t + 1] = xy[2 * (int64_t)k[t] + 1];
        lev[t] = vertex_level ? vertex_level[k[t]] : 0;
        con[t] = (uint8_t)(v_if[k[t]] || v_bnd[k[t]]);
      }
      const int v = k[orc_coarsening_choice_2d(cr, lev, con)];
      if (keyset_insert(&removed, (uint64_t)v)) { rem_cells[nr] = (int32_t)c; rem_vertices[nr] = v; ++nr; }
    }
  }
  *n_ins = ni; *n_rem = nr;
  int64_t nis = 0, nrs = 0;
  if (imark) {
    for (int64_t s = 0; s < ns; ++s) {                    /* mmesh.hh:869-908 */
      if (imark[s] == 1) {
        if (keyset_insert(&inserted, edge_key(segs[2 * s], segs[2 * s + 1]))) ins_segs[nis++] = (int32_t)s;
      } else if (imark[s] == -1) {
        int v = -1;
        for (int j = 0; j < 2 && v < 0; ++j) if (v_cnt[segs[2 * s + j]] == 2) v = segs[2 * s + j];
        if (v >= 0 && keyset_insert(&removed, (uint64_t)v)) { rem_segs[nrs] = (int32_t)s; rem_seg_vertices[nrs] = v; ++nrs; }
      }
    }
  }
  if (n_ins_segs) *n_ins_segs = nis;
  if (n_rem_segs) *n_rem_segs = nrs;
  free(v_if); free(v_bnd); free(v_cnt); free(iface.slot); free(inserted.slot); free(removed.slot);
}
/* 3-D bulk loop of adapt() (mmesh.hh:828-851; coarsening never fires in 3-D, ratioindicator.hh:151-156).
   isInterface(edge) in 3-D = the edge belongs to an interface triangle (mmesh.hh:624-645). */
void orc_refine_candidates_3d(const int32_t* cells, const uint8_t* perm, int64_t nc, const int8_t* mark,
                              const uint8_t* longest_edge, const int32_t* tris, int64_t ns,
                              int32_t* ins_cells, int32_t* ins_edge_vertices, int64_t* n_ins) {
  static const int EV[6][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 }, { 0, 3 }, { 1, 3 }, { 2, 3 } };
  keyset iface, inserted;
  keyset_init(&iface, 3 * ns); keyset_init(&inserted, nc);
  for (int64_t t = 0; t < ns; ++t) {
    const int32_t* q = tris + 3 * t;
    keyset_insert(&iface, edge_key(q[0], q[1])); keyset_insert(&iface, edge_key(q[0], q[2])); keyset_insert(&iface, edge_key(q[1], q[2]));
  }
  int64_t ni = 0;
  for (int64_t c = 0; c < nc; ++c) {
    if (mark[c] != 1) continue;
    int k[4];
    for (int t = 0; t < 4; ++t) k[t] = cells[4 * c + ((perm[c] >> (2 * t)) & 3)];
    const int e = longest_edge[c];
    const int a = k[EV[e][0]], b = k[EV[e][1]];
    if (keyset_has(&iface, edge_key(a, b))) continue;
    if (keyset_insert(&inserted, edge_key(a, b))) {
      ins_cells[ni] = (int32_t)c;
      if (ins_edge_vertices) { ins_edge_vertices[2 * ni] = a; ins_edge_vertices[2 * ni + 1] = b; }
      ++ni;
    }
  }
  *n_ins = ni;
  free(iface.slot); free(inserted.slot);
}
/* 3-D: 6 edges in DUNE tetrahedron edge order; coarsening disabled (ratioindicator.hh:151-156) so only the
   refine count decides; volume / circumradius cannot influence the result and are not evaluated. */
void orc_mark_3d(const double* xyz, const double* dist, const int32_t* cells, const uint8_t* perm, int64_t nc,
                 const orc_params* prm, double maxDist, int8_t* mark, uint8_t* longest_edge, int64_t counts[2]) {
  static const int EV[6][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 }, { 0, 3 }, { 1, 3 }, { 2, 3 } };
  int64_t nref = 0;
#pragma omp parallel for num_threads(g_threads) reduction(+ : nref) schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    const double* cr[4]; double dsum = 0.0;
    for (int k = 0; k < 4; ++k) {
      int v = cells[4 * c + ((perm[c] >> (2 * k)) & 3)];
      cr[k] = xyz + 3 * (int64_t)v; dsum += dist[v];
    }
    dsum /= 4;
    double d = (dsum < maxDist) ? dsum : maxDist;
    const double l = d / maxDist;
    const double maxH = (1. - l) * prm->maxH + l * prm->factor * prm->maxH;
    int refine = 0; double longestLen = -1e100; int longestIdx = -1;
    for (int i = 0; i < 6; ++i) {
      const double len = edge_len3d(cr[EV[i][0]], cr[EV[i][1]]);
      if (len > maxH) refine++;
      if (len > longestLen) { longestLen = len; longestIdx = i; }
    }
    int m = refine > 0 ? 1 : 0;
    mark[c] = (int8_t)m; if (longest_edge) longest_edge[c] = (uint8_t)longestIdx;
    nref += (m > 0);
  }
  if (counts) { counts[0] = nref; counts[1] = 0; }
}
/* Values asserted by dune/mmesh/test/test-mmesh-2d.cc:139-223 (center, volume, edge centers) */
void orc_cell_geometry_2d(const double* xy, const int32_t* cell, uint8_t perm, double* center2, double* volume,
                          double* edge_centers6, double* circumcenter2) {
  static const int EV[3][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 } };
  double c[3][2]; corners2d(xy, cell, perm, c);
  double j00 = c[1][0] - c[0][0], j01 = c[1][1] - c[0][1], j10 = c[2][0] - c[0][0], j11 = c[2][1] - c[0][1];
  const double third = 1.0 / 3.0;
  if (center2) { center2[0] = (c[0][0] + third * j00) + third * j10; center2[1] = (c[0][1] + third * j01) + third * j11; }
  if (volume) *volume = fabs(j00 * j11 - j10 * j01) * 0.5;
  if (edge_centers6)
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 2; ++k)
      edge_centers6[2 * i + k] = c[EV[i][0]][k] + 0.5 * (c[EV[i][1]][k] - c[EV[i][0]][k]);
  if (circumcenter2) {
    double g1[2] = { (c[0][0] + 1.0 * j00) + 0.0 * j10, (c[0][1] + 1.0 * j01) + 0.0 * j11 };
    double g2[2] = { (c[0][0] + 0.0 * j00) + 1.0 * j10, (c[0][1] + 0.0 * j01) + 1.0 * j11 };
    circumcenter2d(c[0], g1, g2, circumcenter2);
  }
}
/* Values asserted by dune/mmesh/test/test-mmesh-3d.cc:137-152,192 (center, volume);
   circumcenter: CGAL/Cartesian/function_objects.h:2203-2241 */
void orc_cell_geometry_3d(const double* xyz, const int32_t* cell, uint8_t perm, double* center3, double* volume,
                          double* circumcenter3) {
  double c[4][3];
  for (int k = 0; k < 4; ++k) for (int j = 0; j < 3; ++j) c[k][j] = xyz[3 * (int64_t)cell[(perm >> (2 * k)) & 3] + j];
  double J[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J[i][j] = c[i + 1][j] - c[0][j];
  if (center3) for (int j = 0; j < 3; ++j) center3[j] = ((c[0][j] + 0.25 * J[0][j]) + 0.25 * J[1][j]) + 0.25 * J[2][j];
  if (volume) *volume = fabs(det3(J[0][0], J[0][1], J[0][2], J[1][0], J[1][1], J[1][2], J[2][0], J[2][1], J[2][2])) * (1.0 / 6.0);
  if (circumcenter3) {
    const double *p = c[0], *q = c[1], *r = c[2], *s = c[3];
    double qpx = q[0] - p[0], qpy = q[1] - p[1], qpz = q[2] - p[2];
    double qp2 = qpx * qpx + qpy * qpy + qpz * qpz;
    double rpx = r[0] - p[0], rpy = r[1] - p[1], rpz = r[2] - p[2];
    double rp2 = rpx * rpx + rpy * rpy + rpz * rpz;
    double spx = s[0] - p[0], spy = s[1] - p[1], spz = s[2] - p[2];
    double sp2 = spx * spx + spy * spy + spz * spz;
    double num_x = det3(qpy, qpz, qp2, rpy, rpz, rp2, spy, spz, sp2);
    double num_y = det3(qpx, qpz, qp2, rpx, rpz, rp2, spx, spz, sp2);
    double num_z = det3(qpx, qpy, qp2, rpx, rpy, rp2, spx, spy, sp2);
    double den = det3(qpx, qpy, qpz, rpx, rpy, rpz, spx, spy, spz);
    double inv = 1 / (2 * den);
    circumcenter3[0] = p[0] + num_x * inv; circumcenter3[1] = p[1] - num_y * inv; circumcenter3[2] = p[2] + num_z * inv;
  }
}

/* ======================================================================================= */
/* 6. Clipping + projection                                                                */
/* ======================================================================================= */
#define MAXPOLY 12

/* dune/mmesh/grid/polygoncutting.hh:100-114 lineIntersectionPoint */
static inline void line_isect(const double* p1, const double* p2, const double* q1, const double* q2, double* out) {
  double dPx = p1[0] - p2[0], dPy = p1[1] - p2[1];
  double dQx = q1[0] - q2[0], dQy = q1[1] - q2[1];
  double scale = dPx * dQy - dQx * dPy;
  double fq = q1[0] * q2[1] - q1[1] * q2[0];
  double fp = p1[0] * p2[1] - p1[1] * p2[0];
  dPx *= fq; dPy *= fq;
  dQx *= fp; dQy *= fp;
  double ix = dQx - dPx, iy = dQy - dPy;
  out[0] = ix / scale; out[1] = iy / scale;
}
/* one clip edge (c0 -> c1) of polygoncutting.hh:31-93 applied to a polygon of nin points; returns the point count, -1 on
   overflow of MAXPOLY */
static int sh_clip_edge(const double in[][2], int nin, const double* c0, const double* c1, double eps, double out[][2]) {
  int nout = 0;
  for (int ii = 0; ii < nin; ++ii) {
    const int in2 = (ii + 1) % nin;
    double dQx = c1[0] - c0[0], dQy = c1[1] - c0[1];
    double q2xq1 = c0[1] * c1[0] - c0[0] * c1[1];
    double first = -dQx * in[ii][1] + dQy * in[ii][0] + q2xq1;
    double second = -dQx * in[in2][1] + dQy * in[in2][0] + q2xq1;
    if (first > eps && second < -eps) {
      line_isect(in[ii], in[in2], c0, c1, out[nout]); nout++;
      out[nout][0] = in[in2][0]; out[nout][1] = in[in2][1]; nout++;
    } else if (first < -eps && second > eps) {
      line_isect(in[ii], in[in2], c0, c1, out[nout]); nout++;
    } else if (first >= -eps && second > eps) {
      continue;
    } else {
      out[nout][0] = in[in2][0]; out[nout][1] = in[in2][1]; nout++;
    }
    if (nout > MAXPOLY - 2) return -1; /* cannot happen for triangle/triangle (bound 9) */
  }
  return nout;
}
/* polygoncutting.hh:17-96 sutherlandHodgman; returns point count (capacity MAXPOLY, rigorous bound 9) */
static int sh_clip(const double subj[3][2], const double clip[3][2], double eps, double out[MAXPOLY][2]) {
  double in[MAXPOLY][2]; int nout = 3;
  for (int i = 0; i < 3; ++i) { out[i][0] = subj[i][0]; out[i][1] = subj[i][1]; }
  for (int ci = 0; ci < 3; ++ci) {
    const int cn = (ci + 1) % 3;
    int nin = nout; memcpy(in, out, sizeof(double) * 2 * (size_t)nin);
    nout = sh_clip_edge(in, nin, clip[ci], clip[cn], eps, out);
    if (nout < 0) return -1;
  }
  return nout;
}
/* one clip edge on a general polygon (<= 6 points in, <= 9 out): unit-test hook for the product's clip stage */
int orc_clip_edge(const double* poly, int n, const double* q1, const double* q2, double eps, double* out) {
  double in[MAXPOLY][2], o[MAXPOLY][2];
  if (n < 0 || n > 6) return -1;
  for (int i = 0; i < n; ++i) { in[i][0] = poly[2 * i]; in[i][1] = poly[2 * i + 1]; }
  const int m = n > 0 ? sh_clip_edge(in, n, q1, q2, eps, o) : 0;
  for (int i = 0; i < m; ++i) { out[2 * i] = o[i][0]; out[2 * i + 1] = o[i][1]; }
  return m;
}
int orc_sutherland_hodgman(const double* subj6, const double* clip6, double eps, double* out) {
  double s[3][2], c[3][2], o[MAXPOLY][2];
  memcpy(s, subj6, sizeof s); memcpy(c, clip6, sizeof c);
  int n = sh_clip(s, c, eps, o);
  for (int i = 0; i < n; ++i) { out[2 * i] = o[i][0]; out[2 * i + 1] = o[i][1]; }
  return n;
}
void orc_sutherland_hodgman_batch(const double* subj, const double* clip, int64_t n, double eps, double* out18, int32_t* count) {
  for (int64_t i = 0; i < n; ++i) {
    double o[2 * MAXPOLY];
    const int m = orc_sutherland_hodgman(subj + 6 * i, clip + 6 * i, eps, o);
    count[i] = m;
    for (int k = 0; k < 2 * m && k < 18; ++k) out18[18 * i + k] = o[k];
  }
}
/* polygoncutting.hh:121-137 polygonArea */
double orc_polygon_area(const double* pts, int n) {
  if (n < 3) return 0.0;
  double area = 0.0;
  for (int i = 0; i < n; ++i) {
    const int j = (i + 1) % n;
    area += pts[2 * i] * pts[2 * j + 1];
    area -= pts[2 * j] * pts[2 * i + 1];
  }
  return 0.5 * area;
}
/* dune/mmesh/grid/cachingentity.hh:154-167 intersectionVolume: subject = cached vertex_ (no corner recompute),
   clip = TDS-order points; clip epsilon fixed at the reference's 1e-14 */
double orc_intersection_volume(const double* old6, const double* new6) {
  double o[2 * MAXPOLY];
  int n = orc_sutherland_hodgman(old6, new6, 1e-14, o);
  return fabs(orc_polygon_area(o, n));
}

typedef struct { double c[3][2]; double e[3][2]; } cut_in;

/* dune/mmesh/grid/cutsettriangulation.hh:31-62.  old6 = CachingEntity::vertex_ (cachingentity.hh:83-89),
   corners re-derived through AffineGeometry::corner (:40); new6 = host->vertex(i)->point() TDS order (:42). */
static int cutpoly(const double* old6, const double* new6, const orc_params* prm, double pts[MAXPOLY][2],
                   double c_unswapped[3][2], double origin[2]) {
  double c[3][2], e[3][2];
  double j00 = old6[2] - old6[0], j01 = old6[3] - old6[1], j10 = old6[4] - old6[0], j11 = old6[5] - old6[1];
  c[0][0] = old6[0]; c[0][1] = old6[1];
  c[1][0] = (old6[0] + 1.0 * j00) + 0.0 * j10; c[1][1] = (old6[1] + 1.0 * j01) + 0.0 * j11;
  c[2][0] = (old6[0] + 0.0 * j00) + 1.0 * j10; c[2][1] = (old6[1] + 0.0 * j01) + 1.0 * j11;
  for (int i = 0; i < 3; ++i) { e[i][0] = new6[2 * i]; e[i][1] = new6[2 * i + 1]; }
  origin[0] = origin[1] = 0.0;
  if (prm->clip_translate) { /* scale-aware variant (NOT reference arithmetic): work relative to e[0] */
    origin[0] = e[0][0]; origin[1] = e[0][1];
    for (int i = 0; i < 3; ++i) {
      c[i][0] -= origin[0]; c[i][1] -= origin[1];
      e[i][0] -= origin[0]; e[i][1] -= origin[1];
    }
  }
  if (c_unswapped) memcpy(c_unswapped, c, sizeof c);
  /* :46-49 orientation check truncated to int (reference quirk) */
  double od = (c[1][1] - c[0][1]) * (c[2][0] - c[1][0]) - (c[1][0] - c[0][0]) * (c[2][1] - c[1][1]);
  int o = (fabs(od) < 2147483648.0) ? (int)od : 0;
  if (o > 0) { double t0 = c[1][0], t1 = c[1][1]; c[1][0] = c[2][0]; c[1][1] = c[2][1]; c[2][0] = t0; c[2][1] = t1; }
  int n = sh_clip(c, e, prm->clip_eps, pts);
  return n < 3 ? 0 : n;
}
static int cutset(const double* old6, const double* new6, const orc_params* prm, double tris[][6],
                  double c_unswapped[3][2], double origin[2]) {
  double pts[MAXPOLY][2];
  int n = cutpoly(old6, new6, prm, pts, c_unswapped, origin);
  if (n < 3) return 0;
  int nt = 0;
  for (int i = 1; i < n - 1; ++i) {
    double a00 = pts[i][0] - pts[0][0], a01 = pts[i][1] - pts[0][1];
    double a10 = pts[i + 1][0] - pts[0][0], a11 = pts[i + 1][1] - pts[0][1];
    double vol = fabs(a00 * a11 - a10 * a01) * 0.5;
    if (vol > prm->drop_area) {
      tris[nt][0] = pts[0][0]; tris[nt][1] = pts[0][1];
      tris[nt][2] = pts[i][0]; tris[nt][3] = pts[i][1];
      tris[nt][4] = pts[i + 1][0]; tris[nt][5] = pts[i + 1][1];
      nt++;
    }
  }
  return nt;
}
int orc_cutset(const double* old6, const double* new6, const orc_params* prm, double* tris) {
  double t[MAXPOLY][6], cu[3][2], org[2];
  int n = cutset(old6, new6, prm, t, cu, org);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 6; ++k) tris[6 * i + k] = t[i][k] + ((k & 1) ? org[1] : org[0]);
  return n;
}

/* AffineGeometry::local (dune-geometry, un-vendored): J^{-T} (x - origin), 2x2 inverse via detInv */
typedef struct { double o[2]; double i00, i01, i10, i11; double vol; } aff2;
static inline void aff2_init(aff2* g, const double c0[2], const double c1[2], const double c2[2]) {
  double a00 = c1[0] - c0[0], a01 = c1[1] - c0[1], a10 = c2[0] - c0[0], a11 = c2[1] - c0[1];
  double det = a00 * a11 - a10 * a01;
  double detInv = 1.0 / det;
  g->o[0] = c0[0]; g->o[1] = c0[1];
  g->i00 = a11 * detInv; g->i10 = -a10 * detInv; g->i01 = -a01 * detInv; g->i11 = a00 * detInv;
  g->vol = fabs(det) * 0.5;
}
static inline void aff2_local(const aff2* g, const double x[2], double* xi, double* eta) {
  double d0 = x[0] - g->o[0], d1 = x[1] - g->o[1];
  double l0 = 0.0, l1 = 0.0;
  l0 += d0 * g->i00; l1 += d0 * g->i01;   /* mtv: y[j] += x[i] * A[i][j] */
  l0 += d1 * g->i10; l1 += d1 * g->i11;
  *xi = l0; *eta = l1;
}

/* mmesh.hh:1138-1165 adapt(handle) loop.  Transfer definitions (dune-fem is un-vendored):
   P0 : u_new[K]  (=|+=) (|T| / |K|) * u_old[K']   per cut triangle, in the reference's visiting order
        (FV restrictLocal with weight |son|/|father|).
   P1 : local L2 projection onto the nodal basis of the id-ordered corners of K.  For a cut triangle T with
        vertices v_j the product of two linear functions is integrated exactly by
            int_T u phi = |T| * (1/12) * ( (sum_j u_j)(sum_j phi_j) + sum_j u_j phi_j ),   (1/12, 1/3: rounded constants)
        u_j = u_old(v_j) = u0 + g.(v_j - o) with g the gradient of the affine old function, phi_j = barycentric
        coordinates of v_j in K (J^{-T}(v_j - origin), cf. AffineGeometry::local).  phi_0 = 1 - phi_1 - phi_2 is eliminated through int_T u = |T|/3 sum_j u_j.
        Each (old,new) pair accumulates its own partial sums; pairs are then added in CSR order. */
void orc_project(const double* xy, const int32_t* cells, const uint8_t* perm, const int64_t* new_off,
                 const int32_t* new_cell, int64_t nnew, const double* old_xy, const int32_t* old_idx,
                 const double* dofs_old, double* dofs_new, int ncomp, const orc_params* prm, double* mass_new,
                 double* covered_area, int64_t* n_pieces) {
  long double mass = 0.0L, cov = 0.0L; int64_t npc = 0;
#pragma omp parallel for num_threads(g_threads) reduction(+ : mass, cov, npc) schedule(static)
  for (int64_t i = 0; i < nnew; ++i) {
    const int32_t K = new_cell[i];
    double e6[6], k[3][2];
    for (int t = 0; t < 3; ++t) { e6[2 * t] = xy[2 * (int64_t)cells[3 * K + t]]; e6[2 * t + 1] = xy[2 * (int64_t)cells[3 * K + t] + 1]; }
    corners2d(xy, cells + 3 * K, perm[K], k);
    double acc0 = 0.0; int initialize = 1;           /* P0 */
    double R1 = 0.0, R2 = 0.0, U = 0.0;              /* P1 totals over pairs */
    double volK;
    { aff2 g0; aff2_init(&g0, k[0], k[1], k[2]); volK = g0.vol; }       /* element.geometry().volume() */
    for (int64_t p = new_off[i]; p < new_off[i + 1]; ++p) {
      double cu[3][2], org[2], pts[MAXPOLY][2];
      int n = cutpoly(old_xy + 6 * p, e6, prm, pts, cu, org);
      if (n < 3) continue;
      const double* uo = dofs_old + (int64_t)ncomp * old_idx[p];
      double uj[MAXPOLY], f1[MAXPOLY], f2[MAXPOLY];
      if (ncomp == 3) {
        aff2 gK, gO; double kk[3][2];
        for (int t = 0; t < 3; ++t) { kk[t][0] = k[t][0] - org[0]; kk[t][1] = k[t][1] - org[1]; }
        aff2_init(&gK, kk[0], kk[1], kk[2]);
        aff2_init(&gO, cu[0], cu[1], cu[2]);
        /* u_old is affine: u(x) = u0 + g . (x - o), g = J^{-1} (u1 - u0, u2 - u0) of the cached old cell */
        const double du1 = uo[1] - uo[0], du2 = uo[2] - uo[0];
        const double gx = du1 * gO.i00 + du2 * gO.i01, gy = du1 * gO.i10 + du2 * gO.i11;
        for (int j = 0; j < n; ++j) {
          const double d0 = pts[j][0] - gO.o[0], d1 = pts[j][1] - gO.o[1];
          uj[j] = uo[0] + (d0 * gx + d1 * gy);
          const double e0 = pts[j][0] - gK.o[0], e1 = pts[j][1] - gK.o[1];
          f1[j] = e0 * gK.i00 + e1 * gK.i10;                 /* barycentric coordinates of v_j in K (phi_1, phi_2) */
          f2[j] = e0 * gK.i01 + e1 * gK.i11;
        }
      }
      double r1 = 0.0, r2 = 0.0, u = 0.0;            /* per-pair partial sums */
      for (int t = 1; t < n - 1; ++t) {
        double a00 = pts[t][0] - pts[0][0], a01 = pts[t][1] - pts[0][1];
        double a10 = pts[t + 1][0] - pts[0][0], a11 = pts[t + 1][1] - pts[0][1];
        double aT = fabs(a00 * a11 - a10 * a01) * 0.5;
        if (!(aT > prm->drop_area)) continue;                          /* cutsettriangulation.hh:60 */
        cov += aT; npc++;
        if (ncomp == 1) {
          double w = aT / volK;
          if (initialize) acc0 = w * uo[0]; else acc0 += w * uo[0];
          initialize = 0;
        } else {
          const double su = uj[0] + uj[t] + uj[t + 1];
          const double s1 = f1[0] + f1[t] + f1[t + 1], s2 = f2[0] + f2[t] + f2[t + 1];
          const double d1 = uj[0] * f1[0] + uj[t] * f1[t] + uj[t + 1] * f1[t + 1];
          const double d2 = uj[0] * f2[0] + uj[t] * f2[t] + uj[t + 1] * f2[t + 1];
          const double w12 = aT * (1.0 / 12.0);
          r1 += w12 * (su * s1 + d1);
          r2 += w12 * (su * s2 + d2);
          u += (aT * (1.0 / 3.0)) * su;
        }
      }
      R1 += r1; R2 += r2; U += u;
    }
    if (ncomp == 1) {
      dofs_new[K] = acc0;
      mass += (long double)volK * acc0;
    } else {
      const double R0 = U - R1 - R2;
      const double f = 3.0 / volK;
      double u0 = f * (4.0 * R0 - U), u1 = f * (4.0 * R1 - U), u2 = f * (4.0 * R2 - U);
      dofs_new[3 * (int64_t)K] = u0; dofs_new[3 * (int64_t)K + 1] = u1; dofs_new[3 * (int64_t)K + 2] = u2;
      mass += (long double)volK * ((u0 + u1 + u2) / 3.0);
    }
  }
  if (mass_new) *mass_new = (double)mass;
  if (covered_area) *covered_area = (double)cov;
  if (n_pieces) *n_pieces = npc;
}

/* dune/mmesh/interface/grid.hh:388-415 with FV (P0) data: father = longer one.
   old longer  -> prolongLocal(father=old, son=new): new = old value (copy)
   else        -> restrictLocal(father=new, son=old, init): new (=|+=) (|old|/|new|) * old */
void orc_project_interface_p0(const double* new_len, const int64_t* new_off, int64_t nnew, const double* old_len,
                              const int32_t* old_idx, const double* dofs_old, double* dofs_new) {
  for (int64_t i = 0; i < nnew; ++i) {
    int initialize = 1; double acc = 0.0;
    for (int64_t p = new_off[i]; p < new_off[i + 1]; ++p) {
      double u = dofs_old[old_idx[p]];
      if (old_len[p] > new_len[i]) acc = u;
      else { double w = old_len[p] / new_len[i]; if (initialize) acc = w * u; else acc += w * u; }
      initialize = 0;
    }
    dofs_new[i] = acc;
  }
}

/* mmesh.hh:1194-1218 + 1919-1941: sequential, order dependent -- restated statement by statement */
int64_t orc_component_numbers(int64_t nz, const int64_t* zone_off, const int32_t* zone_cells, int64_t component_count,
                              int32_t* zone_number, int32_t* cell_number) {
  int markAgain = 1;
  while (markAgain) {                                                      /* :1204 */
    markAgain = 0;
    for (int64_t k = 0; k < nz; ++k) {                                     /* :1212-1226: insert_ then remove_, in order */
      int64_t componentNumber = 0;                                         /* getComponentNumber_, :1922 */
      for (int64_t e = zone_off[k]; e < zone_off[k + 1]; ++e) {
        const int32_t num = cell_number[zone_cells[e]];
        if (num > 0) {                                                     /* :1926 */
          if (componentNumber != 0 && num != componentNumber) {            /* :1928-1929 */
            markAgain = 1;
            componentNumber = num < componentNumber ? num : componentNumber;
          } else
            componentNumber = num;
        }
      }
      if (componentNumber == 0) componentNumber = component_count++;       /* :1938 */
      for (int64_t e = zone_off[k]; e < zone_off[k + 1]; ++e) cell_number[zone_cells[e]] = (int32_t)componentNumber;   /* :1867-1870 */
      zone_number[k] = (int32_t)componentNumber;
    }
  }
  return component_count;
}

/* mmesh.hh:920-1088 ensureInterfaceMovement (dim == 2), restated on index arrays; see the header for the conventions */
int orc_ensure_interface_movement_2d(const double* x, int64_t nv, const int32_t* cells, const uint8_t* perm, int64_t nc,
                                     const uint8_t* vflags, const int32_t* facets, int64_t ns, const int32_t* iverts, int64_t niv,
                                     const double* ishifts, const int32_t* level, uint8_t* removed_mask, int32_t* removed,
                                     int64_t* n_removed, int32_t* ins_edge, double* ins_point, int32_t* ins_v0, int64_t* ins_level,
                                     int32_t* ins_crossing, uint8_t* ins_comp_is_edge, int64_t* n_inserted) {
  int change = 0;
  *n_removed = 0; *n_inserted = 0;
  double* xt = (double*)malloc((size_t)nv * 2 * sizeof(double));
  double* sh = (double*)malloc((size_t)(niv ? niv : 1) * 2 * sizeof(double));
  int32_t* ivof = (int32_t*)malloc((size_t)nv * sizeof(int32_t));      /* interface leaf index of a bulk vertex */
  int32_t* nincident = (int32_t*)calloc((size_t)nv, sizeof(int32_t));
  memcpy(xt, x, (size_t)nv * 2 * sizeof(double));
  memcpy(sh, ishifts, (size_t)niv * 2 * sizeof(double));
  for (int64_t v = 0; v < nv; ++v) ivof[v] = -1;
  for (int64_t i = 0; i < niv; ++i) ivof[iverts[i]] = (int32_t)i;
  for (int64_t f = 0; f < 2 * ns; ++f) nincident[facets[f]]++;
  int rc = 0;
#define SIGNED_AREA(c) (((xt[2 * cells[3 * (c) + 1]] - xt[2 * cells[3 * (c)]]) * (xt[2 * cells[3 * (c) + 2] + 1] - xt[2 * cells[3 * (c)] + 1]) - \
                         (xt[2 * cells[3 * (c) + 2]] - xt[2 * cells[3 * (c)]]) * (xt[2 * cells[3 * (c) + 1] + 1] - xt[2 * cells[3 * (c)] + 1])) / 2)
  for (int64_t c = 0; c < nc && rc == 0; ++c)                           /* :926-930 check if grid is still valid */
    if (SIGNED_AREA(c) <= 0.0) rc = -1;
  if (rc == 0) {
    for (int64_t i = 0; i < niv; ++i) {                                  /* :933 moveInterface(shifts) */
      xt[2 * iverts[i]] = x[2 * iverts[i]] + sh[2 * i]; xt[2 * iverts[i] + 1] = x[2 * iverts[i] + 1] + sh[2 * i + 1];
    }
    static const int EV[3][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 } };         /* DUNE edge numbering of the id-sorted corners */
    for (int64_t c = 0; c < nc && rc == 0; ++c) {                        /* :935 */
      if (!(SIGNED_AREA(c) <= 0.0)) continue;
      int32_t k[3];
      for (int j = 0; j < 3; ++j) k[j] = cells[3 * c + ((perm[c] >> (2 * j)) & 3)];    /* subEntity<dim>(j): id order */
      int allInterface = 1;
      for (int j = 0; j < 3; ++j) {                                      /* :945-962 */
        const int32_t v = k[j];
        if (!(vflags[v] & 1)) {
          if (vflags[v] & 2) continue;                                   /* boundaryFlag(v) == 1 */
          if (!removed_mask[v]) { removed_mask[v] = 1; removed[(*n_removed)++] = v; change = 1; }
          allInterface = 0;
        }
      }
      if (!allInterface) continue;
      int ie = -1;
      for (int j = 0; j < 3 && rc == 0; ++j) {                           /* :966-980 isInterface(edge) */
        const int32_t a = k[EV[j][0]], b = k[EV[j][1]];
        int isif = 0;
        if ((vflags[a] & 1) && (vflags[b] & 1))
          for (int64_t f = 0; f < ns && !isif; ++f)
            isif = (facets[2 * f] == a && facets[2 * f + 1] == b) || (facets[2 * f] == b && facets[2 * f + 1] == a);
        if (isif) { if (ie >= 0) rc = -2; else ie = j; }
      }
      if (rc) break;
      if (ie < 0) {                                                      /* :983-1006 */
        int allNonRemovable = 1;
        for (int j = 0; j < 3; ++j)
          if (nincident[k[j]] == 2) {                                    /* isRemoveable, longestedgerefinement.hh:160-168 */
            allNonRemovable = 0;
            if (!removed_mask[k[j]]) { removed_mask[k[j]] = 1; removed[(*n_removed)++] = k[j]; change = 1; }
          }
        if (allNonRemovable) rc = -3;
        continue;
      }
      const int32_t ea = k[EV[ie][0]], eb = k[EV[ie][1]], third = k[3 - EV[ie][0] - EV[ie][1]];
      int32_t crossing = -1;
      for (int64_t f = 0; f < ns && rc == 0; ++f)                        /* :1021-1036 incidentInterfaceElements(iThirdVertex) */
        if (facets[2 * f] == third || facets[2 * f + 1] == third) { if (crossing >= 0) rc = -4; else crossing = (int32_t)f; }
      if (rc) break;
      /* :1039-1046: corners through AffineGeometry::corner (c1 = c0 + 1.0 * (v1 - c0)) */
      const double c0[2] = { xt[2 * ea], xt[2 * ea + 1] };
      const double c1[2] = { c0[0] + 1.0 * (xt[2 * eb] - c0[0]), c0[1] + 1.0 * (xt[2 * eb + 1] - c0[1]) };
      const int32_t fa = facets[2 * crossing], fb = facets[2 * crossing + 1];
      const double e0[2] = { xt[2 * fa], xt[2 * fa + 1] };
      const double e1[2] = { e0[0] + 1.0 * (xt[2 * fb] - e0[0]), e0[1] + 1.0 * (xt[2 * fb + 1] - e0[1]) };
      double xi[2];
      line_isect(c0, c1, e0, e1, xi);                                    /* polygoncutting.hh:100-114 */
      double dot = 0.0; dot += (c0[0] - xi[0]) * (c1[0] - xi[0]); dot += (c0[1] - xi[1]) * (c1[1] - xi[1]);
      if (dot > 0.) continue;                                            /* :1049 */
      const int64_t n = (*n_inserted)++;
      ins_edge[2 * n] = ea; ins_edge[2 * n + 1] = eb; ins_point[2 * n] = xi[0]; ins_point[2 * n + 1] = xi[1];
      ins_v0[n] = third; ins_level[n] = (level ? (int64_t)level[third] : 0) + 1; ins_crossing[n] = crossing;
      {
        double la = 0.0, lb = 0.0, t;
        t = c1[0] - c0[0]; la += t * t; t = c1[1] - c0[1]; la += t * t;
        t = e1[0] - e0[0]; lb += t * t; t = e1[1] - e0[1]; lb += t * t;
        ins_comp_is_edge[n] = sqrt(la) > sqrt(lb);                       /* :1060-1063 the edge with the larger volume */
      }
      change = 1;
      const int32_t idx = ivof[third];                                   /* :1072-1076 move third vertex back */
      xt[2 * third] = xt[2 * third] - sh[2 * idx]; xt[2 * third + 1] = xt[2 * third + 1] - sh[2 * idx + 1];
      sh[2 * idx] = 0.0; sh[2 * idx + 1] = 0.0;
    }
  }
#undef SIGNED_AREA
  free(xt); free(sh); free(ivof); free(nincident);
  return rc ? rc : change;
}

/* interface/grid.hh:388-415 with DG-P1 data; see the header for the definitions */
static inline double seg_len(const double* c) {                /* AffineGeometry<1,2>::volume(): sqrt(J J^T) */
  double dx = c[2] - c[0], dy = c[3] - c[1], q = 0.0; q += dx * dx; q += dy * dy; return sqrt(q);
}
static inline double seg_local(const double* c, double px, double py) {     /* father.geometry().local(p) */
  double dx = c[2] - c[0], dy = c[3] - c[1];
  double s = 0.0; s += (px - c[0]) * dx; s += (py - c[1]) * dy;
  double n2 = 0.0; n2 += dx * dx; n2 += dy * dy;
  return s / n2;
}
void orc_project_interface_p1(const double* new_xy, const int64_t* new_off, int64_t nnew, const double* old_xy, const int32_t* old_idx,
                              const double* dofs_old, double* dofs_new) {
  for (int64_t i = 0; i < nnew; ++i) {
    const double* nw = new_xy + 4 * i;
    const double ln = seg_len(nw);
    double a0 = 0.0, a1 = 0.0; int initialize = 1;
    for (int64_t p = new_off[i]; p < new_off[i + 1]; ++p) {
      const double* od = old_xy + 4 * p;
      const double u0 = dofs_old[2 * (int64_t)old_idx[p]], u1 = dofs_old[2 * (int64_t)old_idx[p] + 1];
      const double lo = seg_len(od);
      if (lo > ln) {                                             /* prolongLocal(father = old, son = new, true) */
        const double t0 = seg_local(od, nw[0], nw[1]), t1 = seg_local(od, nw[2], nw[3]);
        a0 = (1.0 - t0) * u0 + t0 * u1;
        a1 = (1.0 - t1) * u0 + t1 * u1;
      } else {                                                   /* restrictLocal(father = new, son = old, initialize) */
        const double ta = seg_local(nw, od[0], od[1]), tb = seg_local(nw, od[2], od[3]);
        /* int_son u phi_i with the exact vertex rule len/6 (2 ua pa + ua pb + ub pa + 2 ub pb), phi_0 = 1 - t, phi_1 = t */
        const double pa0 = 1.0 - ta, pb0 = 1.0 - tb, pa1 = ta, pb1 = tb;
        const double w = lo * (1.0 / 6.0);
        const double r0 = w * (((2.0 * u0) * pa0 + u0 * pb0) + (u1 * pa0 + (2.0 * u1) * pb0));
        const double r1 = w * (((2.0 * u0) * pa1 + u0 * pb1) + (u1 * pa1 + (2.0 * u1) * pb1));
        /* M_f^-1 = (2 / len_f) [2 -1; -1 2] */
        const double f = 2.0 / ln;
        const double c0 = f * (2.0 * r0 - r1), c1 = f * (2.0 * r1 - r0);
        if (initialize) { a0 = c0; a1 = c1; } else { a0 += c0; a1 += c1; }
      }
      initialize = 0;
    }
    dofs_new[2 * i] = a0; dofs_new[2 * i + 1] = a1;
  }
}

/* ratioindicator.hh:62-91 in 3-D */
void orc_indicator_init_3d(const double* xyz, const int32_t* cells, int64_t nc, const int32_t* tris, int64_t ns, orc_params* prm) {
  const double K = 2.0, k = K / (2.0 * 4.0);
  double maxH = 0.0, minH = 1e100;
  for (int64_t s = 0; s < ns; ++s)
    for (int e = 0; e < 3; ++e) {
      double h = edge_len3d(xyz + 3 * (int64_t)tris[3 * s + e], xyz + 3 * (int64_t)tris[3 * s + (e + 1) % 3]);
      maxH = (maxH < h) ? h : maxH; minH = (h < minH) ? h : minH;
    }
  maxH = K * maxH; minH = k * minH;
  double maxh = 0.0, minh = 1e100;
  for (int64_t c = 0; c < nc; ++c)
    for (int a = 0; a < 4; ++a) for (int b = a + 1; b < 4; ++b) {
      double h = edge_len3d(xyz + 3 * (int64_t)cells[4 * c + a], xyz + 3 * (int64_t)cells[4 * c + b]);
      maxh = (maxh < h) ? h : maxh; minh = (h < minh) ? h : minh;
    }
  if (ns == 0) { maxH = K * maxh; minH = k * minh; }
  prm->maxH = maxH; prm->minH = minH; prm->factor = maxh / minh;
}
/* mmesh.hh:745-748 + ratioindicator.hh:106-160 for the triangles of a 3-D interface */
void orc_mark_interface_3d(const double* xyz, const int32_t* tris, int64_t ns, const orc_params* prm, double maxDist, int8_t* mark) {
  for (int64_t s = 0; s < ns; ++s) {
    double dist = (0.0 < maxDist) ? 0.0 : maxDist;                 /* std::min(maxDist_, distance_(ielement) == 0) */
    const double l = dist / maxDist;
    const double maxH = (1. - l) * prm->maxH + l * prm->factor * prm->maxH;
    int refine = 0;
    for (int e = 0; e < 3; ++e) {
      double len = edge_len3d(xyz + 3 * (int64_t)tris[3 * s + e], xyz + 3 * (int64_t)tris[3 * s + (e + 1) % 3]);
      if (len > maxH) refine++;
    }
    mark[s] = (int8_t)(refine > 0 ? 1 : 0);                        /* coarse > 0 returns -1 only if dim != 3 */
  }
}

/* ---- Delaunay restoration by edge flips (sequential Lawson; see mmesh_oracle.h) -------------------- */
static int slot_of(const int32_t* nb, int64_t cell, int32_t target) {
  for (int k = 0; k < 3; ++k) if (nb[3 * cell + k] == target) return k;
  return -1;
}
int64_t orc_restore_delaunay_2d(const double* xy, int32_t* cells, int32_t* nb, int64_t nc, const int32_t* segs, int64_t ns,
                                int64_t max_flips) {
  keyset constrained;
  if (!keyset_init(&constrained, ns)) return -1;
  for (int64_t s = 0; s < ns; ++s) keyset_insert(&constrained, edge_key(segs[2 * s], segs[2 * s + 1]));
  /* stack of (cell, slot) edges still to examine: all of them at the start */
  int64_t cap = 3 * nc + 16, top = 0, flips = 0;
  int64_t* stack = (int64_t*)malloc((size_t)cap * sizeof(int64_t));
  if (!stack) { free(constrained.slot); return -1; }
  for (int64_t c = nc - 1; c >= 0; --c) for (int k = 2; k >= 0; --k) stack[top++] = 3 * c + k;
  while (top > 0) {
    const int64_t e = stack[--top];
    const int64_t f = e / 3; const int i = (int)(e % 3);
    const int32_t n = nb[3 * f + i];
    if (n < 0) continue;
    const int ni = slot_of(nb, n, (int32_t)f);
    if (ni < 0) { flips = -1; break; }
    const int ccw_i = (i + 1) % 3, cw_i = (i + 2) % 3, ccw_ni = (ni + 1) % 3, cw_ni = (ni + 2) % 3;
    const int32_t apex = cells[3 * f + i], a = cells[3 * f + ccw_i], b = cells[3 * f + cw_i], opp = cells[3 * (int64_t)n + ni];
    if (keyset_has(&constrained, edge_key(a, b))) continue;
    if (orc_incircle(xy + 2 * (int64_t)apex, xy + 2 * (int64_t)a, xy + 2 * (int64_t)b, xy + 2 * (int64_t)opp) <= 0) continue;
    if (max_flips >= 0 && flips >= max_flips) { flips = -1; break; }
    /* Triangulation_data_structure_2.h:885-915 */
    const int32_t tr = nb[3 * f + ccw_i], bl = nb[3 * (int64_t)n + ccw_ni];
    const int tri = tr >= 0 ? slot_of(nb, tr, (int32_t)f) : -1, bli = bl >= 0 ? slot_of(nb, bl, n) : -1;
    cells[3 * f + cw_i] = opp;
    cells[3 * (int64_t)n + cw_ni] = apex;
    nb[3 * f + i] = bl;            if (bli >= 0) nb[3 * (int64_t)bl + bli] = (int32_t)f;
    nb[3 * f + ccw_i] = n;         nb[3 * (int64_t)n + ccw_ni] = (int32_t)f;
    nb[3 * (int64_t)n + ni] = tr;  if (tri >= 0) nb[3 * (int64_t)tr + tri] = n;
    ++flips;
    /* the four outer edges of the new pair (propagating_flip re-examines the two facing the apex; a sweep over all edges
       needs all four) */
    if (top + 4 > cap) {
      cap = 2 * cap + 16;
      int64_t* g = (int64_t*)realloc(stack, (size_t)cap * sizeof(int64_t));
      if (!g) { flips = -1; break; }
      stack = g;
    }
    stack[top++] = 3 * f + i; stack[top++] = 3 * f + cw_i; stack[top++] = 3 * (int64_t)n + ni; stack[top++] = 3 * (int64_t)n + cw_ni;
  }
  free(stack); free(constrained.slot);
  return flips;
}
