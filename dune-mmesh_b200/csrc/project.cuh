// project.cuh -- DG-P1 conservative projection old -> new cells, organised as a tile of pair slots per CTA.
//
// Reference: driver loop dune/mmesh/grid/mmesh.hh:1143-1161; CutSetTriangulation dune/mmesh/grid/cutsettriangulation.hh:31-62;
// Sutherland-Hodgman + lineIntersectionPoint dune/mmesh/grid/polygoncutting.hh:17-114; local coordinates
// dune/mmesh/grid/entity.hh:573-594.  The DG-P1 transfer arithmetic is the oracle's definition (dune-fem is un-vendored,
// SURVEY A8): per (old, new) pair partial sums r1, r2, u over the fan triangles, added per new cell in CSR order.
//
// One CTA owns NT/4 consecutive new cells of the pair CSR and walks their pairs NT at a time:
//   A  cell threads   new-cell corners -> the three clip-edge constants, inverse affine map, volume       (once per cell)
//   B  pair threads   old corners -> subject triangle -> three clip stages in shared memory [slot][column]
//   C  compaction     columns holding a polygon (>= 3 points) are appended to a dense list (warp-aggregated)
//   D  dense threads  fan triangulation + quadrature of one non-empty polygon -> per-pair partial sums
//   E  cell threads   partial sums of the cell's pairs added in CSR order; at the end the local mass-matrix solve
// When every cell of the tile has exactly four pairs the pair -> column map is transposed (column = q * (NT/4 + 2) + c
// for the q-th pair of cell c), so that a warp works on the same q of 32 cells: the lanes of a warp then see polygons
// of the same kind (e.g. all "cell against its own old copy"), instead of one of each kind per four lanes.
//
// Every function below is host-compilable: tests/native/host_project.cc runs the phases thread by thread on the CPU
// against the oracle (no GPU needed).  Compile with -fmad=false (device) / -ffp-contract=off (host); the quadrature uses
// explicit fma() only when FMA = true (DOF tolerance 1e-12; areas, the drop test and the piece count never use it).
#pragma once
#include <math.h>
#include <stdint.h>
#include "clip.cuh"
#include "geom.cuh"

namespace mmb {

struct Aff2 { double ox, oy, i00, i01, i10, i11, vol; };
// AffineGeometry of a triangle (dune-geometry): jacobian rows c1 - c0, c2 - c0, inverse through detInv, volume |det| / 2
MMB_HD Aff2 aff2_make(double c0x, double c0y, double c1x, double c1y, double c2x, double c2y) {
  const double a00 = c1x - c0x, a01 = c1y - c0y, a10 = c2x - c0x, a11 = c2y - c0y;
  const double det = a00 * a11 - a10 * a01;
  const double detInv = 1.0 / det;
  Aff2 g;
  g.ox = c0x; g.oy = c0y;
  g.i00 = a11 * detInv; g.i10 = -a10 * detInv; g.i01 = -a01 * detInv; g.i11 = a00 * detInv;
  g.vol = fabs(det) * 0.5;
  return g;
}

constexpr int MAXPOLY = 12;   // generic path: rigorous bound for triangle x triangle with the epsilon cases is 9
struct Poly { double x[MAXPOLY], y[MAXPOLY]; int n; };

// Sutherland-Hodgman, polygoncutting.hh:17-96, in the reference's own loop structure (generic path: intersection
// volumes, and pairs whose polygon does not fit the tile's slots).
MMB_HD void sh_clip(const double cx[3], const double cy[3], const double ex[3], const double ey[3], double eps, Poly& A,
                    Poly& B, Poly*& result) {
  Poly* in = &B; Poly* out = &A;
  out->n = 3;
  for (int i = 0; i < 3; ++i) { out->x[i] = cx[i]; out->y[i] = cy[i]; }
  for (int ci = 0; ci < 3; ++ci) {
    const int cn = ci == 2 ? 0 : ci + 1;
    const double q1x = ex[ci], q1y = ey[ci], q2x = ex[cn], q2y = ey[cn];
    const double dQx = q2x - q1x, dQy = q2y - q1y;
    const double q2xq1 = q1y * q2x - q1x * q2y;
    const double ndQx = -dQx;
    const double iQx = q1x - q2x, iQy = q1y - q2y;          // deltaQ of lineIntersectionPoint
    const double fq = q1x * q2y - q1y * q2x;
    { Poly* t = in; in = out; out = t; }
    const int n = in->n;
    int m = 0;
    if (n > 0) {
      const double f0 = ndQx * in->y[0] + dQy * in->x[0] + q2xq1;
      double first = f0;
      double ax = in->x[0], ay = in->y[0];
      for (int ii = 0; ii < n; ++ii) {
        const int i2 = ii + 1 == n ? 0 : ii + 1;
        const double bx = in->x[i2], by = in->y[i2];
        const double second = i2 == 0 ? f0 : ndQx * by + dQy * bx + q2xq1;
        const bool c1 = first > eps && second < -eps;
        const bool c2 = first < -eps && second > eps;
        if (c1 || c2) {
          double dPx = ax - bx, dPy = ay - by;
          const double scale = dPx * iQy - iQx * dPy;
          const double fp = ax * by - ay * bx;
          dPx *= fq; dPy *= fq;
          out->x[m] = (iQx * fp - dPx) / scale; out->y[m] = (iQy * fp - dPy) / scale; ++m;
          if (c1) { out->x[m] = bx; out->y[m] = by; ++m; }
        } else if (!(first >= -eps && second > eps)) {
          out->x[m] = bx; out->y[m] = by; ++m;
        }
        first = second; ax = bx; ay = by;
        if (m > MAXPOLY - 2) { m = 0; break; }
      }
    }
    out->n = m;
  }
  result = out;
}

// ---- the tile ------------------------------------------------------------------------------------------------------
template <int NT>
struct ProjTile {
  static constexpr int NCELL = NT / 4;          // new cells per CTA
  static constexpr int PADQ = NCELL + 2;        // transposed map: column(q, c) = q * PADQ + c (the +2 keeps the transposing stores conflict free)
  static constexpr int NCOL = 4 * PADQ;
  double2 P[5][NCOL];      // subject triangle; output of clip edge 1 (<= 5 points); finally the per-pair partial sums
  double2 Q[6][NCOL];      // outputs of clip edges 0 (<= 4) and 2 (<= 6)
  double EC[9][NCELL];     // per new cell and clip edge st: dQx, dQy, q2xq1 at rows 3 st .. 3 st + 2
  double2 G[3][NCELL];     // inverse affine map of the new cell: origin, (i00, i01), (i10, i11)
  double2 org[NCELL];      // clip_translate origin (0, 0 otherwise)
  double vol[NCELL];       // element.geometry().volume() of the new cell
  double acc[4][NCELL];    // R1, R2, U, covered area: sums over the cell's pairs in CSR order
  int32_t npc[NCELL];      // kept cut triangles of the cell
  int32_t K[NCELL];        // the new cell's index in the mesh
  int32_t poff[NCELL + 1]; // first pair of each cell relative to the tile's first pair
  uint8_t owner[NCOL];     // ragged tiles: cell of each pair slot
  uint8_t npts[NCOL];      // polygon size per column (0xFF: generic path)
  uint8_t npcs[NCOL];      // kept cut triangles per column
  uint8_t dense[NCOL];     // pair slots (j) that hold a polygon
  int32_t ndense;
  int32_t uniform;
};
constexpr uint8_t PROJ_HARD = 0xFF;

struct ProjArgs {
  const double2* xy; const int32_t *cv0, *cv1, *cv2; const uint8_t* perm;
  const long long* new_off; const int32_t* new_cell; int64_t i_begin, nnew;
  const double* old_xy; const int32_t* old_idx; const double* dofs_old; double* dofs_new;
  DevParams prm;
};

template <int NT> MMB_HD int proj_column(bool uniform, int j) {
  return uniform ? (j & 3) * ProjTile<NT>::PADQ + (j >> 2) : j;
}

MMB_HD double2 proj_ld2(const double2* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// A: cell c of the tile (new cell i of the CSR); returns its number of pairs
template <int NT>
MMB_HD int proj_cell_setup(ProjTile<NT>& T, const ProjArgs& a, int c, int64_t i, long long p_lo) {
  const int K = a.new_cell[i];
  const long long p0 = a.new_off[i], p1 = a.new_off[i + 1];
  const int v[3] = { a.cv0[K], a.cv1[K], a.cv2[K] };
  const unsigned pm = a.perm[K];
  double ex[3], ey[3], kx[3], ky[3];
  for (int t = 0; t < 3; ++t) { const double2 q = proj_ld2(a.xy + v[t]); ex[t] = q.x; ey[t] = q.y; }       // TDS order (cutsettriangulation.hh:42)
  for (int t = 0; t < 3; ++t) { const int s = (pm >> (2 * t)) & 3; kx[t] = s == 0 ? ex[0] : (s == 1 ? ex[1] : ex[2]); ky[t] = s == 0 ? ey[0] : (s == 1 ? ey[1] : ey[2]); }
  T.vol[c] = aff2_make(kx[0], ky[0], kx[1], ky[1], kx[2], ky[2]).vol;           // element.geometry().volume()
  double orgx = 0.0, orgy = 0.0;
  if (a.prm.clip_translate) {
    orgx = ex[0]; orgy = ey[0];
    for (int t = 0; t < 3; ++t) { ex[t] -= orgx; ey[t] -= orgy; }
  }
  const Aff2 gK = aff2_make(kx[0] - orgx, ky[0] - orgy, kx[1] - orgx, ky[1] - orgy, kx[2] - orgx, ky[2] - orgy);
  T.G[0][c] = make_double2(gK.ox, gK.oy); T.G[1][c] = make_double2(gK.i00, gK.i01); T.G[2][c] = make_double2(gK.i10, gK.i11);
  T.org[c] = make_double2(orgx, orgy);
  for (int st = 0; st < 3; ++st) {
    const int sn = st == 2 ? 0 : st + 1;
    clip_edge_constants(ex[st], ey[st], ex[sn], ey[sn], T.EC[3 * st][c], T.EC[3 * st + 1][c], T.EC[3 * st + 2][c]);
  }
  T.K[c] = K;
  T.poff[c] = (int32_t)(p0 - p_lo);
  T.poff[c + 1] = (int32_t)(p1 - p_lo);          // == the next cell's own entry; the last cell closes the table
  T.acc[0][c] = 0.0; T.acc[1][c] = 0.0; T.acc[2][c] = 0.0; T.acc[3][c] = 0.0; T.npc[c] = 0;
  return (int)(p1 - p0);
}
// ragged tiles: cell c claims its pair slots of the chunk [cl, cl + NT) (cl relative to the tile's first pair)
template <int NT>
MMB_HD void proj_paint(ProjTile<NT>& T, int c, int cl) {
  const int lo = T.poff[c] > cl ? T.poff[c] : cl, hi = T.poff[c + 1] < cl + NT ? T.poff[c + 1] : cl + NT;
  for (int p = lo; p < hi; ++p) T.owner[p - cl] = (uint8_t)c;
}
// the old cell of pair p as cutsettriangulation.hh:40 sees it: corners through AffineGeometry::corner, translated
MMB_HD void proj_old_corners(const double* old_xy, long long p, double2 org, double cx[3], double cy[3]) {
  const double2* o6 = reinterpret_cast<const double2*>(old_xy + 6 * p);
  const double2 o0 = proj_ld2(o6), o1 = proj_ld2(o6 + 1), o2 = proj_ld2(o6 + 2);
  const double j00 = o1.x - o0.x, j01 = o1.y - o0.y, j10 = o2.x - o0.x, j11 = o2.y - o0.y;
  cx[0] = o0.x; cy[0] = o0.y;
  cx[1] = (o0.x + 1.0 * j00) + 0.0 * j10; cy[1] = (o0.y + 1.0 * j01) + 0.0 * j11;
  cx[2] = (o0.x + 0.0 * j00) + 1.0 * j10; cy[2] = (o0.y + 0.0 * j01) + 1.0 * j11;
  for (int t = 0; t < 3; ++t) { cx[t] -= org.x; cy[t] -= org.y; }          // org = (0, 0) unless clip_translate: x - 0.0 == x
}
// B1: pair slot j of the chunk: subject triangle into column(j)
template <int NT>
MMB_HD void proj_load_pair(ProjTile<NT>& T, const ProjArgs& a, int j, long long p, bool uniform) {
  const int c = uniform ? (j >> 2) : T.owner[j];
  const int col = proj_column<NT>(uniform, j);
  double cx[3], cy[3];
  proj_old_corners(a.old_xy, p, T.org[c], cx, cy);
  // cutsettriangulation.hh:46-49: orientation value truncated to int
  const double od = (cy[1] - cy[0]) * (cx[2] - cx[1]) - (cx[1] - cx[0]) * (cy[2] - cy[1]);
  const int o = (fabs(od) < 2147483648.0) ? (int)od : 0;
  const bool sw = o > 0;
  T.P[0][col] = make_double2(cx[0], cy[0]);
  T.P[1][col] = sw ? make_double2(cx[2], cy[2]) : make_double2(cx[1], cy[1]);
  T.P[2][col] = sw ? make_double2(cx[1], cy[1]) : make_double2(cx[2], cy[2]);
}
// B2: the three clip stages on one column; returns true when the column holds something for phase D
template <int NT>
MMB_HD bool proj_clip_column(ProjTile<NT>& T, const ProjArgs& a, int col, int c) {
  auto ldP = [&](int k, double& x, double& y) { const double2 v = T.P[k][col]; x = v.x; y = v.y; };
  auto ldQ = [&](int k, double& x, double& y) { const double2 v = T.Q[k][col]; x = v.x; y = v.y; };
  auto stP = [&](int k, double x, double y) { T.P[k][col] = make_double2(x, y); };
  auto stQ = [&](int k, double x, double y) { T.Q[k][col] = make_double2(x, y); };
  const double eps = a.prm.clip_eps;
  int n = 3;
  n = clip_stage_core<3, 4>(ldP, n, T.EC[0][c], T.EC[1][c], T.EC[2][c], eps, stQ);
  if (n >= 0) n = clip_stage_core<4, 5>(ldQ, n, T.EC[3][c], T.EC[4][c], T.EC[5][c], eps, stP);
  if (n >= 0) n = clip_stage_core<5, 6>(ldP, n, T.EC[6][c], T.EC[7][c], T.EC[8][c], eps, stQ);
  const bool hard = n < 0;
  const bool keep = hard || n >= 3;                              // cutsettriangulation.hh:55 `if (points.size() < 3) return;`
  T.npts[col] = hard ? PROJ_HARD : (uint8_t)n;
  if (!keep) { T.P[0][col] = make_double2(0.0, 0.0); T.P[1][col] = make_double2(0.0, 0.0); T.npcs[col] = 0; }
  return keep;
}
// D: fan (points[0], points[j-1], points[j]) of one polygon (cutsettriangulation.hh:57-61) and the exact vertex rule
//    int_T u phi = |T|/12 ((sum u)(sum phi) + sum u phi) for the affine old function against the new cell's barycentrics
template <bool FMA, class Get>
MMB_HD void proj_fan(Get get, int nC, double u0, double gx, double gy, double gox, double goy, double2 g0, double2 g1,
                     double2 g2, double drop_area, double& r1, double& r2, double& u, double& cov, unsigned& npc) {
  auto eval = [&](const double2 P, double& uj, double& f1, double& f2) {
    const double d0 = P.x - gox, d1 = P.y - goy;
    const double e0 = P.x - g0.x, e1 = P.y - g0.y;
    if (FMA) {
      uj = fma(d0, gx, fma(d1, gy, u0));
      f1 = fma(e0, g1.x, e1 * g2.x);
      f2 = fma(e0, g1.y, e1 * g2.y);
    } else {
      uj = u0 + (d0 * gx + d1 * gy);                       // affine old function: u0 + grad . (x - o)
      f1 = e0 * g1.x + e1 * g2.x;                          // barycentric coordinates in the new cell
      f2 = e0 * g1.y + e1 * g2.y;
    }
  };
  const double2 F = get(0);
  double2 Pp = get(1);
  double fu, ff1, ff2, pu, pf1, pf2;
  eval(F, fu, ff1, ff2);
  eval(Pp, pu, pf1, pf2);
  for (int j = 2; j < nC; ++j) {
    const double2 P = get(j);
    double uj, f1, f2;
    eval(P, uj, f1, f2);
    const double a00 = Pp.x - F.x, a01 = Pp.y - F.y, a10 = P.x - F.x, a11 = P.y - F.y;
    const double aT = fabs(a00 * a11 - a10 * a01) * 0.5;
    if (aT > drop_area) {                                  // cutsettriangulation.hh:60
      cov += aT; npc++;
      const double su = fu + pu + uj;
      const double s1 = ff1 + pf1 + f1, s2 = ff2 + pf2 + f2;
      const double w12 = aT * (1.0 / 12.0);
      if (FMA) {
        const double d1 = fma(fu, ff1, fma(pu, pf1, uj * f1));
        const double d2 = fma(fu, ff2, fma(pu, pf2, uj * f2));
        r1 = fma(w12, fma(su, s1, d1), r1);
        r2 = fma(w12, fma(su, s2, d2), r2);
        u = fma(aT * (1.0 / 3.0), su, u);
      } else {
        const double d1 = fu * ff1 + pu * pf1 + uj * f1;
        const double d2 = fu * ff2 + pu * pf2 + uj * f2;
        r1 += w12 * (su * s1 + d1);
        r2 += w12 * (su * s2 + d2);
        u += (aT * (1.0 / 3.0)) * su;
      }
    }
    Pp = P; pu = uj; pf1 = f1; pf2 = f2;
  }
}
// generic path: the polygon of pair p did not fit the fast path's slots (several crossings of one kind; degenerate input
// only).  The whole pair is redone from global memory in the reference's own loop structure; kept out of line so that
// it costs the fast path neither registers nor spills.
struct ProjPartial { double r1, r2, u, cov; unsigned npc; };
template <bool FMA>
MMB_HDN ProjPartial proj_pair_generic(const ProjArgs& a, int K, long long p) {
  const int v[3] = { a.cv0[K], a.cv1[K], a.cv2[K] };
  const unsigned pm = a.perm[K];
  double ex[3], ey[3], kx[3], ky[3];
  for (int t = 0; t < 3; ++t) { const double2 q = proj_ld2(a.xy + v[t]); ex[t] = q.x; ey[t] = q.y; }
  for (int t = 0; t < 3; ++t) { const int s = (pm >> (2 * t)) & 3; kx[t] = s == 0 ? ex[0] : (s == 1 ? ex[1] : ex[2]); ky[t] = s == 0 ? ey[0] : (s == 1 ? ey[1] : ey[2]); }
  double2 org = make_double2(0.0, 0.0);
  if (a.prm.clip_translate) {
    org = make_double2(ex[0], ey[0]);
    for (int t = 0; t < 3; ++t) { ex[t] -= org.x; ey[t] -= org.y; }
  }
  const Aff2 gK = aff2_make(kx[0] - org.x, ky[0] - org.y, kx[1] - org.x, ky[1] - org.y, kx[2] - org.x, ky[2] - org.y);
  double cx[3], cy[3];
  proj_old_corners(a.old_xy, p, org, cx, cy);
  const Aff2 gO = aff2_make(cx[0], cy[0], cx[1], cy[1], cx[2], cy[2]);
  const double* uo = a.dofs_old + 3 * (int64_t)a.old_idx[p];
  const double u0 = uo[0], u1 = uo[1], u2 = uo[2];
  const double du1 = u1 - u0, du2 = u2 - u0;
  const double gx = du1 * gO.i00 + du2 * gO.i01, gy = du1 * gO.i10 + du2 * gO.i11;
  const double od = (cy[1] - cy[0]) * (cx[2] - cx[1]) - (cx[1] - cx[0]) * (cy[2] - cy[1]);
  const int o = (fabs(od) < 2147483648.0) ? (int)od : 0;
  if (o > 0) { double t0 = cx[1]; cx[1] = cx[2]; cx[2] = t0; t0 = cy[1]; cy[1] = cy[2]; cy[2] = t0; }
  Poly A, B; Poly* poly;
  sh_clip(cx, cy, ex, ey, a.prm.clip_eps, A, B, poly);
  ProjPartial r = { 0.0, 0.0, 0.0, 0.0, 0u };
  if (poly->n >= 3)
    proj_fan<FMA>([&](int k) { return make_double2(poly->x[k], poly->y[k]); }, poly->n, u0, gx, gy, gO.ox, gO.oy, make_double2(gK.ox, gK.oy),
                  make_double2(gK.i00, gK.i01), make_double2(gK.i10, gK.i11), a.prm.drop_area, r.r1, r.r2, r.u, r.cov, r.npc);
  return r;
}
template <int NT, bool FMA>
MMB_HD void proj_integrate(ProjTile<NT>& T, const ProjArgs& a, int j, long long p, bool uniform) {
  const int c = uniform ? (j >> 2) : T.owner[j];
  const int col = proj_column<NT>(uniform, j);
  const int nC = T.npts[col];
  double r1 = 0.0, r2 = 0.0, u = 0.0, cov = 0.0; unsigned npc = 0;
  if (nC != PROJ_HARD) {
    double cx[3], cy[3];
    proj_old_corners(a.old_xy, p, T.org[c], cx, cy);
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
  T.P[0][col] = make_double2(r1, r2);
  T.P[1][col] = make_double2(u, cov);
  T.npcs[col] = (uint8_t)npc;
}
// E: per-pair partial sums of cell c's pairs inside the chunk, in CSR order
template <int NT>
MMB_HD void proj_gather(ProjTile<NT>& T, int c, int cl, bool uniform) {
  const int lo = T.poff[c] > cl ? T.poff[c] : cl, hi = T.poff[c + 1] < cl + NT ? T.poff[c + 1] : cl + NT;
  double R1 = T.acc[0][c], R2 = T.acc[1][c], U = T.acc[2][c], cov = T.acc[3][c]; int npc = T.npc[c];
  for (int p = lo; p < hi; ++p) {
    const int col = proj_column<NT>(uniform, p - cl);
    const double2 ra = T.P[0][col], rb = T.P[1][col];
    R1 += ra.x; R2 += ra.y; U += rb.x; cov += rb.y; npc += T.npcs[col];
  }
  T.acc[0][c] = R1; T.acc[1][c] = R2; T.acc[2][c] = U; T.acc[3][c] = cov; T.npc[c] = npc;
}
// local mass-matrix solve of the new cell (nodal DG-P1 on id-ordered corners); returns the cell's mass
template <int NT>
MMB_HD double proj_finish(const ProjTile<NT>& T, const ProjArgs& a, int c) {
  const double R1 = T.acc[0][c], R2 = T.acc[1][c], U = T.acc[2][c], volK = T.vol[c];
  const double R0 = U - R1 - R2;
  const double f = 3.0 / volK;
  const double r0 = f * (4.0 * R0 - U), r1 = f * (4.0 * R1 - U), r2 = f * (4.0 * R2 - U);
  double* o = a.dofs_new + 3 * (int64_t)T.K[c];
  o[0] = r0; o[1] = r1; o[2] = r2;
  return volK * ((r0 + r1 + r2) / 3.0);
}

}  // namespace mmb
