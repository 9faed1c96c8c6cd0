// geom.cuh -- the per-facet distance and the RatioIndicator arithmetic of the hot path, host-compilable so that the very
// code the kernels run can be compared with the oracle without a GPU (tests/native/host_geom.cc).
// Compile with -fmad=false (device) / -ffp-contract=off (host).
#pragma once
#include <math.h>
#include "predicates.cuh"      // MMB_HD, det2
#include "clip.cuh"            // host shim of double2

namespace mmb {

struct DevParams {
  double minH, maxH, factor, distProportion, edgeRatio, radiusRatio, clip_eps, drop_area;
  int clip_translate;
};

struct SegLeaf { double c0x, c0y, c1x, c1y, n0, n1; };   // facet corners (AffineGeometry::corner) + Plane normal

// interface facet (v0, v1) as Distance::computeDistance sees it (distance.hh:143-156): corners through
// AffineGeometry::corner, unit normal of Plane (triangulationwrapper.hh:14-40: n = a - b, swap, negate first, normalise)
MMB_HD SegLeaf seg_leaf_make(double v0x, double v0y, double v1x, double v1y) {
  SegLeaf f;
  f.c0x = v0x; f.c0y = v0y;
  f.c1x = v0x + 1.0 * (v1x - v0x);                     // corner(1) = origin + J^T e_1
  f.c1y = v0y + 1.0 * (v1y - v0y);
  double n0 = f.c0x - f.c1x, n1 = f.c0y - f.c1y;       // Plane: n_ = a - b
  { const double tt = n0; n0 = n1; n1 = tt; }          // std::swap(n_[0], n_[1])
  n0 *= -1.0;
  double nn = 0.0; nn += n0 * n0; nn += n1 * n1; nn = sqrt(nn);
  f.n0 = n0 / nn; f.n1 = n1 / nn;
  return f;
}
MMB_HD double seg_distance(const SegLeaf& f, double px, double py) {
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

// RatioIndicator::operator() for a bulk triangle, corners in id order (ratioindicator.hh:106-160)
MMB_HD double edge_len2d(double2 a, double2 b) {
  const double jx = b.x - a.x, jy = b.y - a.y;
  double s = 0.0; s += jx * jx; s += jy * jy;
  return sqrt(s);
}
MMB_HD int indicator2d(const double2 c0, const double2 c1, const double2 c2, double d0, double d1,
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

}  // namespace mmb
