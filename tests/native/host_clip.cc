// CPU build of the PRODUCT clip arithmetic (dune-mmesh_b200/csrc/clip.cuh): the three clip stages exactly as the
// projection kernel chains them (3 -> <=4 -> <=6 -> <=9 points, ping-pong buffers), so that the case analysis on
// comparison bitmasks can be checked against the oracle's reference-structured Sutherland-Hodgman without a GPU
// (tests/test_host_clip.py).  Not the oracle.
#include "clip.cuh"
extern "C" {
// subj6 / clip6: three (x, y) corners each; out: up to 9 points; returns the point count
int hc_clip(const double* subj6, const double* clip6, double eps, double* out) {
  double2 P[6], Q[9];
  for (int k = 0; k < 3; ++k) P[k] = make_double2(subj6[2 * k], subj6[2 * k + 1]);
  const double ex[4] = { clip6[0], clip6[2], clip6[4], clip6[0] }, ey[4] = { clip6[1], clip6[3], clip6[5], clip6[1] };
  auto ldP = [&](int k, double& x, double& y) { x = P[k].x; y = P[k].y; };
  auto ldQ = [&](int k, double& x, double& y) { x = Q[k].x; y = Q[k].y; };
  auto stP = [&](int k, double x, double y) { P[k] = make_double2(x, y); };
  auto stQ = [&](int k, double x, double y) { Q[k] = make_double2(x, y); };
  int n = 3;
  n = mmb::clip_stage<3>(ldP, n, ex[0], ey[0], ex[1], ey[1], eps, stQ);
  n = mmb::clip_stage<4>(ldQ, n, ex[1], ey[1], ex[2], ey[2], eps, stP);
  n = mmb::clip_stage<6>(ldP, n, ex[2], ey[2], ex[3], ey[3], eps, stQ);
  for (int k = 0; k < n; ++k) { out[2 * k] = Q[k].x; out[2 * k + 1] = Q[k].y; }
  return n;
}
// one clip stage on a general (possibly non-convex) polygon of n <= 6 points: several crossings of one kind occur, which
// the triangle/triangle chain above practically never produces
int hc_clip_edge(const double* poly, int n, const double* q1, const double* q2, double eps, double* out) {
  double2 P[6], Q[9];
  for (int k = 0; k < n; ++k) P[k] = make_double2(poly[2 * k], poly[2 * k + 1]);
  auto ldP = [&](int k, double& x, double& y) { x = P[k].x; y = P[k].y; };
  auto stQ = [&](int k, double x, double y) { Q[k] = make_double2(x, y); };
  const int m = mmb::clip_stage<6>(ldP, n, q1[0], q1[1], q2[0], q2[1], eps, stQ);
  for (int k = 0; k < m; ++k) { out[2 * k] = Q[k].x; out[2 * k + 1] = Q[k].y; }
  return m;
}
void hc_clip_batch(const double* subj, const double* clip, long n, double eps, double* out /*[n][18]*/, int* count) {
  for (long i = 0; i < n; ++i) count[i] = hc_clip(subj + 6 * i, clip + 6 * i, eps, out + 18 * i);
}
}
