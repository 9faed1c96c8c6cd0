// CPU build of the PRODUCT predicate header (dune-mmesh_b200/csrc/predicates.cuh) so that the interval and
// expansion stages can be unit-tested without a GPU (tests/test_host_predicates.py).  Not the oracle.
#include "predicates.cuh"
extern "C" {
int hp_orient2d(const double* p, int stage) {
  double det;
  if (stage == 1) return mmb::orient2d_filter(p[0], p[1], p[2], p[3], p[4], p[5], det);
  if (stage == 2) return mmb::orient2d_interval(p[0], p[1], p[2], p[3], p[4], p[5]);
  if (stage == 3) return mmb::orient2d_exact(p[0], p[1], p[2], p[3], p[4], p[5]);
  int s = mmb::orient2d_filter(p[0], p[1], p[2], p[3], p[4], p[5], det);
  return s != mmb::SIGN_UNKNOWN ? s : mmb::orient2d_robust(p[0], p[1], p[2], p[3], p[4], p[5]);
}
int hp_incircle(const double* p, int stage) {
  if (stage == 1) return mmb::incircle_filter(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
  if (stage == 2) return mmb::incircle_interval(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
  if (stage == 3) return mmb::incircle_exact(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
  int s = mmb::incircle_filter(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
  return s != mmb::SIGN_UNKNOWN ? s : mmb::incircle_robust(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
}
int hp_orient3d(const double* p, int stage) {
  double det;
  if (stage == 1) return mmb::orient3d_filter(p, p + 3, p + 6, p + 9, det);
  if (stage == 2) return mmb::orient3d_interval(p, p + 3, p + 6, p + 9);
  if (stage == 3) return mmb::orient3d_exact(p, p + 3, p + 6, p + 9);
  int s = mmb::orient3d_filter(p, p + 3, p + 6, p + 9, det);
  return s != mmb::SIGN_UNKNOWN ? s : mmb::orient3d_robust(p, p + 3, p + 6, p + 9);
}
void hp_batch(int which, int stage, const double* pts, long n, signed char* out) {
  const int k = which == 0 ? 6 : (which == 1 ? 8 : 12);
  for (long i = 0; i < n; ++i)
    out[i] = (signed char)(which == 0 ? hp_orient2d(pts + k * i, stage)
                                      : which == 1 ? hp_incircle(pts + k * i, stage) : hp_orient3d(pts + k * i, stage));
}
}
