// CPU build of the PRODUCT distance / RatioIndicator arithmetic (dune-mmesh_b200/csrc/geom.cuh, the functions k_distance2d
// and k_mark2d call) for comparison with the oracle without a GPU (tests/test_host_geom.py).  Not the oracle.
#include "geom.cuh"
extern "C" {
void hg_seg_distance_batch(const double* p, const double* v0, const double* v1, long n, double* out) {
  for (long i = 0; i < n; ++i) {
    const mmb::SegLeaf f = mmb::seg_leaf_make(v0[2 * i], v0[2 * i + 1], v1[2 * i], v1[2 * i + 1]);
    out[i] = mmb::seg_distance(f, p[2 * i], p[2 * i + 1]);
  }
}
// corners6: id-ordered corners of each triangle; d3: the three vertex distances; prm8: minH, maxH, factor,
// distProportion, edgeRatio, radiusRatio, clip_eps, drop_area
void hg_indicator_batch(const double* corners6, const double* d3, long n, const double* prm8, double maxDist,
                        signed char* mark, unsigned char* longest) {
  mmb::DevParams prm;
  prm.minH = prm8[0]; prm.maxH = prm8[1]; prm.factor = prm8[2]; prm.distProportion = prm8[3];
  prm.edgeRatio = prm8[4]; prm.radiusRatio = prm8[5]; prm.clip_eps = prm8[6]; prm.drop_area = prm8[7]; prm.clip_translate = 0;
  for (long i = 0; i < n; ++i) {
    const double* c = corners6 + 6 * i;
    int le = -1;
    const int m = mmb::indicator2d(make_double2(c[0], c[1]), make_double2(c[2], c[3]), make_double2(c[4], c[5]),
                                   d3[3 * i], d3[3 * i + 1], d3[3 * i + 2], prm, maxDist, le);
    mark[i] = (signed char)m; longest[i] = (unsigned char)le;
  }
}
}
