// CPU run of the PRODUCT DG-P1 projection phases (dune-mmesh_b200/csrc/project.cuh, the code k_project_tile runs): the
// CTA is emulated thread by thread, a barrier is the end of a loop over the threads.  tests/test_host_project.py compares
// the result with the oracle (bit for bit without FMA).  Not the oracle, and not used by the product.
#include <vector>
#include "project.cuh"
using namespace mmb;

template <int NT, bool FMA>
static void run_tiles(const ProjArgs& a, double* mass_out, double* cov_out, long* npc_out, int force_ragged) {
  constexpr int NCELL = NT / 4;
  static ProjTile<NT> T;
  long double mass = 0, cov = 0; long npc = 0;
  const int64_t nblocks = (a.nnew - a.i_begin + NCELL - 1) / NCELL;
  for (int64_t b = 0; b < nblocks; ++b) {
    const int64_t i0 = a.i_begin + b * NCELL;
    const int ncell = (int)((a.nnew - i0) < NCELL ? (a.nnew - i0) : NCELL);
    const long long p_lo = a.new_off[i0];
    const int np_tile = (int)(a.new_off[i0 + ncell] - p_lo);
    bool uniform = !force_ragged;
    for (int t = 0; t < ncell; ++t) uniform &= (proj_cell_setup<NT>(T, a, t, i0 + t, p_lo) == 4);
    T.ndense = 0;
    for (int cl = 0; cl < np_tile; cl += NT) {
      const int np = (np_tile - cl) < NT ? (np_tile - cl) : NT;
      if (!uniform) for (int t = 0; t < ncell; ++t) proj_paint<NT>(T, t, cl);
      for (int t = 0; t < np; ++t) proj_load_pair<NT>(T, a, t, p_lo + cl + t, uniform);
      for (int t = 0; t < NT; ++t) {
        int j = t; bool live = t < np;
        if (uniform) { const int q = t / NCELL, c = t - q * NCELL; j = 4 * c + q; live = c < ncell; }
        if (live && proj_clip_column<NT>(T, a, proj_column<NT>(uniform, j), uniform ? (j >> 2) : T.owner[j]))
          T.dense[T.ndense++] = (uint8_t)j;
      }
      const int nd = T.ndense;
      for (int t = 0; t < nd; ++t) { const int jd = T.dense[t]; proj_integrate<NT, FMA>(T, a, jd, p_lo + cl + jd, uniform); }
      T.ndense = 0;
      for (int t = 0; t < ncell; ++t) proj_gather<NT>(T, t, cl, uniform);
    }
    for (int t = 0; t < ncell; ++t) { mass += proj_finish<NT>(T, a, t); cov += T.acc[3][t]; npc += T.npc[t]; }
  }
  *mass_out = (double)mass; *cov_out = (double)cov; *npc_out = npc;
}

extern "C" void hp_project(int nt, int fma_on, int force_ragged, const double* xy, const int32_t* cv0, const int32_t* cv1, const int32_t* cv2,
                           const uint8_t* perm, const long long* new_off, const int32_t* new_cell, long nnew, const double* old_xy,
                           const int32_t* old_idx, const double* dofs_old, double* dofs_new, double clip_eps, double drop_area,
                           int clip_translate, double* mass, double* cov, long* npc) {
  ProjArgs a;
  a.xy = reinterpret_cast<const double2*>(xy); a.cv0 = cv0; a.cv1 = cv1; a.cv2 = cv2; a.perm = perm;
  a.new_off = new_off; a.new_cell = new_cell; a.i_begin = 0; a.nnew = nnew; a.old_xy = old_xy; a.old_idx = old_idx;
  a.dofs_old = dofs_old; a.dofs_new = dofs_new;
  a.prm = DevParams{ 0, 0, 1, 1, 4, 30, clip_eps, drop_area, clip_translate };
  if (nt == 128) { if (fma_on) run_tiles<128, true>(a, mass, cov, npc, force_ragged); else run_tiles<128, false>(a, mass, cov, npc, force_ragged); }
  else { if (fma_on) run_tiles<256, true>(a, mass, cov, npc, force_ragged); else run_tiles<256, false>(a, mass, cov, npc, force_ragged); }
}
