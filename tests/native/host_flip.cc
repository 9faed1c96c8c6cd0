// Host build of the PRODUCT flip logic (dune-mmesh_b200/csrc/flip.cuh: mark / select / apply and the derived edge list), run
// phase by phase over all cells exactly like the three kernels do -- every phase reads the state the previous one left, the
// atomic minimum becomes a plain minimum -- for the CPU test tests/test_host_flip.py.  No device call is made.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../dune-mmesh_b200/csrc/flip.cuh"
using namespace mmb;

extern "C" {
// cells / nb [nc][3] in/out; perm [nc] out (may be null); keys = sorted (min << 32 | max) interface edges; vid may be null.
// result: [0] flips, [1] rounds, [2] converged, [3] illegal left, [4] constrained illegal, [5] inverted, [6] broken
// trace (may be null): number of flips per round, up to max_rounds entries
static uint32_t* g_log = nullptr; static long g_log_cap = 0, g_log_n = 0;
// as hf_restore, additionally logging the flips as edge ids 3 * cell + slot (the arguments of TDS::flip) in application order
int hf_restore(const double* xy, long nv, int32_t* cells, int32_t* nb, uint8_t* perm, long nc, const uint8_t* v_if,
               const unsigned long long* keys, int nkeys, const int64_t* vid, int max_rounds, long long* result, int32_t* trace);
long hf_restore_log(const double* xy, long nv, int32_t* cells, int32_t* nb, uint8_t* perm, long nc, const uint8_t* v_if,
                    const unsigned long long* keys, int nkeys, const int64_t* vid, int max_rounds, long long* result, int32_t* trace,
                    uint32_t* log, long log_cap) {
  g_log = log; g_log_cap = log_cap; g_log_n = 0;
  hf_restore(xy, nv, cells, nb, perm, nc, v_if, keys, nkeys, vid, max_rounds, result, trace);
  g_log = nullptr;
  return g_log_n;
}
int hf_restore(const double* xy, long nv, int32_t* cells, int32_t* nb, uint8_t* perm, long nc, const uint8_t* v_if,
               const unsigned long long* keys, int nkeys, const int64_t* vid, int max_rounds, long long* result, int32_t* trace) {
  (void)nv;
  std::vector<int32_t> cv0(nc), cv1(nc), cv2(nc);
  for (long c = 0; c < nc; ++c) { cv0[c] = cells[3 * c]; cv1[c] = cells[3 * c + 1]; cv2[c] = cells[3 * c + 2]; }
  std::vector<unsigned> claim(nc);
  std::vector<uint8_t> ill(nc), sel(nc);
  FlipCounts fc; std::memset(&fc, 0, sizeof(fc));
  int rounds = 0, converged = 0;
  for (;;) {
    fc.illegal = fc.constrained = fc.inverted = fc.broken = fc.domain = 0;
    std::fill(claim.begin(), claim.end(), FLIP_NONE);
    for (long c = 0; c < nc; ++c)
      ill[c] = (uint8_t)flip_mark_cell(c, xy, cv0.data(), cv1.data(), cv2.data(), nb, v_if, keys, nkeys,
                                       [&](int64_t cell, unsigned id) { claim[cell] = std::min(claim[cell], id); },
                                       [&](int which) { (&fc.illegal)[which]++; });
    if (fc.broken || fc.inverted || fc.domain) break;
    if (fc.illegal == 0) { converged = 1; break; }
    if (rounds >= max_rounds) break;
    for (long c = 0; c < nc; ++c) sel[c] = (uint8_t)flip_select_cell(c, ill[c], nb, claim.data());
    const unsigned long long before = fc.flips;
    for (long c = 0; c < nc; ++c)
      if (sel[c]) {
        if (g_log && g_log_n < g_log_cap) g_log[g_log_n] = 3u * (uint32_t)c + (uint32_t)(sel[c] - 1);
        if (g_log) ++g_log_n;
        flip_apply_cell(c, sel[c] - 1, cv0.data(), cv1.data(), cv2.data(), nb, vid, perm); ++fc.flips;
      }
    if (trace) trace[rounds] = (int32_t)(fc.flips - before);
    ++rounds;
  }
  for (long c = 0; c < nc; ++c) { cells[3 * c] = cv0[c]; cells[3 * c + 1] = cv1[c]; cells[3 * c + 2] = cv2[c]; }
  result[0] = (long long)fc.flips; result[1] = rounds; result[2] = converged; result[3] = (long long)fc.illegal;
  result[4] = (long long)fc.constrained; result[5] = (long long)fc.inverted; result[6] = (long long)fc.broken;
  return 0;
}
// The same rounds with the bookkeeping of ACTIVE cells that mmb_restore_delaunay uses (mmesh_b200.cu): after the first round only
// the cells around a flip and the cells whose illegal edge had to wait are tested again.  Must give the same arrays, flips and
// rounds as hf_restore.  result[7] = cells tested over all sweeps (full sweeps count nc each).
int hf_restore_active(const double* xy, long nv, int32_t* cells, int32_t* nb, uint8_t* perm, long nc, const uint8_t* v_if,
                      const unsigned long long* keys, int nkeys, const int64_t* vid, int max_rounds, long long* result) {
  (void)nv;
  std::vector<int32_t> cv0(nc), cv1(nc), cv2(nc);
  for (long c = 0; c < nc; ++c) { cv0[c] = cells[3 * c]; cv1[c] = cells[3 * c + 1]; cv2[c] = cells[3 * c + 2]; }
  std::vector<unsigned> claim(nc);
  std::vector<uint8_t> ill(nc), sel(nc), act(nc, 1), act_next(nc);
  FlipCounts fc; std::memset(&fc, 0, sizeof(fc));
  int rounds = 0, converged = 0;
  bool full = true;
  long long tested = 0;
  for (;;) {
    fc.illegal = fc.constrained = fc.inverted = fc.broken = fc.domain = 0;
    std::fill(claim.begin(), claim.end(), FLIP_NONE);
    for (long c = 0; c < nc; ++c) {
      if (!act[c]) { ill[c] = 0; continue; }
      ++tested;
      ill[c] = (uint8_t)flip_mark_cell(c, xy, cv0.data(), cv1.data(), cv2.data(), nb, v_if, keys, nkeys,
                                       [&](int64_t cell, unsigned id) { claim[cell] = std::min(claim[cell], id); },
                                       [&](int which) { (&fc.illegal)[which]++; });
    }
    if (fc.broken || fc.inverted || fc.domain) break;
    const bool done = fc.illegal == 0, out_of_rounds = rounds >= max_rounds;
    if (done || out_of_rounds) {
      if (!full) { std::fill(act.begin(), act.end(), 1); full = true; continue; }
      converged = done ? 1 : 0;
      break;
    }
    std::fill(act_next.begin(), act_next.end(), 0);
    for (long c = 0; c < nc; ++c) sel[c] = (uint8_t)flip_select_cell(c, ill[c], nb, claim.data());
    for (long c = 0; c < nc; ++c) {
      if (sel[c]) {
        const int32_t n = nb[3 * c + (sel[c] - 1)];
        flip_apply_cell(c, sel[c] - 1, cv0.data(), cv1.data(), cv2.data(), nb, vid, perm); ++fc.flips;
        act_next[c] = 1; act_next[n] = 1;
        for (int j = 0; j < 3; ++j) {
          const int32_t a = nb[3 * c + j], b = nb[3 * (long)n + j];
          if (a >= 0) act_next[a] = 1;
          if (b >= 0) act_next[b] = 1;
        }
      } else if (ill[c]) act_next[c] = 1;
    }
    act.swap(act_next);
    full = false;
    ++rounds;
  }
  for (long c = 0; c < nc; ++c) { cells[3 * c] = cv0[c]; cells[3 * c + 1] = cv1[c]; cells[3 * c + 2] = cv2[c]; }
  result[0] = (long long)fc.flips; result[1] = rounds; result[2] = converged; result[3] = (long long)fc.illegal;
  result[4] = (long long)fc.constrained; result[5] = (long long)fc.inverted; result[6] = (long long)fc.broken; result[7] = tested;
  return 0;
}
// the derived edge list: returns the number of edges written (edges [cap][4])
long hf_edges(const int32_t* cells, const int32_t* nb, long nc, int32_t* edges, long cap) {
  std::vector<int32_t> cv0(nc), cv1(nc), cv2(nc);
  for (long c = 0; c < nc; ++c) { cv0[c] = cells[3 * c]; cv1[c] = cells[3 * c + 1]; cv2[c] = cells[3 * c + 2]; }
  long m = 0;
  for (long c = 0; c < nc; ++c) {
    int32_t q[12];
    const int k = flip_edge_emit(c, cv0.data(), cv1.data(), cv2.data(), nb, q);
    if (k != flip_edge_count(c, nb)) return -1;
    for (int e = 0; e < k; ++e, ++m) if (m < cap) std::memcpy(edges + 4 * m, q + 4 * e, 16);
  }
  return m;
}
uint8_t hf_perm(int64_t a, int64_t b, int64_t c) { return flip_perm_of(a, b, c); }
}
