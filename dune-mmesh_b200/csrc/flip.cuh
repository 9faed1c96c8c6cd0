// flip.cuh -- Delaunay restoration by edge flips on the resident snapshot.
//
// What it replaces: after a vertex insertion CGAL walks the star of the new vertex and flips every edge whose opposite
// vertex lies strictly inside the circumcircle (restore_Delaunay / propagating_flip, CGAL/Delaunay_triangulation_2.h:
// 936-1022, test side_of_oriented_circle == ON_POSITIVE_SIDE, :969,999); the flip itself is
// Triangulation_data_structure_2::flip (CGAL/Triangulation_data_structure_2.h:885-915).  The reference never flips an
// interface edge: the interface is a constraint of the mesh (dune/mmesh/cgal/triangulationwrapper.hh:65-125 re-meshes a
// hole on the two sides of a constraint separately).  Here the same test drives a data-parallel Lawson iteration over
// all cells of the snapshot (cells in TDS order + face neighbours): the legality pass of the hot path tells the host
// HOW MANY edges are illegal after a move, this pass removes them without a round trip through the host TDS.
//
// One round = three data-parallel phases (every phase is one kernel; the same functions run on the host in
// tests/native/host_flip.cc):
//   mark    each cell tests its edges towards higher-numbered neighbours with the exact in-circle chain (static filter ->
//           interval -> expansions, predicates.cuh) and lowers claim[] of both cells to the edge id 3 c + k  (atomic min);
//   select  an illegal edge owns its two cells iff both claims carry its id; of two owners that touch (one's cell is a
//           face neighbour of the other's) only the one with the lower id goes on -- the globally lowest owner always
//           does, so every round flips at least one edge;
//   apply   the selected edges flip, following CGAL's slot convention (f keeps vertex(i), cw(i) becomes the opposite
//           vertex; n keeps vertex(ni), cw(ni) becomes f's apex), and repair the two back pointers of the outer cells.
// No two flips of a round share or touch a cell, so a round equals some sequential order of the same flips and the
// iteration terminates like Lawson's algorithm; the limit (the constrained Delaunay triangulation) is unique for
// points in general position, which is what the parity tests compare.
#pragma once
#include <stdint.h>
#include "predicates.cuh"

namespace mmb {

constexpr unsigned FLIP_NONE = 0xffffffffu;

struct FlipCounts {
  unsigned long long illegal;       // illegal unconstrained edges seen by the last mark phase
  unsigned long long constrained;   // illegal interface edges (never flipped)
  unsigned long long inverted;      // cells that are not positively oriented (exact sign): the input is no triangulation
  unsigned long long broken;        // neighbour table entries without a mirror entry
  unsigned long long domain;        // predicates outside the exact stage's exponent range
  unsigned long long flips;         // flips applied (accumulated over rounds)
};

MMB_HD int flip_incircle_sign(const double* p, const double* q, const double* r, const double* t) {
  int s = incircle_filter(p[0], p[1], q[0], q[1], r[0], r[1], t[0], t[1]);
  if (s == SIGN_UNKNOWN) s = incircle_robust(p[0], p[1], q[0], q[1], r[0], r[1], t[0], t[1]);
  return s;
}
MMB_HD int flip_orient_sign(const double* p, const double* q, const double* r) {
  double det;
  int s = orient2d_filter(p[0], p[1], q[0], q[1], r[0], r[1], det);
  if (s == SIGN_UNKNOWN) s = orient2d_robust(p[0], p[1], q[0], q[1], r[0], r[1]);
  return s;
}
MMB_HD int flip_slot_of(const int32_t* nb, int64_t cell, int32_t target) {
  const int32_t* p = nb + 3 * cell;
  return p[0] == target ? 0 : (p[1] == target ? 1 : (p[2] == target ? 2 : -1));
}
MMB_HD int32_t flip_vertex(const int32_t* cv0, const int32_t* cv1, const int32_t* cv2, int64_t c, int k) {
  return k == 0 ? cv0[c] : (k == 1 ? cv1[c] : cv2[c]);
}
MMB_HD void flip_set_vertex(int32_t* cv0, int32_t* cv1, int32_t* cv2, int64_t c, int k, int32_t v) {
  if (k == 0) cv0[c] = v; else if (k == 1) cv1[c] = v; else cv2[c] = v;
}
// MMesh::isInterface(edge) (mmesh.hh:607-621) on the snapshot: both vertices flagged and the (min, max) key present
MMB_HD bool flip_is_constrained(int32_t a, int32_t b, const uint8_t* v_if, const unsigned long long* keys, int nkeys) {
  if (nkeys <= 0 || !v_if[a] || !v_if[b]) return false;
  const unsigned lo_v = (unsigned)(a < b ? a : b), hi_v = (unsigned)(a < b ? b : a);
  const unsigned long long key = ((unsigned long long)lo_v << 32) | hi_v;
  int lo = 0, hi = nkeys - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const unsigned long long km = keys[mid];
    if (km == key) return true;
    if (km < key) lo = mid + 1; else hi = mid - 1;
  }
  return false;
}

// ---- phase 1: mark.  Returns the 3-bit mask of illegal, flippable edges k of cell c (only edges towards n > c are tested,
// so every interior edge has one owner).  claim_min(cell, id) lowers claim[cell] to id.
template <class ClaimMin, class Count>
MMB_HD unsigned flip_mark_cell(int64_t c, const double* xy, const int32_t* cv0, const int32_t* cv1, const int32_t* cv2,
                               const int32_t* nb, const uint8_t* v_if, const unsigned long long* keys, int nkeys,
                               ClaimMin claim_min, Count count) {
  const int32_t v[3] = { cv0[c], cv1[c], cv2[c] };
  const double* P[3] = { xy + 2 * (int64_t)v[0], xy + 2 * (int64_t)v[1], xy + 2 * (int64_t)v[2] };
  const int so = flip_orient_sign(P[0], P[1], P[2]);
  if (so == SIGN_DOMAIN) count(4); else if (so <= 0) count(2);
  unsigned mask = 0;
  for (int k = 0; k < 3; ++k) {
    const int32_t n = nb[3 * c + k];
    if (n <= (int32_t)c) continue;                          // boundary (-1) or owned by the lower cell
    const int j = flip_slot_of(nb, n, (int32_t)c);
    if (j < 0) { count(3); continue; }
    const int32_t y = flip_vertex(cv0, cv1, cv2, n, j);
    const int k1 = k == 2 ? 0 : k + 1, k2 = k == 0 ? 2 : k - 1;
    // (x, a, b) = (v[k], v[k1], v[k2]) is a rotation of the TDS order, hence counter-clockwise
    const int s = flip_incircle_sign(P[k], P[k1], P[k2], xy + 2 * (int64_t)y);
    if (s == SIGN_DOMAIN) { count(4); continue; }
    if (s <= 0) continue;                                    // ON_POSITIVE_SIDE only (Delaunay_triangulation_2.h:969)
    if (flip_is_constrained(v[k1], v[k2], v_if, keys, nkeys)) { count(1); continue; }
    count(0);
    mask |= 1u << k;
    const unsigned id = 3u * (unsigned)c + (unsigned)k;
    claim_min(c, id); claim_min((int64_t)n, id);
  }
  return mask;
}

// ---- phase 2: select.  Returns k + 1 if edge k of cell c flips this round, else 0.
MMB_HD bool flip_is_owner(unsigned id, const int32_t* nb, const unsigned* claim) {
  const int64_t cf = id / 3u; const int kf = (int)(id % 3u);
  const int32_t nf = nb[3 * cf + kf];
  return nf >= 0 && claim[cf] == id && claim[nf] == id;
}
MMB_HD int flip_select_cell(int64_t c, unsigned mask, const int32_t* nb, const unsigned* claim) {
  if (mask == 0u) return 0;
  const unsigned mine = claim[c];
  if (mine == FLIP_NONE || mine / 3u != (unsigned)c) return 0;      // the cell's lowest illegal edge belongs to a lower cell
  const int k = (int)(mine % 3u);
  if (!((mask >> k) & 1u)) return 0;
  const int32_t n = nb[3 * c + k];
  if (claim[n] != mine) return 0;
  const int j = flip_slot_of(nb, n, (int32_t)c);
  const int32_t outer[4] = { nb[3 * c + (k + 1) % 3], nb[3 * c + (k + 2) % 3], nb[3 * (int64_t)n + (j + 1) % 3], nb[3 * (int64_t)n + (j + 2) % 3] };
  for (int t = 0; t < 4; ++t) {
    const int32_t o = outer[t];
    if (o < 0) continue;
    const unsigned f = claim[o];
    if (f != FLIP_NONE && f < mine && flip_is_owner(f, nb, claim)) return 0;   // a touching owner with priority goes first
  }
  return k + 1;
}

// ---- phase 3: apply (Triangulation_data_structure_2.h:885-915).  vid may be null (persistent id = index).
MMB_HD uint8_t flip_perm_of(int64_t i0, int64_t i1, int64_t i2) {
  // perm bits 2t..2t+1 = TDS slot of the t-th corner in ascending persistent id (dune/mmesh/grid/common.hh:55-65)
  int s0 = 0, s1 = 1, s2 = 2;
  int64_t a = i0, b = i1, c = i2;
  if (a > b) { const int64_t t = a; a = b; b = t; const int u = s0; s0 = s1; s1 = u; }
  if (b > c) { const int64_t t = b; b = c; c = t; const int u = s1; s1 = s2; s2 = u; }
  if (a > b) { const int64_t t = a; a = b; b = t; const int u = s0; s0 = s1; s1 = u; }
  return (uint8_t)(s0 | (s1 << 2) | (s2 << 4));
}
MMB_HD void flip_apply_cell(int64_t f, int i, int32_t* cv0, int32_t* cv1, int32_t* cv2, int32_t* nb, const int64_t* vid, uint8_t* perm) {
  const int32_t n = nb[3 * f + i];
  const int ni = flip_slot_of(nb, n, (int32_t)f);
  const int ccw_i = (i + 1) % 3, cw_i = (i + 2) % 3, ccw_ni = (ni + 1) % 3, cw_ni = (ni + 2) % 3;
  const int32_t tr = nb[3 * f + ccw_i];                     // across (v_cw, apex): moves to n
  const int32_t bl = nb[3 * (int64_t)n + ccw_ni];           // across (v_ccw, opposite): moves to f
  const int tri = tr >= 0 ? flip_slot_of(nb, tr, (int32_t)f) : -1;
  const int bli = bl >= 0 ? flip_slot_of(nb, bl, n) : -1;
  const int32_t apex = flip_vertex(cv0, cv1, cv2, f, i), opp = flip_vertex(cv0, cv1, cv2, n, ni);
  flip_set_vertex(cv0, cv1, cv2, f, cw_i, opp);
  flip_set_vertex(cv0, cv1, cv2, n, cw_ni, apex);
  nb[3 * f + i] = bl;            if (bli >= 0) nb[3 * (int64_t)bl + bli] = (int32_t)f;
  nb[3 * f + ccw_i] = n;         nb[3 * (int64_t)n + ccw_ni] = (int32_t)f;
  nb[3 * (int64_t)n + ni] = tr;  if (tri >= 0) nb[3 * (int64_t)tr + tri] = n;
  if (perm) {
    const int64_t cells2[2] = { f, (int64_t)n };
    for (int t = 0; t < 2; ++t) {
      const int64_t c = cells2[t];
      const int32_t a = cv0[c], b = cv1[c], d = cv2[c];
      perm[c] = vid ? flip_perm_of(vid[a], vid[b], vid[d]) : flip_perm_of(a, b, d);
    }
  }
}

// ---- derived edge list (same layout as the snapshot's: ordered by (lower cell, slot); quad = (v[k+1], v[k+2], v[k], apex
// of the other cell or -1), see include/mmesh_b200.h)
MMB_HD int flip_edge_count(int64_t c, const int32_t* nb) {
  int m = 0;
  for (int k = 0; k < 3; ++k) { const int32_t n = nb[3 * c + k]; m += (n < 0 || n > (int32_t)c); }
  return m;
}
MMB_HD int flip_edge_emit(int64_t c, const int32_t* cv0, const int32_t* cv1, const int32_t* cv2, const int32_t* nb, int32_t* out4) {
  int m = 0;
  const int32_t v[3] = { cv0[c], cv1[c], cv2[c] };
  for (int k = 0; k < 3; ++k) {
    const int32_t n = nb[3 * c + k];
    if (!(n < 0 || n > (int32_t)c)) continue;
    int32_t d = -1;
    if (n >= 0) { const int j = flip_slot_of(nb, n, (int32_t)c); d = j >= 0 ? flip_vertex(cv0, cv1, cv2, n, j) : -1; }
    out4[4 * m] = v[(k + 1) % 3]; out4[4 * m + 1] = v[(k + 2) % 3]; out4[4 * m + 2] = v[k]; out4[4 * m + 3] = d;
    ++m;
  }
  return m;
}

}  // namespace mmb
