// C wrapper around the PRODUCT host function Dune::B200::buildProjectionPairs (dune-mmesh_b200/host/pairs.hh) for the CPU
// test tests/test_host_pairs.py.  No device call is made (the class template MMesh<dim> is only declared, never instantiated).
#include <cstring>
#include "../../dune-mmesh_b200/host/pairs.hh"
using namespace Dune::B200;

static Snapshot<2> snap(long nv, const double* x, long nc, const int32_t* cells, const uint8_t* perm) {
  Snapshot<2> s;
  s.x.resize(nv); s.cells.resize(nc); s.perm.assign(perm, perm + nc);
  for (long i = 0; i < nv; ++i) s.x[i] = { x[2 * i], x[2 * i + 1] };
  for (long i = 0; i < nc; ++i) s.cells[i] = { cells[3 * i], cells[3 * i + 1], cells[3 * i + 2] };
  return s;
}
static ProjectionPairs g_pairs;
extern "C" {
// returns the number of pairs (or -1 with the message in err); the arrays are fetched with hpair_get
long hpair_build(long nvo, const double* xo, const int64_t* ido, long nco, const int32_t* co, const uint8_t* po, long nvn,
                 const double* xn, const int64_t* idn, long ncn, const int32_t* cn, const uint8_t* pn, const int32_t* comp_old,
                 long* nnew, long* ncomp, char* err, int errlen) {
  try {
    const Snapshot<2> so = snap(nvo, xo, nco, co, po), sn = snap(nvn, xn, ncn, cn, pn);
    std::vector<int64_t> io(ido, ido + nvo), in(idn, idn + nvn);
    std::vector<int32_t> comp;
    if (comp_old) comp.assign(comp_old, comp_old + nco);
    g_pairs = buildProjectionPairs(so, io, sn, in, comp_old ? &comp : nullptr);
    *nnew = (long)g_pairs.newCell.size(); *ncomp = (long)g_pairs.components;
    return (long)g_pairs.oldIndex.size();
  } catch (const std::exception& e) {
    std::strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0;
    return -1;
  }
}
void hpair_get(int64_t* new_off, int32_t* new_cell, double* old_xy, int32_t* old_idx, int32_t* comp_old, int32_t* comp_new) {
  std::memcpy(new_off, g_pairs.newOff.data(), g_pairs.newOff.size() * 8);
  std::memcpy(new_cell, g_pairs.newCell.data(), g_pairs.newCell.size() * 4);
  if (!g_pairs.oldIndex.empty()) {
    std::memcpy(old_xy, g_pairs.oldCorners.data(), g_pairs.oldCorners.size() * 48);
    std::memcpy(old_idx, g_pairs.oldIndex.data(), g_pairs.oldIndex.size() * 4);
  }
  std::memcpy(comp_old, g_pairs.componentOfOld.data(), g_pairs.componentOfOld.size() * 4);
  std::memcpy(comp_new, g_pairs.componentOfNew.data(), g_pairs.componentOfNew.size() * 4);
}
}
