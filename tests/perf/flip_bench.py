"""Delaunay restoration by device flips (mmb_restore_delaunay) on the config-3 mesh after one vertex move, timed beside the
oracle's sequential Lawson iteration on the host.  usage: flip_bench.py [nx] [amp] [nooracle]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
import numpy as np
from mmesh_b200 import capi, synth
from oracle import pyoracle as oracle
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
amp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
t0 = time.time()
m = synth.config3(nx=nx, ny=nx // 2)
s = synth.shifts_uniform(m, amp * m.h, 30)
s[np.unique(m.segs.ravel())] = 0.0
for _ in range(60):                                          # ensureVertexMovement: no inverted cell
    valid, _, ninv = oracle.trial_validate(m.xy + s, None, m.cells)
    if ninv == 0:
        break
    s[np.unique(m.cells[valid == 0].ravel())] *= 0.5
xy = m.xy + s
print("mesh: %d cells, %d edges, %d facets (setup %.1f s)" % (m.nc, m.edges.shape[0], m.segs.shape[0], time.time() - t0), flush=True)
ctx = capi.Context(2, 0)
ctx.upload_mesh(xy, m.cells, m.perm, m.edges, m.segs)
ctx.upload_neighbours(m.nb)
_, nill = ctx.legality()
ctx.synchronize()
ctx.timing_enable(True); ctx.timing_reset()
t1 = time.perf_counter()
res = ctx.restore_delaunay(vertex_id=m.vid)
ctx.synchronize()
t2 = time.perf_counter()
tm = ctx.timing_get(); ctx.timing_enable(False)
print("device: %d illegal edges -> %s in %.2f ms wall" % (nill, res, 1e3 * (t2 - t1)), flush=True)
for k in sorted(tm, key=lambda k: -tm[k][0]):
    print("  %-16s %9.3f ms total  %5d launches  %.4f ms each" % (k, tm[k][0], tm[k][1], tm[k][0] / tm[k][1]))
cells, perm, nb, edges = ctx.download_topology()
flags, nill2 = ctx.legality()
print("after: %d illegal edges (constrained %d)" % (nill2, res["n_constrained_illegal"]), flush=True)
if len(sys.argv) > 3 and sys.argv[3] == "nooracle":
    sys.exit(0)
t3 = time.perf_counter()
oc, onb, oflips = oracle.restore_delaunay_2d(xy, m.cells, m.nb, m.segs)
t4 = time.perf_counter()
same = np.array_equal(np.unique(np.sort(oc, axis=1), axis=0), np.unique(np.sort(cells, axis=1), axis=0))
print("oracle (sequential Lawson, 1 core): %d flips in %.1f ms; same triangulation: %s" % (oflips, 1e3 * (t4 - t3), same), flush=True)
