"""Config 5 (3-D, 8 M Kuhn tets): vertex movement + orient3d validity + distance estimate + marks + P0 trace/skeleton."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
import numpy as np
from mmesh_b200 import capi, synth
from oracle import pyoracle as oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 110
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
t0 = time.time(); m = synth.lattice3d(n); print("gen %.1fs" % (time.time() - t0), "tets", m.nc, "vertices", m.nv, "facets", m.tris.shape[0])
s = synth.shifts_uniform(m, 0.05 * m.h, 50)
ctx = capi.Context(3, 0)
ctx.upload_mesh(m.xyz, m.cells, m.perm, None, m.tris)
prm = ctx.get_params(); prm.minH, prm.maxH, prm.factor = 0.25 * m.h, 1.3 * m.h, 1.5; ctx.set_params(prm)
ctx.upload_shifts(s)
for _ in range(3):
    ctx.adapt_step(None, 0, want_result=False); ctx.scale_shifts(-1.0)
ctx.synchronize()
ctx.timing_enable(True); ctx.timing_reset()
for _ in range(steps):
    ctx.adapt_step(None, 0, want_result=False); ctx.scale_shifts(-1.0)
ctx.synchronize()
tm = ctx.timing_get()
tot = sum(v[0] for v in tm.values()) / steps
res = ctx.adapt_step(None, 0)
print("config5: %d tets, %.3f ms/step -> %.3g cells/s" % (m.nc, tot, m.nc / (tot * 1e-3)), res.asdict())
for name, (ms, cnt) in sorted(tm.items(), key=lambda kv: -kv[1][0]):
    print("  %-18s %8.4f ms" % (name, ms / cnt))
# the oracle on the host: validity and marks over all cells, the brute-force distance estimate (distance.hh:157-173, O(Nv * Ns)) on a
# sample of the vertices and scaled to all of them
d_gpu, _ = ctx.distance_update()
xyz = ctx.download_coords()                                 # the mesh where the steps above left it
for thr in (1, os.cpu_count()):
    oracle.set_threads(thr)
    t0 = time.perf_counter(); oracle.trial_validate(xyz, s, m.cells, dim=3); t1 = time.perf_counter()
    nsub = 4096 * thr
    vidx = np.linspace(0, m.nv - 1, nsub).astype(np.int64)
    t2 = time.perf_counter(); od = oracle.distance_3d_subset(xyz, vidx, m.tris); t3 = time.perf_counter()
    oprm = oracle.Params.from_buffer_copy(bytes(ctx.get_params()))
    t4 = time.perf_counter(); oracle.mark(xyz, d_gpu, m.cells, m.perm, oprm, oprm.distProportion * float(d_gpu.max()), dim=3); t5 = time.perf_counter()
    t_dist = (t3 - t2) * m.nv / nsub
    tot_cpu = (t1 - t0) + t_dist + (t5 - t4)
    print("  oracle, %2d thread(s): validity %.0f ms, marks %.0f ms, brute-force distance %.0f ms (from %d of %d vertices) -> %.3g cells/s" % (
        thr, 1e3 * (t1 - t0), 1e3 * (t5 - t4), 1e3 * t_dist, nsub, m.nv, m.nc / tot_cpu))
    assert np.array_equal(od, d_gpu[vidx])
oracle.set_threads(1)
