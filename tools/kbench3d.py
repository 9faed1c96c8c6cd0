"""Config 5 (3-D, 8 M Kuhn tets): vertex movement + orient3d validity + distance estimate + marks + P0 trace/skeleton."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
import numpy as np
from mmesh_b200 import capi, synth
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
