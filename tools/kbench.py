"""Quick per-kernel timing on a reduced config-3 mesh (development aid; bench.py is the contract)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
import numpy as np
from mmesh_b200 import capi, synth
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
m = synth.config3(nx=nx, ny=nx // 2)
s = synth.shifts_uniform(m, 0.3 * m.h, 30)
new_off, new_cell, old_xy, old_idx = synth.pairs_self_and_neighbours(m, m.xy - s)
k = np.stack([np.take_along_axis(m.cells, ((m.perm >> (2 * i)) & 3)[:, None].astype(np.int64), 1)[:, 0] for i in range(3)], 1)
pk = (m.xy - s)[k]
dofs = np.ascontiguousarray(np.sin(np.pi * pk[..., 0]) * np.sin(np.pi * pk[..., 1]))
ctx = capi.Context(2, 0)
ctx.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs)
ctx.indicator_init()
ctx.upload_pairs(new_off, new_cell, old_xy, old_idx)
ctx.project(3, dofs, download=False)
ctx.upload_shifts(s)
for _ in range(3):
    ctx.adapt_step(None, 3, want_result=False); ctx.scale_shifts(-1.0)
ctx.synchronize()
ctx.timing_enable(True); ctx.timing_reset()
t0 = time.perf_counter()
for _ in range(steps):
    ctx.adapt_step(None, 3, want_result=False); ctx.scale_shifts(-1.0)
ctx.synchronize()
wall = (time.perf_counter() - t0) / steps * 1e3
tm = ctx.timing_get()
tot = sum(v[0] for v in tm.values()) / steps
print("variant", os.environ.get("MMB_PROJ_VARIANT", "0"), "cells", m.nc, "wall ms/step %.3f" % wall, "sum kernels %.3f" % tot,
      "-> %.3g cells/s" % (m.nc / (tot * 1e-3)))
for name, (ms, cnt) in sorted(tm.items(), key=lambda kv: -kv[1][0]):
    print("  %-18s %8.4f ms" % (name, ms / cnt))
