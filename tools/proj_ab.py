"""A/B of the DG-P1 projection kernel variants on the config-3 mesh (development aid; needs a library built with
`make EXTRA=-DMMB_DEV_VARIANTS`).  Variant 20 = round-1 kernel (one thread per new cell); others = tile kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
import numpy as np
from mmesh_b200 import capi, synth
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 30, 31, 32, 33, 34, 35]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
m = synth.config3(nx=nx, ny=nx // 2)
s = synth.shifts_uniform(m, 0.3 * m.h, 30)
new_off, new_cell, old_xy, old_idx = synth.pairs_self_and_neighbours(m, m.xy - s)
k = np.stack([np.take_along_axis(m.cells, ((m.perm >> (2 * i)) & 3)[:, None].astype(np.int64), 1)[:, 0] for i in range(3)], 1)
pk = (m.xy - s)[k]
dofs = np.ascontiguousarray(np.sin(np.pi * pk[..., 0]) * np.sin(np.pi * pk[..., 1]) + 2.0)
ctx = capi.Context(2, 0)
ctx.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs)
ctx.upload_pairs(new_off, new_cell, old_xy, old_idx)
for mode, (tr, drop) in (("reference constants", (0, 1e-8)), ("scale-aware", (1, 0.0))):
    p = ctx.get_params(); p.clip_translate = tr; p.drop_area = drop; ctx.set_params(p)
    ref = None
    for v in variants:
        os.environ["MMB_PROJ_VARIANT"] = str(v)
        out, mass, cov, npc = ctx.project(3, dofs, download=True)
        for _ in range(2):
            ctx.project(3, None, download=False)
        ctx.synchronize(); ctx.timing_enable(True); ctx.timing_reset()
        for _ in range(reps):
            ctx.project(3, None, download=False)
        ctx.synchronize()
        tm = ctx.timing_get(); ctx.timing_enable(False); ctx.timing_reset()
        ms = tm["project_p1"][0] / tm["project_p1"][1] + (tm["project_p1_cells"][0] / tm["project_p1_cells"][1] if "project_p1_cells" in tm else 0.0)
        if ref is None:
            ref = (out, mass, cov, npc)
        err = np.abs(out - ref[0]).max() / np.abs(ref[0]).max()
        print("%-20s variant %2d  %.4f ms  pieces %d (%s)  mass rel %.2e  cov rel %.2e  dof max-rel diff %.2e %s" % (
            mode, v, ms, npc, "same" if npc == ref[3] else "DIFFERENT", abs(mass - ref[1]) / abs(ref[1]), abs(cov - ref[2]) / abs(ref[2]), err,
            "bitwise" if np.array_equal(out, ref[0]) else ""), flush=True)
