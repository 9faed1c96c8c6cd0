"""distance-field kernel: development counters + timing, warm and cold (needs a MMB_DEV_VARIANTS build)"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
import numpy as np
from mmesh_b200 import capi, synth
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = synth.config3(nx=nx, ny=nx // 2) if nx > 0 else synth.config2()
s = synth.shifts_uniform(m, 0.3 * m.h, 30)
ctx = capi.Context(2, 0)
ctx.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs)
L = capi.lib()
def run(tag):
    ctx.timing_enable(True); ctx.timing_reset()
    ctx.distance_update()
    tm = ctx.timing_get(); ctx.timing_enable(False); ctx.timing_reset()
    out = (C.c_ulonglong * 8)()
    if hasattr(L, "mmb_dev_counters"):
        L.mmb_dev_counters(ctx._h, out)
    w = max(out[0], 1)
    wide = tm["distance2d_wide"][0] / tm["distance2d_wide"][1] if "distance2d_wide" in tm else 0.0
    print("%-28s distance2d %.4f ms  wide %.4f ms  bvh %.4f ms | warps %d flat %.3f  groups/warp %.1f  group-passes/warp %.1f  evals/warp(union) %.1f  evals/lane %.2f" % (
        tag, tm["distance2d"][0] / tm["distance2d"][1], wide, tm["bvh_build2d"][0] / tm["bvh_build2d"][1], out[0], out[1] / w, out[2] / max(out[1], 1), out[3] / w, out[4] / w, out[5] / w / 32), flush=True)
run("cold (after upload)")
run("warm, mesh unchanged")
ctx.move_vertices(s); run("warm, after move +s")
ctx.move_vertices(-s); run("warm, after move -s")
ctx.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs); run("cold again")
if nx > 0 and len(sys.argv) > 2:                             # one shard of an 8-way partition: 1/8 of the vertices, the whole interface
    from mmesh_b200 import partition
    for rank in (0, 3, 5):
        sh = partition.make_shard(m, 8, rank, all_halo_lists=False)
        ctx = capi.Context(2, 0)
        ctx.upload_mesh(sh.xy, sh.cells, sh.perm, sh.edges, sh.segs)
        run("shard %d of 8 (%d vertices), cold" % (rank, sh.xy.shape[0]))
        run("shard %d of 8, warm" % rank)
        run("shard %d of 8, warm" % rank)
