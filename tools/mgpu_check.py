"""Run under torchrun: the sharded step must reproduce the single-GPU step (counts bit-exact, mass to 1e-13).
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mgpu_check.py [nx]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
import numpy as np
import torch
import torch.distributed as dist
from mmesh_b200 import capi, multigpu, partition, synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = synth.config3(nx=nx, ny=nx // 2)
s = synth.shifts_uniform(m, 0.3 * m.h, 30)
new_off, new_cell, old_xy, old_idx = synth.pairs_self_and_neighbours(m, m.xy - s)
rng = np.random.default_rng(1)
dofs = rng.uniform(size=(m.nc, 3))
minH, maxH, factor = multigpu.global_indicator_params(m)
sh = partition.make_shard(m, world, rank, new_off, new_cell, old_xy, old_idx)
ctx = capi.Context(2, lr, stream.cuda_stream)
sm = multigpu.ShardedMesh(torch, dist, ctx, sh)
prm = ctx.get_params(); prm.minH, prm.maxH, prm.factor = minH, maxH, factor; prm.drop_area = 1e-12
prm.move_policy = 1          # the torch-driven phases below have no all-reduce of the move gate: always move
ctx.set_params(prm)
ctx.upload_pairs(sh.new_off, sh.new_cell, sh.old_xy, sh.old_idx)
ctx.project(3, dofs[sh.lcells], download=False)
sl = s[sh.lverts].copy(); sl[~sh.owned_vertex] = 0.0
ctx.upload_shifts(sl)
sm.step(3)
torch.cuda.synchronize()
got = sm.result()
# the host-buffer (pipelined) sharded step must give the same scalars and the same fields as step() + downloads
lnc, lne = sh.cells.shape[0], sh.edges.shape[0]
ctx.upload_coords(sh.xy)                                        # undo the move bit for bit
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
sl_h, dofs_h = pin(sl), pin(dofs[sh.lcells])
out = dict(dofs=torch.empty((lnc, 3), dtype=torch.float64).pin_memory(), marks=torch.empty(lnc, dtype=torch.int8).pin_memory(),
           longest=torch.empty(lnc, dtype=torch.uint8).pin_memory(), valid=torch.empty(lnc, dtype=torch.uint8).pin_memory(),
           orient=torch.empty(lnc, dtype=torch.int8).pin_memory(), eflag=torch.empty(max(lne, 1), dtype=torch.uint8).pin_memory())
sm.step_host(3, sl_h, dofs_h, lnc, out["dofs"], out["marks"], out["longest"], out["valid"], out["orient"], out["eflag"])
got_host = sm.result()
torch.cuda.synchronize()
host_fields = {k: v.numpy().copy() for k, v in out.items()}
# resident path again on restored coordinates, fields downloaded afterwards
ctx.upload_coords(sh.xy)
ctx.project(3, dofs[sh.lcells], download=False)
ctx.upload_shifts(sl)
sm.step(3)
torch.cuda.synchronize()
import ctypes
Lb = capi.lib()
res_f = dict(dofs=np.empty((lnc, 3)), marks=np.empty(lnc, np.int8), longest=np.empty(lnc, np.uint8), valid=np.empty(lnc, np.uint8),
             orient=np.empty(lnc, np.int8), eflag=np.empty(max(lne, 1), np.uint8))
Pn = lambda a: a.ctypes.data_as(ctypes.c_void_p)
ctx._ck(Lb.mmb_download_fields(ctx._h, ctypes.c_int(3), Pn(res_f["dofs"]), Pn(res_f["marks"]), Pn(res_f["longest"]), Pn(res_f["valid"]),
                               Pn(res_f["orient"]), Pn(res_f["eflag"])))
host_ok = True
own = slice(sh.own_lo, sh.own_hi)
for k in ("marks", "longest", "valid", "orient"):
    if not np.array_equal(host_fields[k][own], res_f[k][own]):
        host_ok = False; print("rank", rank, "HOST FIELD MISMATCH", k)
if not np.array_equal(host_fields["eflag"][:lne], res_f["eflag"][:lne]):
    host_ok = False; print("rank", rank, "HOST FIELD MISMATCH eflag")
if not np.array_equal(host_fields["dofs"][own], res_f["dofs"][own]):
    host_ok = False; print("rank", rank, "HOST FIELD MISMATCH dofs", np.abs(host_fields["dofs"][own] - res_f["dofs"][own]).max())
for k in got:
    if got_host[k] != got[k] and not (k in ("mass_new", "covered_area") and abs(got_host[k] - got[k]) <= 1e-13 * abs(got[k])):
        host_ok = False; print("rank", rank, "HOST SCALAR MISMATCH", k, got_host[k], got[k])
hflag = torch.tensor([1 if host_ok else 0], device="cuda")
dist.all_reduce(hflag, op=dist.ReduceOp.MIN)


ok = True
if rank == 0:
    ref = capi.Context(2, lr)
    ref.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs)
    p2 = ref.indicator_init()
    assert (p2.minH, p2.maxH, p2.factor) == (minH, maxH, factor), ((p2.minH, p2.maxH, p2.factor), (minH, maxH, factor))
    p2.drop_area = 1e-12; p2.move_policy = 1; ref.set_params(p2)
    ref.upload_pairs(new_off, new_cell, old_xy, old_idx)
    ref.project(3, dofs, download=False)
    r = ref.adapt_step(s, 3).asdict()
    for k in ("n_invalid", "n_orient_nonpos", "n_illegal", "n_refine", "n_coarsen", "n_pieces", "max_dist"):
        if got[k] != r[k]:
            ok = False; print("MISMATCH", k, got[k], r[k])
    for k in ("mass_new", "covered_area"):
        if abs(got[k] - r[k]) > 1e-13 * abs(r[k]):
            ok = False; print("MISMATCH", k, got[k], r[k])
    if hflag.item() != 1:
        ok = False; print("host-buffer sharded step differs from the resident one")
# ---- the step with its collectives INSIDE the library (mmb_adapt_step_sharded, NCCL C API, CUDA graph from step 2 on) ----
sm.init_library_comm()
lib_ok = True
for policy in (0, 1):
    prm.move_policy = policy; ctx.set_params(prm)
    fields, scal = [], []
    for rep in range(3):                                     # rep 0 eager, rep 1 captures the graph, rep 2 replays it
        ctx.upload_coords(sh.xy)
        ctx.project(3, dofs[sh.lcells], download=False)
        ctx.upload_shifts(sl)
        scal.append(sm.step_library(3, want_result=True).asdict())
        f = ctx.download_fields(3)
        f["xy"] = ctx.download_coords()
        fields.append(f)
    for rep in (1, 2):                                       # eager, captured and replayed steps agree bit for bit
        for k in fields[0]:
            if not np.array_equal(fields[rep][k], fields[0][k]):
                lib_ok = False; print("rank", rank, "policy", policy, "GRAPH/EAGER MISMATCH", k)
        if scal[rep] != scal[0]:
            lib_ok = False; print("rank", rank, "policy", policy, "GRAPH/EAGER SCALARS", scal[rep], scal[0])
    np.savez("/tmp/mmb_check_%d_%d.npz" % (policy, rank), **{k: (v[own] if k in ("marks", "longest", "valid_fp", "orient_sign", "dofs_new") else v) for k, v in fields[2].items()},
             lverts=sh.lverts, owned=sh.owned_vertex)
    dist.barrier()
    if rank == 0:
        p2.move_policy = policy; ref.set_params(p2)
        ref.upload_coords(m.xy); ref.project(3, dofs, download=False)
        r = ref.adapt_step(s, 3).asdict()
        rf = ref.download_fields(3); rxy = ref.download_coords()
        parts = [np.load("/tmp/mmb_check_%d_%d.npz" % (policy, q)) for q in range(world)]
        for k in ("marks", "longest", "valid_fp", "orient_sign", "dofs_new", "edge_flags"):
            g = np.concatenate([q[k] for q in parts])
            if not np.array_equal(g, rf[k]):
                lib_ok = False; print("policy", policy, "LIBRARY STEP FIELD MISMATCH", k)
        for q in parts:                                      # coordinates of every local vertex (owned and halo) equal the global ones
            if not np.array_equal(q["xy"], rxy[q["lverts"]]):
                lib_ok = False; print("policy", policy, "LIBRARY STEP COORDINATES MISMATCH")
        for k in ("n_invalid", "n_orient_nonpos", "n_illegal", "n_refine", "n_coarsen", "n_pieces", "max_dist", "moved"):
            if scal[2][k] != r[k]:
                lib_ok = False; print("policy", policy, "LIBRARY STEP SCALAR MISMATCH", k, scal[2][k], r[k])
        for k in ("mass_new", "covered_area"):
            if abs(scal[2][k] - r[k]) > 1e-13 * abs(r[k]):
                lib_ok = False; print("policy", policy, "LIBRARY STEP SCALAR MISMATCH", k, scal[2][k], r[k])
        if policy == 0 and not (r["n_invalid"] > 0 and r["moved"] == 0):
            lib_ok = False; print("policy 0 case does not exercise the withheld move", r)
        print("library step policy", policy, "moved", scal[2]["moved"], "OK" if lib_ok else "FAILED")
    dist.barrier()
lflag = torch.tensor([1 if lib_ok else 0], device="cuda")
dist.all_reduce(lflag, op=dist.ReduceOp.MIN)
if rank == 0:
    if lflag.item() != 1:
        ok = False
    print("mgpu_check world", world, "cells", m.nc, "OK" if ok else "FAILED", got)
    ref.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
ctx.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
