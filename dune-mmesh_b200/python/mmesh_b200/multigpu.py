"""One process per GPU: Hilbert-range shards + one-ring halos; torch.distributed (NCCL over NVLink) carries only the
halo shift exchange and two tiny reductions per step (max distance; the additive step scalars).

The kernels and the pack/unpack of the halo buffers are the library's own (libmmesh_b200.so); torch is plumbing:
it wraps the library's device buffers as tensors for the collectives and provides the stream + events.
"""
import json
import os
import time

import numpy as np

from . import capi, partition


class _DevBuf:
    """Expose a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(torch, ptr, nbytes, dtype):
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device="cuda")
    item = {torch.float64: ("<f8", 8), torch.int64: ("<i8", 8)}[dtype]
    return torch.as_tensor(_DevBuf(ptr, (nbytes // item[1],), item[0]), device="cuda")


def exchange_rows(dist, send, recv, send_counts, recv_counts, rank, world):
    """Variable-size all-to-all of rows.  NCCL: one all_to_all_single; gloo (CPU tests): isend/irecv pairs."""
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(recv, send, output_split_sizes=[int(c) for c in recv_counts],
                               input_split_sizes=[int(c) for c in send_counts])
        return
    so = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
    ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    reqs = []
    for p in range(world):
        if p == rank:
            continue
        if recv_counts[p]:
            reqs.append(dist.irecv(recv[ro[p]:ro[p + 1]], src=p))
        if send_counts[p]:
            reqs.append(dist.isend(send[so[p]:so[p + 1]].contiguous(), dst=p))
    for r in reqs:
        r.wait()


class ShardedMesh:
    """A rank's shard living in one mmb context, plus the collectives of one adapt step."""

    def __init__(self, torch, dist, ctx, shard):
        self.torch, self.dist, self.ctx, self.sh = torch, dist, ctx, shard
        ctx.upload_mesh(shard.xy, shard.cells, shard.perm, shard.edges, shard.segs)
        ctx.set_owned_cells(shard.own_lo, shard.own_hi)
        ctx.halo_setup(shard.send_idx, shard.recv_idx)
        self.library_comm = False
        ps, ns = ctx.device_buffer(0)
        pr, nr = ctx.device_buffer(1)
        self.send = device_tensor(torch, ps, ns, torch.float64).view(-1, 2)
        self.recv = device_tensor(torch, pr, nr, torch.float64).view(-1, 2)
        pm, nm = ctx.device_buffer(2)
        self.maxword = device_tensor(torch, pm, nm, torch.int64)
        pv, nvb = ctx.device_buffer(3)
        self.vec = device_tensor(torch, pv, nvb, torch.float64)

    def init_library_comm(self):
        """NCCL communicator INSIDE the library (mmb_comm_init): rank 0 draws the unique id, torch.distributed only carries its
        128 bytes.  Afterwards step_library() runs the whole sharded step, collectives included, as one library call."""
        torch, dist = self.torch, self.dist
        if dist.get_backend() == "nccl":
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.sh.rank == 0:
                idt.copy_(torch.frombuffer(bytearray(capi.Context.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, src=0)
            uid = bytes(idt.cpu().numpy().tobytes())
        else:
            box = [capi.Context.comm_unique_id() if self.sh.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            uid = box[0]
        self.ctx.comm_init(self.sh.world, self.sh.rank, uid)
        self.ctx.halo_counts(self.sh.send_counts, self.sh.recv_counts)
        self.library_comm = True

    def step_library(self, ncomp, want_result=False):
        """mmb_adapt_step_sharded: kernels + halo exchange + all-reduces in one call (CUDA graph from the second step on)"""
        return self.ctx.adapt_step_sharded(None, ncomp, want_result=want_result)

    def exchange_shifts(self):
        self.ctx.halo_pack_shifts()
        exchange_rows(self.dist, self.send, self.recv, self.sh.send_counts, self.sh.recv_counts, self.sh.rank, self.sh.world)
        self.ctx.halo_unpack_shifts()

    def step(self, ncomp):
        """markElements -> ensureVertexMovement -> adapt(handle) -> moveVertices -> legality on this shard.
        The halo exchange runs while the distance field is computed (it does not need the shifts); the MAX
        all-reduce runs while validity + projection are computed (they do not need the global max distance)."""
        dist = self.dist
        nccl = dist.get_backend() == "nccl"
        self.ctx.halo_pack_shifts()
        if nccl:
            w1 = dist.all_to_all_single(self.recv, self.send, output_split_sizes=[int(c) for c in self.sh.recv_counts],
                                        input_split_sizes=[int(c) for c in self.sh.send_counts], async_op=True)
            self.ctx.step_phase_distance()
            w1.wait()
        else:
            exchange_rows(dist, self.send, self.recv, self.sh.send_counts, self.sh.recv_counts, self.sh.rank, self.sh.world)
            self.ctx.step_phase_distance()
        self.ctx.halo_unpack_shifts()                            # halo vertex exchange (owned shifts -> halos)
        w2 = dist.all_reduce(self.maxword, op=dist.ReduceOp.MAX, async_op=True)      # Distance::maximum() over all ranks
        self.ctx.step_phase(1, ncomp)                            # trial validity + projection
        w2.wait()
        self.ctx.step_phase(2, ncomp)                            # marks (need the global max)
        self.ctx.step_phase(3, ncomp)                            # move + legality + reduce vector
        dist.all_reduce(self.vec, op=dist.ReduceOp.SUM)          # counts + mass in one small vector

    def step_host(self, ncomp, shifts_h, dofs_old_h, n_old, dofs_new_h, marks_h, longest_h, valid_h, orient_h, eflag_h):
        """The same step with HOST buffers (torch pinned tensors or None), pipelined inside the library: the uploads run on
        their own stream under Distance::update and the marks, the new DOFs stream back chunk by chunk under the projection.
        Collectives sit between the three library phases; the caller reads self.vec (which orders after every copy)."""
        import ctypes
        dist, L, h = self.dist, capi.lib(), self.ctx._h
        P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        ck = self.ctx._ck
        ck(L.mmb_step_host_begin(h, P(shifts_h), ctypes.c_int(ncomp), ctypes.c_int64(n_old), P(dofs_old_h),
                                 ctypes.c_int(1 if dofs_new_h is not None else 0)))
        w = dist.all_reduce(self.maxword, op=dist.ReduceOp.MAX, async_op=True)       # Distance::maximum() over all ranks
        w.wait()
        ck(L.mmb_step_host_marks(h, P(marks_h), P(longest_h)))
        self.exchange_shifts()                                                       # ordered behind the shift upload
        ck(L.mmb_step_host_finish(h, P(dofs_new_h), P(valid_h), P(orient_h), P(eflag_h), ctypes.c_int(1), None))
        dist.all_reduce(self.vec, op=dist.ReduceOp.SUM)

    def result(self):
        v = self.vec.cpu().numpy()
        md = self.maxword.cpu().numpy().view(np.float64)[0]
        names = ["n_invalid", "n_orient_nonpos", "n_illegal", "n_refine", "n_coarsen", "n_exact_orient", "n_exact_incircle", "n_pieces"]
        out = {k: int(v[i]) for i, k in enumerate(names)}
        out.update(max_dist=float(md), mass_new=float(v[8]), covered_area=float(v[9]))
        return out


def global_indicator_params(m):
    """RatioIndicator::init (ratioindicator.hh:62-91) on the GLOBAL mesh, evaluated on the host so that every rank
    uses identical parameters (same IEEE operations as the device kernel: sqrt(dx*dx + dy*dy))."""
    def lens(a, b):
        d = m.xy[b] - m.xy[a]
        return np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    K, k = 2.0, 2.0 / (2.0 * 4.0)
    bl = lens(m.edges[:, 0], m.edges[:, 1])
    maxh, minh = bl.max(), bl.min()
    if len(m.segs):
        il = lens(m.segs[:, 0], m.segs[:, 1])
        maxH, minH = K * il.max(), k * il.min()
    else:
        maxH, minH = K * maxh, k * minh
    return float(minH), float(maxH), float(maxh / minh)


def bench_main(args, world, rank, local_rank, METRIC, UNIT, algorithmic_bytes, ClockSampler, make_workload, tune_params, projection_mode):
    # NCCL writes its version / debug lines to stdout by default; the bench prints ONE JSON line there
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    w = make_workload(args)
    m = w["mesh"]
    t0 = time.time()
    shard = partition.make_shard(m, world, rank, w["new_off"], w["new_cell"], w["old_xy"], w["old_idx"])
    part_s = time.time() - t0
    ctx = capi.Context(dim=2, device=local_rank, stream=stream.cuda_stream)
    sm = ShardedMesh(torch, dist, ctx, shard)
    sm.init_library_comm()
    minH, maxH, factor = global_indicator_params(m)
    prm = ctx.get_params()
    prm.minH, prm.maxH, prm.factor = minH, maxH, factor
    prm = projection_mode(tune_params(prm), True)
    prm.move_policy = 1                                      # as in the single-GPU bench (stress: a few cells invert)
    ctx.set_params(prm)
    ncomp = 3
    ctx.upload_pairs(shard.new_off, shard.new_cell, shard.old_xy, shard.old_idx)
    ctx.project(ncomp, w["dofs"][shard.lcells], download=False)
    # every rank supplies the shifts of its OWNED vertices only; halo entries are filled by the exchange
    s_local = w["shifts"][shard.lverts].copy()
    s_local[~shard.owned_vertex] = 0.0
    ctx.upload_shifts(s_local)

    def step():
        sm.step_library(ncomp)
        ctx.scale_shifts(-1.0)

    for _ in range(args.warmup + (args.warmup & 1)):        # even count: the mesh is back at its start position
        step()
    torch.cuda.synchronize()
    res_check = sm.step_library(ncomp, want_result=True).asdict()      # GLOBAL scalars; same phase as the single-GPU `check`
    # ---- field-level parity with a single-GPU run of the same step (outside the timed region) ----
    own = slice(shard.own_lo, shard.own_hi)
    f = ctx.download_fields(ncomp)                           # fields of the step just run (shifts = +s at the original positions)
    ctx.scale_shifts(-1.0)
    tmp = os.environ.get("MMB_TMP", "/tmp")
    np.savez(os.path.join(tmp, "mmb_fields_%d.npz" % rank), marks=f["marks"][own], longest=f["longest"][own], valid_fp=f["valid_fp"][own],
             orient_sign=f["orient_sign"][own], edge_flags=f["edge_flags"], dofs_new=f["dofs_new"][own])
    step()                                                   # back to the original positions
    torch.cuda.synchronize()
    dist.barrier()
    fields_match = None
    if rank == 0:
        parts = [np.load(os.path.join(tmp, "mmb_fields_%d.npz" % r)) for r in range(world)]
        glob = {k: np.concatenate([p[k] for p in parts]) for k in parts[0].files}
        ref = capi.Context(dim=2, device=local_rank)
        ref.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs)
        ref.set_params(prm)
        ref.upload_pairs(w["new_off"], w["new_cell"], w["old_xy"], w["old_idx"])
        ref.project(ncomp, w["dofs"], download=False)
        ref.upload_shifts(w["shifts"])
        for _ in range(2):                                   # same phase: two steps warm the hints, the third is compared
            ref.adapt_step(None, ncomp, want_result=False); ref.scale_shifts(-1.0)
        r1 = ref.adapt_step(None, ncomp).asdict()
        rf = ref.download_fields(ncomp)
        ref.close()
        fields_match = {k: bool(np.array_equal(glob[k], rf[k])) for k in glob}
        fields_match["scalars"] = bool(all(res_check[k] == r1[k] for k in ("n_invalid", "n_orient_nonpos", "n_illegal", "n_refine", "n_coarsen", "n_pieces", "max_dist"))
                                       and abs(res_check["mass_new"] - r1["mass_new"]) <= 1e-13 * abs(r1["mass_new"]))
        for r in range(world):
            os.remove(os.path.join(tmp, "mmb_fields_%d.npz" % r))
    dist.barrier()
    # ---- timed region: K sharded steps (one library call each), device-timed, max over ranks ----
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    clocks.begin()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    clocks.end()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                       # max over ranks, device-timed
    ms = float(ms.item())
    launches = ctx.launch_count - l0
    clock_window = "timed region"
    if ms * 1e-3 < 0.25:                                            # short timed region: every rank keeps the same load on
        n_ext = int(np.ceil(0.25 / max(ms / args.steps * 1e-3, 1e-6)))   # (untimed, even step count) for the 20 ms sampler
        n_ext += n_ext & 1
        for _ in range(n_ext):
            step()
        torch.cuda.synchronize()
        clocks.end()
        dist.barrier()
        clock_window = "timed region + %d identical untimed steps" % n_ext
    clk = clocks.stop() if rank == 0 else None
    if clk is not None:
        clk["window"] = clock_window
    # per-kernel times: a separate pass with per-launch events (the timed region replays a CUDA graph, which has none)
    kp = max(2, min(args.steps, 6)) // 2 * 2
    ctx.timing_enable(True)
    ctx.timing_reset()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ea.record(stream)
    for _ in range(kp):
        step()
    eb.record(stream)
    torch.cuda.synchronize()
    ms_events = ea.elapsed_time(eb) / kp
    per_kernel = ctx.timing_get()
    ctx.timing_enable(False)
    ctx.timing_reset()
    dist.barrier()
    # ---- e2e: per-rank HOST buffers (pinned); copies, halo exchange and collectives inside the timed region ----
    import ctypes
    L = capi.lib()
    lnv, lnc, lne = shard.xy.shape[0], shard.cells.shape[0], shard.edges.shape[0]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    sh_h = [pin(s_local), pin(-s_local)]
    dofs_in = pin(w["dofs"][shard.lcells])
    dofs_out = torch.empty((lnc, ncomp), dtype=torch.float64).pin_memory()
    marks_h = torch.empty(lnc, dtype=torch.int8).pin_memory()
    longest_h = torch.empty(lnc, dtype=torch.uint8).pin_memory()
    valid_h = torch.empty(lnc, dtype=torch.uint8).pin_memory()
    orient_h = torch.empty(lnc, dtype=torch.int8).pin_memory()
    eflag_h = torch.empty(max(lne, 1), dtype=torch.uint8).pin_memory()
    P = lambda t: ctypes.c_void_p(t.data_ptr())

    def step_e2e(i):
        sm.step_host(ncomp, sh_h[i & 1], dofs_in, lnc, dofs_out, marks_h, longest_h, valid_h, orient_h, eflag_h)
        vec_h = sm.vec.cpu()                                  # the step's scalars (synchronises the stream => all copies done)
        return vec_h

    for i in range(2):
        step_e2e(i)
    ke = max(2, min(args.steps, 10)) // 2 * 2
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(ke):
        step_e2e(i)
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / ke], device="cuda", dtype=torch.float64)
    dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    io = torch.tensor([lnv * 16 + lnc * ncomp * 8, lnc * ncomp * 8 + 4 * lnc + lne + 128], device="cuda", dtype=torch.float64)
    dist.all_reduce(io, op=dist.ReduceOp.SUM)
    e2e = {"value": m.nc / (float(e2e_ms.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(e2e_ms.item()),
           "h2d_bytes_per_step": int(io[0].item()), "d2h_bytes_per_step": int(io[1].item()), "steps": ke,
           "note": "per rank: shifts + old DG-P1 DOFs in (pinned host, own copy stream), halo exchange + 2 all-reduces between the pipelined library phases, marks/flags/new DOFs/scalars out; bytes summed over ranks"}
    halo = torch.tensor([shard.recv_idx.shape[0], shard.lcells.shape[0] - (shard.c1 - shard.c0), shard.lverts.shape[0]],
                        device="cuda", dtype=torch.float64)
    dist.all_reduce(halo, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_per_step = ms / args.steps
        nc = m.nc
        value = nc / (ms_per_step * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))), "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        lnv, lnc, lne, lns = shard.xy.shape[0], shard.cells.shape[0], shard.edges.shape[0], shard.segs.shape[0]
        kern = []
        for name, (tot_ms, cnt) in per_kernel.items():
            ab = algorithmic_bytes(name, lnv, lnc, lne, lns, shard.old_idx.shape[0], shard.new_cell.shape[0], ncomp)
            avg = tot_ms / cnt
            kern.append({"kernel": name, "avg_ms": avg, "share": tot_ms / (ms_events * kp), "launches_per_step": cnt / kp,
                         "algorithmic_bytes": ab, "achieved_gbs": (ab / (avg * 1e-3) / 1e9) if ab else None,
                         "frac": (ab / (avg * 1e-3) / 1e9 / peak) if ab else None})
        kern.sort(key=lambda k: -k["share"])
        dom = kern[0]
        chk = dict(res_check)
        chk["fields_match_single_gpu"] = fields_match
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": w["name"], "cells": nc, "parallelism": "hilbert-range shards x%d, one-ring halo" % world,
                           "projection": "DG-P1, self+face-neighbour pairs (phi=1), scale-aware clip (clip_translate=1, drop_area=0)",
                           "step": "mmb_adapt_step_sharded: one library call per step, kernels + NCCL collectives replayed as a CUDA graph",
                           "collectives_per_step": "grouped ncclSend/ncclRecv (halo shifts, <= %d vertices/rank), all-reduce MAX (8 B), all-reduce SUM (8 B, move gate), all-reduce SUM (128 B)" % int(halo[0].item()),
                           "halo_cells_max": int(halo[1].item()), "local_vertices_max": int(halo[2].item()),
                           "l2_policy": "per-rank working set %.2f GB per step" % (sum(k["algorithmic_bytes"] or 0 for k in kern) / 1e9)},
                "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                             "frac": dom["frac"], "traffic": None, "scope": "rank 0 shard; per-kernel events from a separate pass of %d steps (%.4f ms/step with events)" % (kp, ms_events)},
                "cpu_baseline": None,
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "kernels": kern, "check": chk,
                "setup_s": w["gen_s"], "partition_s": part_s}
        print(json.dumps(line), flush=True)
    ctx.close()
    dist.destroy_process_group()
