"""Under torchrun: per-rank pinned H2D + D2H of the e2e step's byte counts, all ranks at once (host-link contention probe)."""
import os, time, torch, torch.distributed as dist
r, w, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h2d, d2h = 537_000_000 // w, 495_000_000 // w
a = torch.empty(h2d, dtype=torch.uint8).pin_memory(); b = torch.empty(d2h, dtype=torch.uint8).pin_memory()
da = torch.empty(h2d, dtype=torch.uint8, device="cuda"); db = torch.empty(d2h, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
for mode in ("serial", "overlap"):
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            if mode == "serial":
                da.copy_(a, non_blocking=True); b.copy_(db, non_blocking=True)
            else:
                da.copy_(a, non_blocking=True)
                with torch.cuda.stream(s2):
                    b.copy_(db, non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
    t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if r == 0: print("world", w, mode, "per-step copies: %.2f ms  (%.1f GB/s per rank per direction avg)" % (t.item() * 1e3, (h2d + d2h) / 2 / t.item() / 1e9))
dist.destroy_process_group()
