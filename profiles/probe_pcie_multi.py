"""Pinned host<->device copy bandwidth with ALL ranks copying at once (torchrun, one process per
GPU): what the box's host memory / PCIe root complexes give N GPUs together — the ceiling of the
end-to-end (host-buffer) number at N GPUs.  Rank 0 prints per-rank and aggregate figures."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.zero_(); h_out.zero_()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=6):
    torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        ts = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(ts, t)
        return [float(x) for x in ts]
    return [float(t)]
run(True, True, 1)
for name, h2d, d2h in (("H2D only", True, False), ("D2H only", False, True), ("both ways", True, True)):
    ts = run(h2d, d2h)
    if rank == 0:
        per = [n / t / 1e9 for t in ts]
        tot = sum(per) * (2 if h2d and d2h else 1)
        print(f"{world} GPU(s) at once, {name}: per rank and direction {min(per):.1f}-{max(per):.1f} GB/s, aggregate {tot:.1f} GB/s")
if world > 1: dist.destroy_process_group()
