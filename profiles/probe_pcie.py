"""Pinned host<->device copy bandwidth on this box (floor for the end-to-end number)."""
import time, torch
n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
run(True, True, 1)
a = run(True, False); b = run(False, True); c = run(True, True)
print(f"H2D {n/a/1e9:.1f} GB/s  D2H {n/b/1e9:.1f} GB/s  both: {n/c/1e9:.1f} GB/s each way ({2*n/c/1e9:.1f} total)")
for sz in (1 << 20, 4 << 20, 16 << 20):
    k = n // sz
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s2):
        for i in range(k): h_out[i*sz:(i+1)*sz].copy_(d_out[i*sz:(i+1)*sz], non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"D2H in {sz>>20} MiB pieces: {n/t/1e9:.1f} GB/s")
