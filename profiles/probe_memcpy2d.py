"""Is a pitched (2-D) device-to-host copy of the 40 new points per row as fast as a flat copy?"""
import ctypes as C, time, torch
rt = C.CDLL("libcudart.so.12")
n = 1 << 20
d = torch.zeros((n, 50), dtype=torch.float64, device="cuda")
h = torch.zeros((n, 50), dtype=torch.float64).pin_memory()
s = torch.cuda.Stream()
def flat():
    with torch.cuda.stream(s): h.copy_(d, non_blocking=True)
def pitched():
    rc = rt.cudaMemcpy2DAsync(C.c_void_p(h.data_ptr() + 80), C.c_size_t(400), C.c_void_p(d.data_ptr() + 80),
                              C.c_size_t(400), C.c_size_t(320), C.c_size_t(n), C.c_int(2), C.c_void_p(s.cuda_stream))
    assert rc == 0, rc
for fn, name, nbytes in ((flat, "flat 400 B rows", n * 400), (pitched, "pitched 320 of 400 B", n * 320)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 5
    print(f"{name}: {t*1e3:.2f} ms, {nbytes/t/1e9:.1f} GB/s payload")
