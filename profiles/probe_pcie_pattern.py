"""Which copy pattern reaches the duplex PCIe rate?  The host entry point's copies (field by
field, 131,072-frame chunks, 1,048,576 frames) issued (A) on three round-robin streams, upload
then download of a chunk on the same stream (what pp_plan_batch_host does), (B) on one upload
and one download stream tied by events, (C) as B with one packed copy per chunk and direction."""
import time
import torch

n, chunk = 1 << 20, 131072
h2d_fields = [8, 8, 8, 8, 4, 80, 80, 4, 4, 48, 96, 96, 96, 96]
d2h_fields = [320, 320, 4, 4, 4, 4, 4]


def bufs(fields):
    host = [torch.empty(n * b, dtype=torch.uint8).pin_memory() for b in fields]
    dev = [[torch.empty(chunk * b, dtype=torch.uint8, device="cuda") for b in fields] for _ in range(3)]
    return host, dev


hi, di = bufs(h2d_fields)
ho, do = bufs(d2h_fields)
streams = [torch.cuda.Stream() for _ in range(3)]
up, down = torch.cuda.Stream(), torch.cuda.Stream()
bi, bo = sum(h2d_fields) * n, sum(d2h_fields) * n


def pattern_a(do_up=True, do_down=True):
    for c in range(n // chunk):
        s = c % 3
        with torch.cuda.stream(streams[s]):
            if do_up:
                for f, b in enumerate(h2d_fields):
                    di[s][f].copy_(hi[f][c * chunk * b:(c + 1) * chunk * b], non_blocking=True)
            if do_down:
                for f, b in enumerate(d2h_fields):
                    ho[f][c * chunk * b:(c + 1) * chunk * b].copy_(do[s][f], non_blocking=True)


def pattern_b(packed=False):
    evs_up = [torch.cuda.Event() for _ in range(n // chunk)]
    evs_down = [torch.cuda.Event() for _ in range(n // chunk)]
    for c in range(n // chunk):
        s = c % 3
        with torch.cuda.stream(up):
            if c >= 3:
                up.wait_event(evs_down[c - 3])  # the slot's previous download is done
            if packed:
                b = sum(h2d_fields)
                pk_di[s].copy_(pk_hi[c * chunk * b:(c + 1) * chunk * b], non_blocking=True)
            else:
                for f, b in enumerate(h2d_fields):
                    di[s][f].copy_(hi[f][c * chunk * b:(c + 1) * chunk * b], non_blocking=True)
            evs_up[c].record(up)
        with torch.cuda.stream(down):
            down.wait_event(evs_up[c])
            if packed:
                b = sum(d2h_fields)
                pk_ho[c * chunk * b:(c + 1) * chunk * b].copy_(pk_do[s], non_blocking=True)
            else:
                for f, b in enumerate(d2h_fields):
                    ho[f][c * chunk * b:(c + 1) * chunk * b].copy_(do[s][f], non_blocking=True)
            evs_down[c].record(down)


pk_hi = torch.empty(n * sum(h2d_fields), dtype=torch.uint8).pin_memory()
pk_ho = torch.empty(n * sum(d2h_fields), dtype=torch.uint8).pin_memory()
pk_di = [torch.empty(chunk * sum(h2d_fields), dtype=torch.uint8, device="cuda") for _ in range(3)]
pk_do = [torch.empty(chunk * sum(d2h_fields), dtype=torch.uint8, device="cuda") for _ in range(3)]


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for name, fn in (("A round-robin streams", pattern_a), ("A uploads only", lambda: pattern_a(True, False)),
                 ("A downloads only", lambda: pattern_a(False, True)),
                 ("B upload + download streams", pattern_b), ("C as B, packed copies", lambda: pattern_b(True))):
    t = timed(fn)
    print(f"{name}: {t * 1e3:.2f} ms per 1M frames = {n / t / 1e6:.1f} M frames/s, "
          f"{bi / t / 1e9:.1f} GB/s up + {bo / t / 1e9:.1f} GB/s down")
