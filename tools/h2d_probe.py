#!/usr/bin/env python
"""Pinned host -> device copy bandwidth of this box (the floor of the e2e number): one copy of the
C2 table size, and the same bytes in 64 MiB pieces on one stream."""
import torch, time
n = 1426719514
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
one = timed(lambda: d.copy_(h, non_blocking=True))
CH = 64 << 20
def chunks():
    for o in range(0, n, CH):
        d[o:o + CH].copy_(h[o:o + CH], non_blocking=True)
many = timed(chunks)
print("H2D %d bytes: single copy %.2f ms (%.1f GB/s); 64 MiB pieces %.2f ms (%.1f GB/s)" % (n, one, n / one / 1e6, many, n / many / 1e6))
d2 = torch.empty(40 << 20, dtype=torch.uint8, device="cuda"); h2 = torch.empty(40 << 20, dtype=torch.uint8).pin_memory()
t = timed(lambda: h2.copy_(d2, non_blocking=True))
print("D2H 40 MiB: %.3f ms (%.1f GB/s)" % (t, (40 << 20) / t / 1e6))
# two streams, alternating 32 MiB pieces: does a second DMA engine add anything on this link?
s2 = torch.cuda.Stream()
def two_streams(piece):
    def f():
        k = 0
        for o in range(0, n, piece):
            with torch.cuda.stream(s if k % 2 == 0 else s2):
                d[o:o + piece].copy_(h[o:o + piece], non_blocking=True)
            k += 1
    return f
def timed2(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) * 1e3)
    return best
for piece in (16 << 20, 32 << 20, 64 << 20):
    t2 = timed2(two_streams(piece))
    print("H2D two streams, %d MiB pieces: %.2f ms (%.1f GB/s, wall clock)" % (piece >> 20, t2, n / t2 / 1e6))
t1 = timed2(chunks)
print("H2D one stream, 64 MiB pieces, wall clock: %.2f ms (%.1f GB/s)" % (t1, n / t1 / 1e6))
