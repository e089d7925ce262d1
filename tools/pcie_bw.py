#!/usr/bin/env python
"""Pinned-memory PCIe bandwidth of the box (the floor of every *_host entry point): D2H / H2D, one and two streams."""
import torch, time
dev = torch.device("cuda:0")
n = 480_000_000
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
print("D2H 480MB pinned: %.2f ms  %.1f GB/s" % ((lambda x: (x * 1e3, n / x / 1e9))(t(lambda: h.copy_(d, non_blocking=True)))))
print("H2D 480MB pinned: %.2f ms  %.1f GB/s" % ((lambda x: (x * 1e3, n / x / 1e9))(t(lambda: d.copy_(h, non_blocking=True)))))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    with torch.cuda.stream(s1): h[: n // 2].copy_(d[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): h[n // 2:].copy_(d[n // 2:], non_blocking=True)
print("D2H 2 streams: %.2f ms  %.1f GB/s" % ((lambda x: (x * 1e3, n / x / 1e9))(t(two))))
def chunks():
    for c in range(8):
        a, b = c * n // 8, (c + 1) * n // 8
        h[a:b].copy_(d[a:b], non_blocking=True)
print("D2H 8 chunks: %.2f ms  %.1f GB/s" % ((lambda x: (x * 1e3, n / x / 1e9))(t(chunks))))
hp = torch.empty(n, dtype=torch.uint8)
print("D2H pageable: %.2f ms" % (t(lambda: hp.copy_(d)) * 1e3))
