#!/usr/bin/env python
"""Host -> device copy rate from pinned memory on this box: one stream against two streams that share the transfer
(what decides whether a large upload is cut in two).  One line per case on stdout."""
import torch

n = 256 << 20
src = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(parts):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s1)
    s2.wait_event(e0)
    for (a, b), st in parts:
        with torch.cuda.stream(st):
            dst[a:b].copy_(src[a:b], non_blocking=True)
    ej = torch.cuda.Event()
    ej.record(s2)
    s1.wait_event(ej)
    e1.record(s1)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for name, parts in [("one stream", [((0, n), s1)]), ("two streams", [((0, n // 2), s1), ((n // 2, n), s2)])]:
    best = min(run(parts) for _ in range(5))
    print("h2d %s: %d MB in %.2f ms = %.1f GB/s" % (name, n >> 20, best, n / best / 1e6))
for name, parts in [("one stream", [((0, n), s1)]), ("two streams", [((0, n // 2), s1), ((n // 2, n), s2)])]:
    def back(parts=parts):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s1)
        s2.wait_event(e0)
        for (a, b), st in parts:
            with torch.cuda.stream(st):
                src[a:b].copy_(dst[a:b], non_blocking=True)
        ej = torch.cuda.Event()
        ej.record(s2)
        s1.wait_event(ej)
        e1.record(s1)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    best = min(back() for _ in range(5))
    print("d2h %s: %d MB in %.2f ms = %.1f GB/s" % (name, n >> 20, best, n / best / 1e6))
