#!/usr/bin/env python3
"""HBM bandwidth probe: write-only (fill) and copy, large buffers, CUDA events.
Context for the roofline of write-dominated configurations (cfg4 writes 16x what it reads)."""
import torch, json
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def timeit(f, reps=10):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
res = {}
res["fill_GBs"] = n / timeit(lambda: a.zero_()) / 1e6
res["copy_rw_GBs"] = 2 * n / timeit(lambda: b.copy_(a)) / 1e6
# PCIe: pinned host <-> device, one direction and both at once (what bounds the end-to-end leg)
h = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
h2 = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
d1 = a[: 1 << 28]; d2 = b[: 1 << 28]
res["h2d_GBs"] = (1 << 28) / timeit(lambda: d1.copy_(h, non_blocking=True)) / 1e6
res["d2h_GBs"] = (1 << 28) / timeit(lambda: h2.copy_(d2, non_blocking=True)) / 1e6
s2 = torch.cuda.Stream()
def both():
    d1.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)
res["h2d_while_d2h_GBs_each"] = (1 << 28) / timeit(both) / 1e6
print(json.dumps(res))
