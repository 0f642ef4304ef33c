"""Host<->device copy bandwidth of this box (pinned memory), one direction and both at once."""
import torch, json
n = 32 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, it=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
import time
def wall(fn, it=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it * 1e3
out = {"h2d_GBs": n / wall(lambda: d.copy_(h, non_blocking=True)) / 1e6, "d2h_GBs": n / wall(lambda: h2.copy_(d2, non_blocking=True)) / 1e6,
       "both_ms_for_32MiB_each": wall(both)}
out["both_GBs_each"] = n / out["both_ms_for_32MiB_each"] / 1e6
print(json.dumps(out))
