"""Raw PCIe copy rates of the box (pinned host memory): H2D alone, D2H alone, both at once — the ceiling of every e2e figure."""
import json, torch
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); torch.cuda.synchronize(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
import time
def wall(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
r = {"h2d_gbs": round(n / wall(h2d) / 1e6, 1), "d2h_gbs": round(n / wall(d2h) / 1e6, 1)}
tb = wall(both)
r["duplex_each_gbs"] = round(n / tb / 1e6, 1)
print(json.dumps(r))
