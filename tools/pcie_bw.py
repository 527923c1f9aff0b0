"""Measures what the host link of this box can move: pinned H2D, D2H, and both at once (the e2e ceiling of bench.py)."""
import time, torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return reps * n / (time.perf_counter() - t0) / 1e9
run(True, True, 1)
print("H2D only  %.1f GB/s" % run(True, False))
print("D2H only  %.1f GB/s" % run(False, True))
print("both      %.1f GB/s each direction" % run(True, True))
