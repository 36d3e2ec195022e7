#!/usr/bin/env python3
"""GPU box probe: can H2D and D2H run concurrently on this host (PCIe full duplex through two copy engines)?"""
import time, torch
n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for name, u, d in (("H2D only", 1, 0), ("D2H only", 0, 1), ("both", 1, 1)):
    t = run(u, d); print(f"{name}: {t*1e3:.2f} ms per 64 MiB each -> {(u + d) * n / t / 1e9:.1f} GB/s aggregate")
