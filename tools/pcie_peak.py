"""Development aid: the PCIe ceiling of the host-resident step (bench.py's e2e).  Copies the bytes one 8192^2 step moves
(2.15 GB each way) between pinned host memory and the device: one direction alone, then both directions at once on
two streams (what e2d_step_host_streamed overlaps)."""
import time

import torch

n = 4 * 8196 * 8196
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
h_in.fill_(1.0)
d_a = torch.empty(n, dtype=torch.float64, device="cuda")
d_b = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
nbytes = n * 8


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


for name, fn in (("H2D alone", h2d), ("D2H alone", d2h), ("H2D + D2H concurrently", both)):
    t = timed(fn)
    print(f"{name:26s} {t*1e3:7.2f} ms  {nbytes/t*1e-9:6.1f} GB/s per direction", flush=True)

