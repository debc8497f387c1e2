#!/usr/bin/env python
"""Host<->device copy rates of this box for the step's transfer sizes (the bound of bench.py's e2e):
H2D alone, D2H alone, and both directions at once on two streams (pinned memory, CUDA events)."""
import sys
import torch

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 11.1
n = int(mb * 1e6) // 4
dev = torch.device("cuda:0")
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device=dev)
d_out = torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=50):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3  # us per rep


for name, f in (("h2d", (1, 0)), ("d2h", (0, 1)), ("duplex", (1, 1))):
    run(*f, reps=5)
    us = run(*f)
    print("%-7s %.1f MB per direction: %8.1f us  -> %6.1f GB/s per direction" % (name, mb, us, mb * 1e6 / us / 1e3))
