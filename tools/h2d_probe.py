"""Host<->device copy rates on this box (pinned vs pageable), to read the e2e number of bench.py against."""
import os
import time
import torch

# under torchrun every rank probes its own GPU at the same time (copy rates when several processes share the host)
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
RANK = os.environ.get("RANK", "0")

n = 100 * 1024 * 1024
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
pag = torch.empty(n, dtype=torch.uint8)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, src in (("pinned", pin), ("pageable", pag)):
    for _ in range(2):
        dev.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dev.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("rank " + RANK + " H2D %-9s %.1f MB in %.2f ms = %.1f GB/s" % (name, n / 1e6, dt * 1e3, n / dt / 1e9))
t0 = time.perf_counter()
for _ in range(5):
    pin.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("rank " + RANK + " D2H pinned    %.1f MB in %.2f ms = %.1f GB/s" % (n / 1e6, dt * 1e3, n / dt / 1e9))
