"""PCIe probe for the e2e leg: pinned host<->device copy bandwidth at the sizes the host-state round trip uses."""
import torch, time
torch.cuda.set_device(0)
for mb in (1.1, 4.4, 64.0):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        t0.record()
        for _ in range(reps): fn()
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        print(f"{name} {mb:5.1f} MB: {ms*1e3:8.1f} us  {n/ms/1e6:6.1f} GB/s")
# both directions at once on two streams
n = int(32e6)
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 20
print(f"duplex 32 MB each way: {dt*1e3:.3f} ms  -> {2*n/dt/1e9:.1f} GB/s total")
