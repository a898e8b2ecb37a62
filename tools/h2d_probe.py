"""Development aid (run under gpurun): pinned host -> device copy bandwidth of this box, alone and beside a running kernel."""
import torch, time
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, 64 << 20, 16 << 20, 4 << 20):
    torch.cuda.synchronize()
    best = 1e9
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for o in range(0, n, chunk):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"H2D {n >> 20} MB in chunks of {chunk >> 20} MB: {best:.2f} ms = {n / best / 1e6:.1f} GB/s")
# beside a memory-heavy kernel on another stream
a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
s2 = torch.cuda.Stream()
torch.cuda.synchronize()
with torch.cuda.stream(s2):
    for _ in range(20):
        b.copy_(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
print(f"H2D beside device copies: {e0.elapsed_time(e1):.2f} ms = {n / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
