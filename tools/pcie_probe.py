"""PCIe D2H/H2D bandwidth with pinned buffers: one stream vs two concurrent streams (copy engines)."""
import time
import torch
n = 100_000_000  # 400 MB of int32
d1 = torch.empty(n, dtype=torch.int32, device="cuda"); d2 = torch.empty(n, dtype=torch.int32, device="cuda")
h1 = torch.empty(n, dtype=torch.int32, pin_memory=True); h2 = torch.empty(n, dtype=torch.int32, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def one_d2h():
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True); h2.copy_(d2, non_blocking=True)
def two_d2h():
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
def one_h2d():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True); d2.copy_(h2, non_blocking=True)
def bidir():
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
for name, fn in (("D2H one stream", one_d2h), ("D2H two streams", two_d2h), ("H2D one stream", one_h2d), ("D2H + H2D concurrently", bidir)):
    t = run(fn)
    print(f"{name}: {0.8 / t:.1f} GB/s ({t * 1e3:.2f} ms for 800 MB)", flush=True)
# chunked: 8 copies of 100 MB
def chunked():
    with torch.cuda.stream(s1):
        for i in range(4):
            a, b = i * (n // 4), (i + 1) * (n // 4)
            h1[a:b].copy_(d1[a:b], non_blocking=True); h2[a:b].copy_(d2[a:b], non_blocking=True)
t = run(chunked)
print(f"D2H 8 x 100 MB one stream: {0.8 / t:.1f} GB/s ({t * 1e3:.2f} ms)")
