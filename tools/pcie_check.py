import torch, time
n = 453*1024*1024//4
h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
h2 = torch.empty(386*1024*1024//4, dtype=torch.float32).pin_memory(); d2 = torch.empty_like(h2, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, k=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
a = t(lambda: d.copy_(h, non_blocking=True)); print(f"H2D 453 MiB alone: {a:.2f} ms = {453*1.048576/a:.1f} GB/s")
b = t(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H 386 MiB alone: {b:.2f} ms = {386*1.048576/b:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print(f"both directions concurrently: {c:.2f} ms per pair")
