import torch, time
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, (a, b) in {"d2h": (h, d), "h2d": (d, h)}.items():
    a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    print(name, round(3 * n / (time.perf_counter() - t) / 1e9, 1), "GB/s")
