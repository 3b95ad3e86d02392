import torch, time
d = torch.empty(100 << 20, dtype=torch.uint8, device="cuda:0")
h = torch.empty(100 << 20, dtype=torch.uint8, pin_memory=True)
for name, (src, dst) in {"D2H": (d, h), "H2D": (h, d)}.items():
    for _ in range(3): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print(name, "%.1f GB/s" % (d.numel() / dt / 1e9), "%.2f ms per 100 MiB" % (dt * 1e3))
