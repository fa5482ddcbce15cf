"""Pinned-memory D2H / H2D bandwidth of the box (what bounds the host-delivery leg of the e2e number)."""
import torch, time
for mb in (64, 800, 4000):
    n = mb * (1 << 20) // 8
    d = torch.empty(n, dtype=torch.float64, device="cuda"); h = torch.empty(n, dtype=torch.float64).pin_memory()
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt2 = (time.perf_counter() - t0) / 3
    print(f"{mb} MB: D2H {mb/1024*1.073741824/dt:.1f} GB/s, H2D {mb/1024*1.073741824/dt2:.1f} GB/s")
