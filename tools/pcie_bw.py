"""Pinned-memory H2D / D2H bandwidth of this box (the ceiling of bench.py's e2e number)."""
import torch, time
dev = torch.device('cuda', 0)
for mb in (46, 185, 740):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device=dev)
    h2 = torch.empty(n // 6, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n // 6, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(both):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
            if both:
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / 10
    run(False); t = run(False); tb = run(True)
    print("%4d MB: H2D alone %.1f GB/s; H2D with concurrent D2H (1/6 size): %.1f GB/s H2D" % (mb, n / t / 1e9, n / tb / 1e9))
