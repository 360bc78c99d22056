"""Raw pinned host -> device copy rate on this box (what bounds the host-pointer entry points): one cudaMemcpyAsync of n MB, and the
same bytes as 8 back-to-back copies; also two streams at once."""
import time
import torch

dev = torch.device("cuda", 0)
for mb in (1, 8, 38, 76):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for mode in ("one", "8 chunks", "2 streams"):
        best = 1e9
        for _ in range(8):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if mode == "one":
                d.copy_(h, non_blocking=True)
            elif mode == "8 chunks":
                c = n // 8
                for k in range(8):
                    d[k * c:(k + 1) * c].copy_(h[k * c:(k + 1) * c], non_blocking=True)
            else:
                half = n // 2
                with torch.cuda.stream(s1):
                    d[:half].copy_(h[:half], non_blocking=True)
                with torch.cuda.stream(s2):
                    d[half:].copy_(h[half:], non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print("H2D %3d MB %-9s: %.3f ms  %.1f GB/s" % (mb, mode, best * 1e3, n / best / 1e9))
