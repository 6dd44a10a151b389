"""Bare pinned host -> device copy rate: one vs two copy streams, several chunk sizes (no kernels)."""
import torch, time
dev = torch.device("cuda", 0)
total = 1843200000
for chunk_mb in (32, 64, 154, 512):
    for nstreams in (1, 2, 4):
        n = chunk_mb * 1000 * 1000
        host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2 * nstreams)]
        devb = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2 * nstreams)]
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        reps = max(1, total // n)
        best = 0.0
        for trial in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(reps):
                s = streams[i % nstreams]
                with torch.cuda.stream(s):
                    devb[i % len(devb)].copy_(host[i % len(host)], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = max(best, reps * n / dt / 1e9)
        print(f"chunk {chunk_mb:4d} MB  streams {nstreams}: {best:6.2f} GB/s")
