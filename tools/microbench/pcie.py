"""PCIe copy rates of the GPU box: pinned H2D / D2H alone and concurrently, by copy size (the ceiling of bench.py's e2e arm)."""
import torch, time
dev = torch.device("cuda")
for mb in (1, 8, 64, 256):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    reps = max(4, 1024 // mb)
    def run(up, dn):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if dn:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); return n * reps / (time.perf_counter() - t0) / 1e9
    run(True, True)
    print(f"{mb:4d} MiB copies: H2D {run(True, False):5.1f} GB/s | D2H {run(False, True):5.1f} GB/s | both: {run(True, True):5.1f} GB/s each way")
