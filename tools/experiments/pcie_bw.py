"""pinned host <-> device copy bandwidth of the box (sizes of the bench's e2e leg), alone and both directions at once"""
import torch, time
dev = torch.device("cuda", 0)
for mb in (3, 17, 64):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def t(fn, reps=10):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
    a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    c = t(both)
    print(f"{mb} MiB: H2D {n/a/1e9:.1f} GB/s ({a*1e3:.3f} ms)  D2H {n/b/1e9:.1f} GB/s ({b*1e3:.3f} ms)  both at once {2*n/c/1e9:.1f} GB/s total ({c*1e3:.3f} ms)")
