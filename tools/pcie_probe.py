import torch, time
dev = torch.device("cuda:0")
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
h2 = torch.empty(1 << 29, dtype=torch.uint8).pin_memory(); d2 = torch.empty(1 << 29, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
dt = t(lambda: d.copy_(h, non_blocking=True)); print("H2D alone GB/s", (1 << 30) / dt / 1e9)
dt = t(lambda: h2.copy_(d2, non_blocking=True)); print("D2H alone GB/s", (1 << 29) / dt / 1e9)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
dt = t(both); print("both: H2D GB/s", (1 << 30) / dt / 1e9, "D2H GB/s (concurrent)", (1 << 29) / dt / 1e9)
