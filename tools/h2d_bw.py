import torch, time
x = torch.empty(1, 3, 1216, 2176).pin_memory()
d = torch.empty_like(x, device="cuda")
s = torch.cuda.Stream()
for n in (1, 4):
    with torch.cuda.stream(s):
        for _ in range(3): d.copy_(x, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20): d.copy_(x, non_blocking=True)
        e1.record(s)
        torch.cuda.synchronize()
    print("H2D pinned 31.75 MB:", x.numel() * 4 * 20 / e0.elapsed_time(e1) / 1e6, "GB/s")
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty_like(big, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dbig.copy_(big, non_blocking=True); torch.cuda.synchronize()
e0.record(); dbig.copy_(big, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("H2D pinned 256 MB:", big.numel() / e0.elapsed_time(e1) / 1e6, "GB/s")
