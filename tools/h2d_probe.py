import torch, time
for mb in (1.5625, 3.125, 12.5, 100):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    reps = 20
    for _ in range(reps): d.copy_(h, non_blocking=True)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(f"{mb:8.3f} MB  {ms*1e3:8.1f} us  {n/ms/1e6:6.1f} GB/s")
