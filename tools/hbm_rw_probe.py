import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")   # 4 GB
y = torch.empty(1 << 30, dtype=torch.float32, device="cuda")
def t(f, n=10):
    f(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.fill_(1.0)); print(f"fill  4 GB write: {ms:.3f} ms  {4.295/ms:.2f} TB/s")
ms = t(lambda: y.copy_(x));   print(f"copy  4+4 GB:     {ms:.3f} ms  {8.59/ms:.2f} TB/s")
ms = t(lambda: x.sum());      print(f"sum   4 GB read:  {ms:.3f} ms  {4.295/ms:.2f} TB/s")
