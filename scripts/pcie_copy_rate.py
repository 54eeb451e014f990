import torch, time
for mb in (0.2, 2.0, 2.77, 4.0, 32.0):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, a, b in (("H2D", h, d), ("D2H", d, h)):
        for _ in range(5): b.copy_(a, non_blocking=True)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): b.copy_(a, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print("%s %.2f MB: %.1f us  %.1f GB/s" % (name, mb, us, n / us / 1e3))
