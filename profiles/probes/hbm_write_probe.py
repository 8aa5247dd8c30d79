"""HBM write / read / copy bandwidth with plain torch ops (run on the GPU box)."""
import torch
n = 1 << 31
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
ms = t(lambda: a.fill_(3)); print("fill  %.1f GB/s written" % (n / ms / 1e6))
ms = t(lambda: b.copy_(a)); print("copy  %.1f GB/s read+written" % (2 * n / ms / 1e6))
a32 = a.view(torch.int32)
ms = t(lambda: a32.sum()); print("read  %.1f GB/s read" % (n / ms / 1e6))
