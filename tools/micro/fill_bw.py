import torch
x = torch.empty(16384, 16384, device="cuda")
y = torch.empty_like(x)
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
gb = x.numel() * 4 / 1e12
print("fill_   %.2f TB/s written" % (gb / t(lambda: x.fill_(1.0))))
print("zero_   %.2f TB/s written" % (gb / t(lambda: x.zero_())))
print("copy_   %.2f TB/s read+written" % (2 * gb / t(lambda: y.copy_(x))))
print("sum     %.2f TB/s read" % (gb / t(lambda: x.sum())))
