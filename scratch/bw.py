import torch
x = torch.empty(int(6.7e9) // 4, dtype=torch.float32, device='cuda')
y = torch.empty_like(x)
def t(f, n=5):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.fill_(1.0)); print('fill 6.7 GB: %.3f ms  %.0f GB/s write' % (ms, 6.7e9 / ms / 1e6))
ms = t(lambda: y.copy_(x)); print('copy 6.7 GB: %.3f ms  %.0f GB/s r+w' % (ms, 2 * 6.7e9 / ms / 1e6))
ms = t(lambda: x.sum()); print('read 6.7 GB: %.3f ms  %.0f GB/s read' % (ms, 6.7e9 / ms / 1e6))
