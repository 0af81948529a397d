"""PCIe probe (diagnostic): does splitting one host->device copy over two streams, or using a numpy-allocated /
cudaHostAlloc'ed buffer, move more than one stream does?"""
import torch, time
n = 18_464_768 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]
def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
def split(k):
    c = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(streams[i]): d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
    return f
for k in (1, 2, 4):
    ms = t(split(k)); print("h2d over", k, "streams:", round(ms, 4), "ms", round(n * 4 / ms / 1e6, 1), "GB/s")
d2 = torch.ones(n, dtype=torch.float32, device="cuda"); h2 = torch.empty(n, dtype=torch.float32).pin_memory()
def split_out(k):
    c = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(streams[i]): h2[i * c:(i + 1) * c].copy_(d2[i * c:(i + 1) * c], non_blocking=True)
    return f
for k in (1, 2):
    ms = t(split_out(k)); print("d2h over", k, "streams:", round(ms, 4), "ms", round(n * 4 / ms / 1e6, 1), "GB/s")
