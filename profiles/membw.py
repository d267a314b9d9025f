"""Measures HBM write-only / read-only / copy bandwidth with plain torch ops (context for the roofline)."""
import torch

dev = torch.device("cuda", 0)
n = 1 << 30


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3


x = torch.empty(n, dtype=torch.int32, device=dev)
y = torch.empty(n, dtype=torch.int32, device=dev)
t = timeit(lambda: x.fill_(7))
print(f"write-only fill_ 4 GiB: {4 * n / t / 1e9:.0f} GB/s")
t = timeit(lambda: torch.cuda.memset if False else x.zero_())
print(f"write-only zero_ 4 GiB: {4 * n / t / 1e9:.0f} GB/s")
t = timeit(lambda: x.sum())
print(f"read-only sum 4 GiB: {4 * n / t / 1e9:.0f} GB/s")
t = timeit(lambda: y.copy_(x))
print(f"copy 4+4 GiB: {8 * n / t / 1e9:.0f} GB/s (read+write)")
# 1:4 read:write mix like encode (u8 in, i32 out): y = x8.to(int32)
x8 = torch.empty(n, dtype=torch.uint8, device=dev)
t = timeit(lambda: y.copy_(x8))
print(f"u8->i32 convert (1 B read + 4 B write per elem): {5 * n / t / 1e9:.0f} GB/s")
y8 = torch.empty(n, dtype=torch.uint8, device=dev)
t = timeit(lambda: y8.copy_(x))
print(f"i32->u8 convert (4 B read + 1 B write per elem): {5 * n / t / 1e9:.0f} GB/s")
# smaller sizes (the 4096^2 RGB working set): 201 MB write
z = torch.empty(50331648, dtype=torch.int32, device=dev)
z8 = torch.empty(50331648, dtype=torch.uint8, device=dev)
t = timeit(lambda: z.copy_(z8), reps=30)
print(f"u8->i32 convert at 4096x4096x3 (252 MB): {5 * 50331648 / t / 1e9:.0f} GB/s, {t * 1e6:.1f} us")
t = timeit(lambda: z8.copy_(z), reps=30)
print(f"i32->u8 convert at 4096x4096x3 (252 MB): {5 * 50331648 / t / 1e9:.0f} GB/s, {t * 1e6:.1f} us")
