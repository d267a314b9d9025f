"""PCIe context for the end-to-end numbers: pinned-memory copy bandwidth, one direction at a time and
both directions at once (two streams), with plain torch copies of the sizes the e2e step moves."""
import torch

dev = torch.device("cuda", 0)
n = 150 * 1000 * 1000  # bytes per direction per e2e step with 16-bit transport (151 MB)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a)
        s2.wait_event(a)
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3


for _ in range(2):
    run(True, True)
t = run(True, False)
print(f"H2D only, {n / 1e6:.0f} MB pinned: {n / t / 1e9:.1f} GB/s")
t = run(False, True)
print(f"D2H only, {n / 1e6:.0f} MB pinned: {n / t / 1e9:.1f} GB/s")
t = run(True, True)
print(f"H2D + D2H at once, {n / 1e6:.0f} MB each way: {n / t / 1e9:.1f} GB/s per direction ({t * 1e3:.2f} ms) "
      f"-> ceiling of an e2e step that moves that much each way: {4096 * 4096 / t / 1e6:.0f} MPix/s")
