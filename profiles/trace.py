"""Per-CTA timeline of one encode and one decode launch (needs the FRI_TRACE=1 variant library)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FRI_CUDA_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "frave_b200", "libfri_cuda_trace.so")
import numpy as np, torch
from frave_b200 import capi

L = C.CDLL(os.environ["FRI_CUDA_LIB"])
L.fri_debug_trace.argtypes = [C.c_void_p, C.c_size_t]
dev = torch.device("cuda", 0)
W, H, Cc = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4096x4096x3").split("x"))
plan = capi.Plan(W, H, Cc)
n = plan.launch_info()["n_groups"]
px = torch.randint(0, 256, (H, W, Cc), device=dev, dtype=torch.int32).to(torch.uint8)
co = torch.empty(plan.coef_shape, dtype=torch.int32, device=dev)
out = torch.empty_like(px)
junk = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
q = np.ones(32, np.int32); q[8] = q[9] = 4


def trace(label, fn):
    for _ in range(3):
        fn()
    junk.zero_()  # flush L2
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    t = np.zeros(3 * n, np.uint64)
    assert L.fri_debug_trace(t.ctypes.data, t.size) == 0
    t = t.reshape(n, 3).astype(np.int64)
    t0 = t[:, 0].min()
    t = (t - t0) / 1e3  # us
    span = t[:, 2].max()
    print(f"{label}: {n} CTAs, span {span:.1f} us; phase1 mean {np.mean(t[:,1]-t[:,0]):.2f} us, phase2 mean {np.mean(t[:,2]-t[:,1]):.2f} us, CTA mean {np.mean(t[:,2]-t[:,0]):.2f} us")
    edges = np.arange(0, span + 4, 4)
    hs, _ = np.histogram(t[:, 0], edges)
    he, _ = np.histogram(t[:, 2], edges)
    act = [(int(((t[:, 0] <= x) & (t[:, 2] > x)).sum())) for x in edges[:-1] + 2]
    print("  t(us)   starts  ends  resident")
    for i in range(len(hs)):
        print(f"  {edges[i]:5.0f}  {hs[i]:6d} {he[i]:6d} {act[i]:6d}")
    if "decode" in label:
        t2 = np.zeros(4 * n, np.uint64)
        L.fri_debug_trace2.argtypes = [C.c_void_p, C.c_size_t]
        assert L.fri_debug_trace2(t2.ctypes.data, t2.size) == 0
        t2 = (t2.reshape(n, 4).astype(np.int64) - t0) / 1e3
        ok = t2[:, 0] > 0
        print(f"  write-out, interior CTAs: barrier -> full chunks issued  thread0 {np.mean((t2[:,0]-t[:,1])[ok]):.2f} us, thread255 {np.mean((t2[:,1]-t[:,1])[ok]):.2f} us; "
              f"-> mixed done thread0 {np.mean((t2[:,2]-t[:,1])[ok]):.2f}, thread255 {np.mean((t2[:,3]-t[:,1])[ok]):.2f}; -> CTA end {np.mean((t[:,2]-t[:,1])[ok]):.2f}")
    first = np.argsort(t[:, 0])[:592]
    print(f"  first wave: phase1 {np.mean(t[first,1]-t[first,0]):.2f} us, phase2 {np.mean(t[first,2]-t[first,1]):.2f} us; later: phase1 {np.mean(np.delete(t[:,1]-t[:,0], first)):.2f} phase2 {np.mean(np.delete(t[:,2]-t[:,1], first)):.2f}")


st = torch.cuda.current_stream().cuda_stream
trace("encode (phase1 = staging, phase2 = tiles)", lambda: plan.encode_device(px.data_ptr(), 1, co.data_ptr(), q, st))
trace("decode (phase1 = tiles, phase2 = write-out)", lambda: plan.decode_device(co.data_ptr(), 1, out.data_ptr(), q, False, st))
