"""Back-to-back dependent launch latency of trivial kernels on this box (platform floor for small-problem attempts)."""
import time, torch
x = torch.zeros(1, device="cuda")
big = torch.zeros(1 << 20, device="cuda")
for name, fn in (("1-element add_", lambda: x.add_(1.0)), ("1M-element add_", lambda: big.add_(1.0))):
    for _ in range(100): fn()
    torch.cuda.synchronize()
    n = 2000
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / n * 1e3:.2f} us per launch (device time, back to back)")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(200): fn()
        torch.cuda.synchronize()
        a.record(s); g.replay(); b.record(s); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / 200 * 1e3:.2f} us per launch inside a CUDA graph")
