"""PCIe probe: pinned H2D alone, D2H alone, both at once on two streams (the ceiling of bench.py's e2e leg)."""
import torch
n = 1179256320 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_a = torch.empty(n, dtype=torch.float64, device="cuda"); d_b = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both(): h2d(); d2h()
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    ms = t(fn); print(f"{name}: {ms:.2f} ms  {n*8/ms/1e6:.1f} GB/s per direction")
def chunks(k):
    c = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(s1): d_a[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
            with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_b[i*c:(i+1)*c], non_blocking=True)
    return f
for k in (8, 64):
    ms = t(chunks(k)); print(f"both, {k} chunks: {ms:.2f} ms  {n*8/ms/1e6:.1f} GB/s per direction")
