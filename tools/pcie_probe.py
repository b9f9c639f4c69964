"""PCIe probe: H2D, D2H, concurrent both, for pinned 256 MiB buffers (CUDA events)."""
import torch
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
def sliced():
    k = 8; m = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d1[i*m:(i+1)*m].copy_(h_in[i*m:(i+1)*m], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*m:(i+1)*m].copy_(d2[i*m:(i+1)*m], non_blocking=True)
for name, fn in (("H2D", h2d), ("D2H", d2h), ("both", both), ("both sliced x8", sliced)):
    ms = timed(fn)
    print(f"{name:16s} {ms:7.3f} ms  -> {n / ms / 1e6:6.1f} GB/s per direction")
