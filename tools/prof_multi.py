"""Time the temporally blocked leapfrog kernel (4 steps per launch) on one L x L lattice."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmc_ft_b200 import _lib
lib = _lib.load()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Lx = int(sys.argv[2]) if len(sys.argv) > 2 else L
a = [torch.randn(1, 2, Lx, L, dtype=torch.float64, device="cuda") for _ in range(4)]
flag = C.c_int(0)
sp = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(lib.u1_leapfrog_steps_rect(a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(), 1, Lx, L,
                                          6.0, 0.06, 0.06, 4, C.byref(flag), sp))
for _ in range(20): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"Lx={Lx} L={L}: {ms*1e3:.1f} us per 4 steps -> {Lx*L*4/ms/1e6:.1f} G site-updates/s, {Lx*L*4*64/ms/1e9:.2f} TB/s algorithmic")
