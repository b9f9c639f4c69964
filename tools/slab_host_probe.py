"""Host-side cost of one SlabHMC trajectory: a lattice small enough that the GPU work is negligible
(L=256, one rank, eager path), so ms/trajectory ~ host issue time of the launches and torch ops."""
import os, sys, time, torch
os.environ.setdefault("U1_SLAB_GRAPH", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmc_ft_b200.sharding import SlabHMC
L = int(sys.argv[1]) if len(sys.argv) > 1 else 256
halo = int(sys.argv[2]) if len(sys.argv) > 2 else 32
slab = SlabHMC(L, 6.0, 50, 0.06, halo=halo)
own = torch.zeros(2, slab.Lx, L, dtype=torch.float64, device="cuda")
for _ in range(5):
    slab.trajectory(own)
torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
for _ in range(n):
    slab.trajectory(own)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"L={L} halo={slab.k}: host issue {1e3*(t1-t0)/n:.3f} ms/trajectory, with drain {1e3*(t2-t0)/n:.3f} ms, "
      f"launches/trajectory {slab.gpu_launches // (n + 5)}, exchanges {slab.halo.n_exchanges // (n + 5)}")
