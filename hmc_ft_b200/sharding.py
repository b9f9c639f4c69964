"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

(i)  Chain sharding (BASELINE configs[1], [3]): independent chains are split over ranks; there is no
     data-path collective — Philox streams are keyed by the GLOBAL chain id, so the ensemble is the
     same for any GPU count; only per-chain observables are gathered.
(ii) Domain decomposition of one large lattice (configs[4], L=4096): 1-D slabs along x.  Each rank
     keeps an extended slab (own rows + k halo rows per side).  Halos of (theta, pi) are exchanged
     with the two ring neighbours once every k leapfrog steps (NCCL send/recv over NVLink); between
     exchanges the extended slab is advanced as a whole and the outer rows, which become stale one
     row per step, are simply never used.  dH needs one 3-double all-reduce per Hamiltonian.
     The reference has no such path; results are checked against the undecomposed chain.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._runtime import WS, require_cuda, stream_ptr


# ----------------------------------------------------------------------------- chain sharding
def shard_chains(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of `total` chains: returns (global id of first local chain, local count).
    Earlier ranks take the remainder, so counts differ by at most one."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def gather_chain_values(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather per-chain values ([..., B_local] on every rank) into [..., total], ordered by
    global chain id.  Works for uneven shards (pads to the largest shard)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [shard_chains(total, world, r)[1] for r in range(world)]
    mx = max(counts)
    pad = torch.zeros(local.shape[:-1] + (mx,), dtype=local.dtype, device=local.device)
    pad[..., : local.shape[-1]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[..., :c] for b, c in zip(bufs, counts)], dim=-1)


# ----------------------------------------------------------------------------- halo exchange
class HaloExchanger:
    """Ring exchange of k boundary rows of an extended slab tensor ext[..., Lx + 2k, L]:
    rows [k, 2k) go to the upper neighbour (rank-1) and become its bottom halo; rows [Lx, Lx+k)
    go to the lower neighbour (rank+1) and become its top halo.  Backend-agnostic (nccl on the
    GPUs, gloo in the CPU tests); world_size 1 degenerates to local periodic copies."""

    def __init__(self, Lx: int, k: int, group=None):
        self.Lx, self.k, self.group = Lx, k, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.up = (self.rank - 1) % self.world
        self.dn = (self.rank + 1) % self.world
        self.n_exchanges = 0
        self.bytes_sent = 0

    def exchange(self, ext: torch.Tensor) -> None:
        k, Lx = self.k, self.Lx
        first = ext[..., k:2 * k, :]                  # my first k own rows
        last = ext[..., Lx:Lx + k, :]                 # my last k own rows
        top, bot = ext[..., 0:k, :], ext[..., Lx + k:Lx + 2 * k, :]
        self.n_exchanges += 1
        if self.world == 1:
            top.copy_(last)
            bot.copy_(first)
            return
        s_up, s_dn = first.contiguous(), last.contiguous()
        r_top, r_bot = torch.empty_like(s_dn), torch.empty_like(s_up)
        # order matters when up == dn (world 2): the peer's sends arrive as (its first, its last)
        ops = [dist.P2POp(dist.isend, s_up, self.up, self.group),
               dist.P2POp(dist.isend, s_dn, self.dn, self.group),
               dist.P2POp(dist.irecv, r_bot, self.dn, self.group),
               dist.P2POp(dist.irecv, r_top, self.up, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        top.copy_(r_top)
        bot.copy_(r_bot)
        self.bytes_sent += (s_up.numel() + s_dn.numel()) * s_up.element_size()


# ----------------------------------------------------------------------------- slab HMC
class SlabHMC:
    """Plain HMC (hmc_u1.py:82-97 semantics) of ONE L x L lattice decomposed into `world` slabs of
    Lx = L/world rows.  Momenta are keyed by the global site index and the accept draw by
    (seed, chain, traj) only, so every rank takes the same Metropolis decision and the chain is
    identical (up to reduction round-off in dH) to the undecomposed one."""

    def __init__(self, lattice_size: int, beta: float, n_steps: int, step_size: float, halo: Optional[int] = None,
                 seed: int = 1331, chain: int = 0, group=None, device=None):
        self.L, self.beta, self.n_steps, self.dt = int(lattice_size), float(beta), int(n_steps), float(step_size)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.L % self.world or self.L % 2:
            raise ValueError("L must be even and divisible by the number of ranks")
        self.Lx = self.L // self.world
        if halo is None:       # auto: trade redundant rows (2k/Lx extra compute) for fewer exchanges (one per k steps)
            halo = max(10, self.Lx // 16)
        self.k = max(1, min(int(halo), self.Lx, self.n_steps))
        self.x0 = self.rank * self.Lx
        self.seed, self.chain = int(seed), int(chain)
        self.traj_counter = 0
        self.gpu_launches = 0
        self._lib = _lib.load()
        self._dev = require_cuda(device)
        self.halo = HaloExchanger(self.Lx, self.k, group)
        Lxe = self.Lx + 2 * self.k
        # ext[f][mu][row][y], f = 0 theta, 1 pi ; two buffers for the out-of-place step kernel
        self._ext = [torch.zeros(2, 2, Lxe, self.L, dtype=torch.float64, device=self._dev) for _ in range(2)]
        self._sums6 = torch.zeros(6, dtype=torch.float64, device=self._dev)     # [old sums | new sums]
        # CUDA-graph replay of the trajectory: on by default for a single rank; with NCCL halo exchanges inside
        # the capture (world > 1) it is opt-in (U1_SLAB_GRAPH=1) - a 2-GPU run with it hung in round 1.
        self._graph_mode = os.environ.get("U1_SLAB_GRAPH", "1" if self.world == 1 else "0") == "1"
        self._graph = None
        self._eager_calls = 0
        self.graph_error: Optional[str] = None

    def initialize(self) -> torch.Tensor:
        return torch.zeros(2, self.Lx, self.L, dtype=torch.float64, device=self._dev)

    # -- pieces ---------------------------------------------------------------------------
    def _call(self, status, what):
        _lib.check(status, what)
        self.gpu_launches += _lib.last_launch_count()

    def trajectory(self, theta_own: torch.Tensor, u: Optional[float] = None) -> Dict[str, torch.Tensor]:
        """One Metropolis step.  theta_own [2,Lx,L] (this rank's rows, contiguous fp64 on the device) is
        updated in place.  Returns 1-element device tensors dH, accept, H, plaq, topo (no host sync).

        After two eager trajectories the whole step (about 25 kernel launches, the halo exchanges and the
        all-reduce) is captured once into a CUDA graph and replayed: the trajectory index lives in a
        device counter that the graph increments (u1_momenta_rows_dev / u1_metropolis_sums_dev), so the
        replay draws fresh momenta and a fresh accept uniform.  `U1_SLAB_GRAPH=0`, an injected `u`, or a
        failed capture keep the eager path; with more than one rank the graph path is opt-in
        (`U1_SLAB_GRAPH=1`): capturing the NCCL send/recv of the halo exchange is not validated."""
        if self._graph_mode and u is None and theta_own.is_cuda:
            key = (theta_own.data_ptr(), tuple(theta_own.shape))
            g = self._graph
            if g is not None and g["key"] == key:
                return self._replay(g)
            if self._eager_calls >= 2 or g is not None:
                g = self._capture(theta_own, key)
                if g is not None:
                    return self._replay(g)
        self._eager_calls += 1
        return self._trajectory_body(theta_own, u, self.traj_counter, None)

    def _replay(self, g) -> Dict[str, torch.Tensor]:
        g["graph"].replay()
        self.traj_counter += 1
        self.gpu_launches += g["launches"]
        self.halo.n_exchanges += g["exchanges"]
        self.halo.bytes_sent += g["bytes"]
        return {k: v.clone() for k, v in g["out"].items()}

    def _capture(self, theta_own: torch.Tensor, key):
        dev = self._dev
        try:
            with torch.cuda.device(dev):
                ctr = torch.zeros(1, dtype=torch.int64, device=dev)      # trajectories since capture
                l0, e0, b0, t0 = self.gpu_launches, self.halo.n_exchanges, self.halo.bytes_sent, self.traj_counter
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    out = self._trajectory_body(theta_own, None, t0, ctr)
                    ctr += 1
            g = {"key": key, "graph": graph, "out": out, "ctr": ctr, "launches": self.gpu_launches - l0,
                 "exchanges": self.halo.n_exchanges - e0, "bytes": self.halo.bytes_sent - b0}
            # capture executes nothing: undo the host-side bookkeeping of the captured pass
            self.gpu_launches, self.halo.n_exchanges, self.halo.bytes_sent, self.traj_counter = l0, e0, b0, t0
            self._graph = g
            return g
        except Exception as e:  # noqa: BLE001 - any capture failure falls back to the eager path for good
            self._graph_mode = False
            self._graph = None
            self.graph_error = f"{type(e).__name__}: {e}"
            return None

    def _trajectory_body(self, theta_own: torch.Tensor, u: Optional[float], traj: int,
                         ctr: Optional[torch.Tensor]) -> Dict[str, torch.Tensor]:
        """All device work of one trajectory, operating in place on the halo-extended buffers (no packing
        copies apart from the halo send buffers): 1 copy-in, momenta, theta halo exchange, H_old sums, blocks
        of <= k fused steps with a (theta,pi) exchange in between, final drift in place, theta halo exchange,
        H_new sums, ONE all-reduce of the six sums, Metropolis, select."""
        L, Lx, k, dt = self.L, self.Lx, self.k, self.dt
        lib, dev = self._lib, self._dev
        ctr_p = None if ctr is None else ctr.data_ptr()
        cur, nxt = self._ext
        own = slice(k, k + Lx)
        plane = (Lx + 2 * k) * L                             # doubles between the mu planes of an ext buffer
        n_own = Lx * L
        s6 = self._sums6
        with torch.cuda.device(dev):
            sp = stream_ptr(dev)
            ws, nb = WS.get(dev, 4096 + 24 * (Lx + 2), "slab")     # up to one partial triple per row
            cur[0, :, own] = theta_own
            # momenta for own rows AND halo rows (global-site keyed: no exchange needed for pi here)
            self._call(lib.u1_momenta_rows_dev(cur[1].data_ptr(), self.seed, self.chain, traj, ctr_p, self.x0 - k,
                                               Lx + 2 * k, L, L, sp), "u1_momenta_rows")
            self.halo.exchange(cur[0])                      # theta halos
            self._call(lib.u1_slab_sums_ext(cur[0, 0, k].data_ptr(), cur[1, 0, k].data_ptr(), plane,
                                            s6[0:3].data_ptr(), Lx, L, ws, nb, sp), "u1_slab_sums_ext")
            done = 0
            while done < self.n_steps:                       # blocks of <= k steps between halo exchanges
                if done > 0:
                    self.halo.exchange(cur)                  # theta and pi, k rows each side
                nblk = min(k, self.n_steps - done)
                in_b = C.c_int(0)
                self._call(lib.u1_leapfrog_steps_rect(cur[0].data_ptr(), cur[1].data_ptr(), nxt[0].data_ptr(),
                                                      nxt[1].data_ptr(), 1, Lx + 2 * k, L, self.beta,
                                                      0.5 * dt if done == 0 else dt, dt, nblk, C.byref(in_b), sp),
                           "u1_leapfrog_steps_rect")
                if in_b.value:
                    cur, nxt = nxt, cur
                done += nblk
            # final half drift + wrap on own rows, in place (hmc_u1.py:78-79), then H_new
            for mu in range(2):
                t = cur[0, mu, k].data_ptr()
                self._call(lib.u1_drift_wrap(t, cur[1, mu, k].data_ptr(), t, 0.5 * dt, n_own, sp), "u1_drift_wrap")
            self.halo.exchange(cur[0])
            self._call(lib.u1_slab_sums_ext(cur[0, 0, k].data_ptr(), cur[1, 0, k].data_ptr(), plane,
                                            s6[3:6].data_ptr(), Lx, L, ws, nb, sp), "u1_slab_sums_ext")
            if self.world > 1:
                dist.all_reduce(s6, group=self.group)
            # every rank takes the same decision from the all-reduced sums and the shared Philox draw
            out = {n: torch.empty(1, dtype=torch.float64, device=dev) for n in ("dH", "H", "plaq", "topo")}
            out["accept"] = torch.empty(1, dtype=torch.int32, device=dev)
            u_d = None if u is None else torch.tensor([float(u)], dtype=torch.float64, device=dev)
            self._call(lib.u1_metropolis_sums_dev(s6[0:3].data_ptr(), s6[3:6].data_ptr(),
                                                  None if u_d is None else u_d.data_ptr(),
                                                  self.seed, self.chain, traj, ctr_p, 1, self.beta, float(L) * L,
                                                  out["dH"].data_ptr(), out["accept"].data_ptr(), out["H"].data_ptr(),
                                                  out["plaq"].data_ptr(), out["topo"].data_ptr(), sp),
                       "u1_metropolis_sums")
            for mu in range(2):
                self._call(lib.u1_select_accepted(theta_own[mu].data_ptr(), cur[0, mu, k].data_ptr(),
                                                  out["accept"].data_ptr(), n_own, 1, sp), "u1_select_accepted")
        if ctr is None:
            # (under capture the buffer roles must stay as they were: every replay starts from them)
            self._ext = [cur, nxt]
            self.traj_counter += 1
        return out
