"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

(i)  Chain sharding (BASELINE configs[1], [3]): independent chains are split over ranks; there is no
     data-path collective — Philox streams are keyed by the GLOBAL chain id, so the ensemble is the
     same for any GPU count; only per-chain observables are gathered.
(ii) Domain decomposition of one large lattice (configs[4], L=4096): 1-D slabs along x.  Each rank
     keeps an extended slab (own rows + k halo rows per side).  Halos of (theta, pi) are exchanged
     with the two ring neighbours once every k leapfrog steps; between exchanges the extended slab is
     advanced as a whole and the outer rows, which become stale one row per step, are simply never
     used.  Transport "peer" (default on GPUs): our own kernels store the edge rows straight into the
     neighbours' memory over NVLink and synchronise through epoch flags (csrc/u1_slab.cu), the
     Hamiltonian sums are all-gathered the same way, so a trajectory contains no NCCL call and is
     replayed as ONE CUDA graph per rank.  Transport "nccl": send/recv + all-reduce through
     torch.distributed (fallback; also what the gloo CPU tests exercise).
     The reference has no such path; results are checked against the undecomposed chain.
"""
from __future__ import annotations

import ctypes as C
import os
import math
import weakref
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


# ----------------------------------------------------------------------------- NVLink peer regions
class PeerRegions:
    """Table of the peer-region device pointers of all ranks of one slab decomposition (csrc/u1_slab.cu).

    PeerRegions.ipc(k, L, group): one process per GPU - every rank allocates its region with u1_peer_alloc,
    the CUDA IPC handles are all-gathered through torch.distributed (plumbing only) and opened with
    u1_peer_open; afterwards the data path never touches torch.distributed again.
    PeerRegions.local(world, k, L): `world` regions in THIS process on the current device (virtual ranks on
    one GPU for tests, and the degenerate world == 1 case)."""

    def __init__(self, lib, rank: int, world: int, ptrs, owned, opened):
        self._lib, self.rank, self.world = lib, rank, world
        self.ptrs = list(ptrs)                        # device pointers, index = rank
        self._owned, self._opened = list(owned), list(opened)
        self.table = (C.c_void_p * world)(*[C.c_void_p(p) for p in self.ptrs])

    @classmethod
    def local(cls, world: int, k: int, L: int, device=None):
        lib = _lib.load()
        dev = require_cuda(device)
        nbytes = lib.u1_slab_peer_region_bytes(int(k), int(L))
        ptrs = []
        with torch.cuda.device(dev):
            for _ in range(world):
                p = C.c_void_p()
                _lib.check(lib.u1_peer_alloc(C.byref(p), nbytes, None), "u1_peer_alloc")
                ptrs.append(p.value)
        shared = {"refs": world, "ptrs": ptrs}
        out = []
        for r in range(world):
            pr = cls(lib, r, world, ptrs, [], [])
            pr._shared = shared
            out.append(pr)
        return out

    @classmethod
    def ipc(cls, k: int, L: int, group=None, device=None):
        lib = _lib.load()
        dev = require_cuda(device)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        nbytes = lib.u1_slab_peer_region_bytes(int(k), int(L))
        hb = lib.u1_ipc_handle_bytes()
        handle = (C.c_ubyte * hb)()
        mine = C.c_void_p()
        err = None
        with torch.cuda.device(dev):
            st = lib.u1_peer_alloc(C.byref(mine), nbytes, handle)
            if st != 0:
                err = f"u1_peer_alloc: status {st} ({lib.u1_status_string(st).decode()})"
            handles = [None] * world
            dist.all_gather_object(handles, None if err else bytes(handle), group=group)
            ptrs, opened = [], []
            for r in range(world):
                if r == rank:
                    ptrs.append(mine.value)
                    continue
                if err or handles[r] is None:
                    err = err or f"rank {r} could not export its peer region"
                    ptrs.append(None)
                    continue
                p = C.c_void_p()
                buf = (C.c_ubyte * hb).from_buffer_copy(handles[r])
                st = lib.u1_peer_open(buf, C.byref(p))
                if st != 0:
                    err = f"u1_peer_open (CUDA IPC) of rank {r}: status {st} ({lib.u1_status_string(st).decode()})"
                    ptrs.append(None)
                    continue
                ptrs.append(p.value)
                opened.append(p.value)
            # every rank learns whether ALL ranks succeeded, so a failure raises everywhere instead of hanging some
            errs = [None] * world
            dist.all_gather_object(errs, err, group=group)
        if any(errs):
            for p in opened:
                lib.u1_peer_close(p)
            if mine.value:
                lib.u1_peer_free(mine.value)
            raise _lib.U1Error("NVLink peer regions unavailable: " + "; ".join(f"rank {r}: {e}" for r, e in enumerate(errs) if e))
        return cls(lib, rank, world, ptrs, [mine.value], opened)

    def status(self):
        """Control words of the own region (host sync): dict with the exchange / Metropolis epochs and `error`."""
        w = (C.c_ulonglong * 8)()
        _lib.check(self._lib.u1_slab_peer_status(self.ptrs[self.rank], w), "u1_slab_peer_status")
        return {"exchange_epoch": int(w[2]), "metropolis_epoch": int(w[5]), "error": int(w[6])}

    def close(self):
        for p in self._opened:
            self._lib.u1_peer_close(p)
        for p in self._owned:
            self._lib.u1_peer_free(p)
        self._opened, self._owned = [], []
        sh = getattr(self, "_shared", None)
        if sh is not None:
            sh["refs"] -= 1
            if sh["refs"] == 0:
                for p in sh["ptrs"]:
                    self._lib.u1_peer_free(p)
            self._shared = None


# ----------------------------------------------------------------------------- slab HMC
class SlabHMC:
    """Plain HMC (hmc_u1.py:82-97 semantics) of ONE L x L lattice decomposed into `world` slabs of
    Lx = L/world rows.  Momenta are keyed by the global site index and the accept draw by
    (seed, chain, traj) only, and every rank adds the Hamiltonian sums of all ranks in the same order,
    so every rank takes the same Metropolis decision and the chain is identical (up to reduction
    round-off in dH) to the undecomposed one.

    transport: "peer" (NVLink peer-memory kernels, whole trajectory in one CUDA graph; default on CUDA),
    "nccl" (torch.distributed send/recv + all-reduce).  `peers` / `rank` / `world` let a test run several
    virtual ranks inside one process (PeerRegions.local)."""

    def __init__(self, lattice_size: int, beta: float, n_steps: int, step_size: float, halo: Optional[int] = None,
                 seed: int = 1331, chain: int = 0, group=None, device=None, transport: Optional[str] = None,
                 peers: Optional[PeerRegions] = None, rank: Optional[int] = None, world: Optional[int] = None):
        self.L, self.beta, self.n_steps, self.dt = int(lattice_size), float(beta), int(n_steps), float(step_size)
        self.group = group
        if peers is not None:
            self.world, self.rank = peers.world, peers.rank
        else:
            self.world = int(world) if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
            self.rank = int(rank) if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        if self.L % self.world or self.L % 2:
            raise ValueError("L must be even and divisible by the number of ranks")
        self.Lx = self.L // self.world
        self.transport = transport or os.environ.get("U1_SLAB_TRANSPORT", "peer")
        if self.transport not in ("peer", "nccl"):
            raise ValueError("transport must be 'peer' or 'nccl'")
        if halo is None:
            env = os.environ.get("U1_SLAB_HALO")
            halo = int(env) if env else (self._choose_halo() if self.transport == "peer" else max(10, self.Lx // 16))
        self.k = max(1, min(int(halo), self.Lx, self.n_steps))
        self.x0 = self.rank * self.Lx
        self.seed, self.chain = int(seed), int(chain)
        self.traj_counter = 0
        self.gpu_launches = 0
        self._lib = _lib.load()
        self._dev = require_cuda(device)
        self.halo = HaloExchanger(self.Lx, self.k, group)
        self.halo.world, self.halo.rank = self.world, self.rank
        self.halo.up, self.halo.dn = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        Lxe = self.Lx + 2 * self.k
        # ext[f][mu][row][y], f = 0 theta, 1 pi ; two work buffers for the out-of-place step kernels, and the state
        # buffer: _st[0] = the accepted theta (kept in extended layout between trajectories, so a trajectory starts
        # without a copy-in and the steps read it in place), _st[1] = the fresh momenta of the current trajectory
        self._ext = [torch.zeros(2, 2, Lxe, self.L, dtype=torch.float64, device=self._dev) for _ in range(2)]
        self._st = torch.zeros(2, 2, Lxe, self.L, dtype=torch.float64, device=self._dev)
        self._state_ref = None                     # weakref to the caller's theta_own the state was taken from
        self._state_version = -1
        # [old sums | new sums]; old[0:2] (plaquette sums of the state) are carried from trajectory to trajectory by the
        # select kernel, old[2] (sum pi^2) comes out of the momenta kernel: H_old costs no pass over the fields
        self._sums6 = torch.zeros(6, dtype=torch.float64, device=self._dev)
        self._sums_red = torch.zeros(6, dtype=torch.float64, device=self._dev)  # nccl transport: all-reduced copy
        self._sums_valid = False                   # old[0:2] belong to the current state
        # private scratch for the sums kernels: a captured graph bakes this pointer in, so it must not be a
        # shared grow-only cache entry.  _ws_m / _ws_d: single-launch reductions (zeroed ticket words)
        self._ws = torch.empty(4096 + 24 * (self.Lx + 2), dtype=torch.uint8, device=self._dev)
        nred = int(self._lib.u1_slab_reduce_workspace_bytes(Lxe, self.L))
        self._ws_m = torch.zeros(nred, dtype=torch.uint8, device=self._dev)
        self._ws_d = torch.zeros(nred, dtype=torch.uint8, device=self._dev)
        self.peers: Optional[PeerRegions] = None
        self.peer_error: Optional[str] = None
        with torch.cuda.device(self._dev):
            _lib.check(self._lib.u1_slab_preload(), "u1_slab_preload")      # no lazy code load behind a spinning kernel
        if self.transport == "peer":
            try:
                if peers is not None:
                    self.peers = peers
                elif self.world == 1:
                    self.peers = PeerRegions.local(1, self.k, self.L, self._dev)[0]
                else:
                    self.peers = PeerRegions.ipc(self.k, self.L, group, self._dev)
            except _lib.U1Error as e:            # e.g. CUDA IPC not permitted in this container (raised on every rank)
                self.peer_error = str(e)
                raise
        # CUDA-graph replay of the whole trajectory.  peer: on by default for any world size (no NCCL inside);
        # nccl: single rank only (capturing NCCL send/recv hung at 2 GPUs in round 1), opt-in otherwise.
        default_graph = "1" if (self.transport == "peer" or self.world == 1) else "0"
        self._graph_mode = os.environ.get("U1_SLAB_GRAPH", default_graph) == "1"
        self._graph = None
        self._eager_calls = 0
        self.graph_error: Optional[str] = None
        # peer transport: exchanges run on a side stream while the main stream advances the tile rows that do not
        # touch halo rows (U1_SLAB_OVERLAP=0: everything on one stream)
        self._overlap = self.transport == "peer" and os.environ.get("U1_SLAB_OVERLAP", "1") == "1"
        self._side = torch.cuda.Stream(device=self._dev) if self._overlap else None
        Lxe = self.Lx + 2 * self.k
        T = int(self._lib.u1_leapfrog_multi_tile_rows(Lxe))
        inner = [i for i in range(T) if 56 * i - 4 >= self.k and 56 * i + 60 <= self.Lx + self.k]
        multi_ok = self.L % 4 == 0 and Lxe >= 64 and self.L >= 64 and os.environ.get("U1_MULTISTEP", "1") != "0"
        # tile rows [i0, i1) read no halo row; [0, i0) and [i1, T) wait for the exchange
        self._split = (inner[0], inner[-1] + 1, T) if (self._overlap and multi_ok and len(inner) >= 2) else None

    def _choose_halo(self) -> int:
        """Halo width for the peer transport from a two-term cost model fitted to the 8-GPU sweep of round 2
        (profiles/r02_slab_8gpu_halo_sweep.json): the temporally blocked kernel works on whole 56-row tile rows, so a
        trajectory costs (launches) x ceil((Lx + 2k)/56) tile rows, plus ~2 tile-row-launches worth of time per exchange
        (an exchange is one ~10 us push/unpack kernel).  Multiples of 4 (steps per launch) only."""
        best, best_cost = 4, float("inf")
        for k in range(4, 33, 4):
            if k > self.Lx or k > max(4, self.n_steps):
                break
            blocks = [min(k, self.n_steps - d) for d in range(0, self.n_steps, k)]
            launches = sum(-(-b // 4) for b in blocks)
            tile_rows = -(-(self.Lx + 2 * k) // 56)
            cost = launches * tile_rows + 2.0 * (len(blocks) + 1)
            if cost < best_cost - 1e-9:
                best, best_cost = k, cost
        return best

    def initialize(self) -> torch.Tensor:
        return torch.zeros(2, self.Lx, self.L, dtype=torch.float64, device=self._dev)

    def close(self):
        self._graph = None
        if self.peers is not None:
            self.peers.close()
            self.peers = None

    def check_peers(self) -> None:
        """Host sync + read of the own peer region's error word; raises if a bounded wait on a peer timed out."""
        if self.peers is None:
            return
        torch.cuda.synchronize(self._dev)
        st = self.peers.status()
        if st["error"]:
            raise _lib.U1Error(f"slab peer exchange timed out (error word {st['error']}, rank {self.rank})")

    # -- pieces ---------------------------------------------------------------------------
    def _call(self, status, what):
        _lib.check(status, what)
        self.gpu_launches += _lib.last_launch_count()

    def trajectory(self, theta_own: torch.Tensor, u: Optional[float] = None) -> Dict[str, torch.Tensor]:
        """One Metropolis step.  theta_own [2,Lx,L] (this rank's rows, contiguous fp64 on the device) is
        updated in place.  Returns 1-element device tensors dH, accept, H, plaq, topo (no host sync).

        After two eager trajectories the whole step (about 30 kernel launches including the halo exchanges
        and the cross-rank Metropolis decision) is captured once into a CUDA graph and replayed: the
        trajectory index lives in a device counter that the graph increments (u1_momenta_rows_dev /
        u1_slab_metropolis_peer) and the exchange epochs in the peer region, so the replay draws fresh momenta
        and a fresh accept uniform.  `U1_SLAB_GRAPH=0`, an injected `u`, or a failed capture keep the eager
        path.  Every rank must call trajectory() the same number of times.

        The accepted state also lives in an extended-layout buffer of this object; theta_own is copied in only when
        it is a different tensor than last time or torch has seen an in-place write to it since (tensor version
        counter).  After writing to theta_own through a raw pointer call invalidate_state()."""
        self._sync_state(theta_own)
        if self._graph_mode and u is None and theta_own.is_cuda and self._sums_valid:
            key = (theta_own.data_ptr(), tuple(theta_own.shape), torch.cuda.current_stream(self._dev).cuda_stream)
            g = self._graph
            if g is not None and g["key"] == key:
                return self._replay(g)
            if self._eager_calls >= 2 or g is not None:
                g = self._capture(theta_own, key)
                if g is not None:
                    return self._replay(g)
        self._eager_calls += 1
        if self._graph is not None:
            self._graph["ctr"].add_(1)            # an eager trajectory between replays: keep the graph's device counter in step
        return self._trajectory_body(theta_own, u, self.traj_counter, None)

    def invalidate_state(self) -> None:
        """Forget the internal copy of the state: the next trajectory() copies theta_own in again."""
        self._state_ref = None
        self._sums_valid = False

    def _sync_state(self, theta_own: torch.Tensor) -> None:
        ref = self._state_ref() if self._state_ref is not None else None
        if ref is theta_own and theta_own._version == self._state_version:
            return
        k, Lx, L = self.k, self.Lx, self.L
        with torch.cuda.device(self._dev):
            self._call(self._lib.u1_slab_copy_in(theta_own.data_ptr(), self._st[0, 0, k].data_ptr(), (Lx + 2 * k) * L, Lx, L,
                                                 stream_ptr(self._dev)), "u1_slab_copy_in")
        self._state_ref = weakref.ref(theta_own)
        self._state_version = theta_own._version
        self._sums_valid = False                   # the next trajectory (eager) sums the new state's plaquettes itself

    def prepare_graph(self, theta_own: torch.Tensor) -> bool:
        """Capture the trajectory graph now (no replay).  trajectory() does this by itself on its third call; a
        caller that drives SEVERAL slabs from one process (virtual ranks on one GPU) must capture all of them
        before replaying any, because torch's capture synchronizes the device and a replay waits for its peers."""
        if not (self._graph_mode and theta_own.is_cuda):
            return False
        key = (theta_own.data_ptr(), tuple(theta_own.shape), torch.cuda.current_stream(self._dev).cuda_stream)
        if self._graph is None or self._graph["key"] != key:
            self._capture(theta_own, key)
        return self._graph is not None

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
                torch.cuda.current_stream(dev).synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=None, capture_error_mode="thread_local"):
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

    def _exchange(self, cur: torch.Tensor, with_pi: bool, rows_up: int, rows_dn: int, sp: int) -> None:
        """Halo exchange of the extended buffer `cur` [2 fields][2][Lx+2k][L]."""
        k, Lx, L = self.k, self.Lx, self.L
        if self.transport == "peer":
            p = self.peers
            self._call(self._lib.u1_slab_exchange(cur[0].data_ptr(), cur[1].data_ptr() if with_pi else None, Lx, k, L,
                                                  rows_up, rows_dn, p.ptrs[p.rank], p.ptrs[(p.rank - 1) % p.world],
                                                  p.ptrs[(p.rank + 1) % p.world], sp), "u1_slab_exchange")
            self.halo.n_exchanges += 1
            self.halo.bytes_sent += (4 if with_pi else 2) * (rows_up + rows_dn) * L * 8
        else:
            self.halo.exchange(cur if with_pi else cur[0])

    def _trajectory_body(self, theta_own: torch.Tensor, u: Optional[float], traj: int,
                         ctr: Optional[torch.Tensor]) -> Dict[str, torch.Tensor]:
        """All device work of one trajectory on the halo-extended buffers (the state is already in self._st[0]):
        momenta, theta halo exchange, H_old sums, blocks of <= k fused steps with a (theta,pi) exchange in between
        (the first block reads the state in place), one-row (theta,pi) exchange, final drift fused into the H_new
        sums, cross-rank sum + Metropolis, select into the state and the caller's theta_own."""
        L, Lx, k, dt = self.L, self.Lx, self.k, self.dt
        lib, dev = self._lib, self._dev
        ctr_p = None if ctr is None else ctr.data_ptr()
        cur, nxt = self._ext
        st = self._st
        plane = (Lx + 2 * k) * L                             # doubles between the mu planes of an ext buffer
        s6 = self._sums6
        peer = self.transport == "peer"
        with torch.cuda.device(dev):
            sp = stream_ptr(dev)
            ws, nb = self._ws.data_ptr(), self._ws.numel()
            # theta halo exchange on the side stream, concurrent with the momenta kernel (it only writes pi)
            main = torch.cuda.current_stream(dev)
            side = self._side if self._overlap else None
            if side is not None:
                side.wait_stream(main)
                self._exchange(st, False, k, k, side.cuda_stream)
            # momenta for own rows AND halo rows (global-site keyed: no exchange needed for pi here); sum pi^2 of the
            # own rows -> s6[2]
            self._call(lib.u1_momenta_rows_pi2_dev(st[1].data_ptr(), self.seed, self.chain, traj, ctr_p, self.x0 - k,
                                                   Lx + 2 * k, L, L, k, Lx, s6[2:3].data_ptr(), self._ws_m.data_ptr(),
                                                   self._ws_m.numel(), sp), "u1_momenta_rows_pi2")
            if side is not None:
                main.wait_stream(side)
            else:
                self._exchange(st, False, k, k, sp)          # theta halos
            if not self._sums_valid:
                # first trajectory after a copy-in: plaquette sums of the state (afterwards the select kernel carries them)
                if ctr is not None:
                    raise _lib.U1Error("slab graph capture needs valid state sums (run one eager trajectory first)")
                tmp = torch.empty(3, dtype=torch.float64, device=dev)
                self._call(lib.u1_slab_sums_ext(st[0, 0, k].data_ptr(), None, plane, tmp.data_ptr(), Lx, L, ws, nb, sp),
                           "u1_slab_sums_ext")
                s6[0:2].copy_(tmp[0:2])
                self._sums_valid = True
            done = 0
            while done < self.n_steps:                       # blocks of <= k steps between halo exchanges
                nblk = min(k, self.n_steps - done)
                first = 0                                    # steps of this block already taken by the split launch
                if done > 0 and self._split is not None:
                    # (theta, pi) exchange on the side stream; meanwhile the first <= 4 steps of the interior tile
                    # rows (their 64-row windows read no halo row), then the boundary tile rows once the halos are in
                    i0, i1, T = self._split
                    first = min(4, nblk)
                    side.wait_stream(main)
                    self._exchange(cur, True, k, k, side.cuda_stream)
                    args = (cur[0].data_ptr(), cur[1].data_ptr(), nxt[0].data_ptr(), nxt[1].data_ptr(), Lx + 2 * k, L,
                            self.beta, dt, dt, first)
                    self._call(lib.u1_leapfrog_multi_rows(*args, i0, i1 - i0, sp), "u1_leapfrog_multi_rows")
                    main.wait_stream(side)
                    if i0 > 0 and T > i1:                    # top and bottom boundary tile rows in one launch
                        self._call(lib.u1_leapfrog_multi_rows2(*args, 0, i0, i1, T - i1, sp), "u1_leapfrog_multi_rows2")
                    elif i0 > 0:
                        self._call(lib.u1_leapfrog_multi_rows(*args, 0, i0, sp), "u1_leapfrog_multi_rows")
                    elif T > i1:
                        self._call(lib.u1_leapfrog_multi_rows(*args, i1, T - i1, sp), "u1_leapfrog_multi_rows")
                    cur, nxt = nxt, cur
                elif done > 0:
                    self._exchange(cur, True, k, k, sp)      # theta and pi, k rows each side
                if done == 0:
                    in_b = C.c_int(0)                        # reads the state in place, first output into cur
                    self._call(lib.u1_leapfrog_steps_rect3(st[0].data_ptr(), st[1].data_ptr(), cur[0].data_ptr(),
                                                           cur[1].data_ptr(), nxt[0].data_ptr(), nxt[1].data_ptr(), 1,
                                                           Lx + 2 * k, L, self.beta, 0.5 * dt, dt, nblk, C.byref(in_b), sp),
                               "u1_leapfrog_steps_rect3")
                    if in_b.value:
                        cur, nxt = nxt, cur
                elif nblk > first:
                    in_b = C.c_int(0)
                    self._call(lib.u1_leapfrog_steps_rect(cur[0].data_ptr(), cur[1].data_ptr(), nxt[0].data_ptr(),
                                                          nxt[1].data_ptr(), 1, Lx + 2 * k, L, self.beta,
                                                          dt, dt, nblk - first, C.byref(in_b), sp),
                               "u1_leapfrog_steps_rect")
                    if in_b.value:
                        cur, nxt = nxt, cur
                done += nblk
            # final half drift + wrap (hmc_u1.py:78-79) fused into the H_new sums: the proposal theta' of the own rows
            # goes to nxt; the plaquettes of the last own row need (theta, pi) of the lower neighbour's first row
            if peer:
                self._exchange(cur, True, 1, 0, sp)
            else:
                self._exchange(cur, True, k, k, sp)
            self._call(lib.u1_slab_drift_sums_ext(cur[0, 0, k].data_ptr(), cur[1, 0, k].data_ptr(), plane,
                                                  nxt[0, 0, k].data_ptr(), 0.5 * dt, s6[3:6].data_ptr(), Lx, L,
                                                  self._ws_d.data_ptr(), self._ws_d.numel(), sp),
                       "u1_slab_drift_sums_ext")
            out = {n: torch.empty(1, dtype=torch.float64, device=dev) for n in ("dH", "H", "plaq", "topo")}
            out["accept"] = torch.empty(1, dtype=torch.int32, device=dev)
            u_d = None if u is None else torch.tensor([float(u)], dtype=torch.float64, device=dev)
            if peer:
                # all ranks' sums gathered through the peer regions and added in rank order inside the kernel
                self._call(lib.u1_slab_metropolis_peer(
                    s6[0:3].data_ptr(), s6[3:6].data_ptr(), None if u_d is None else u_d.data_ptr(), self.seed,
                    self.chain, traj, ctr_p, self.beta, float(L) * L, self.rank, self.world, self.peers.table,
                    out["dH"].data_ptr(), out["accept"].data_ptr(), out["H"].data_ptr(), out["plaq"].data_ptr(),
                    out["topo"].data_ptr(), sp), "u1_slab_metropolis_peer")
            else:
                sr = self._sums_red                           # the local sums stay intact: old[0:2] are carried on
                sr.copy_(s6)
                if self.world > 1:
                    dist.all_reduce(sr, group=self.group)
                self._call(lib.u1_metropolis_sums_dev(sr[0:3].data_ptr(), sr[3:6].data_ptr(),
                                                      None if u_d is None else u_d.data_ptr(),
                                                      self.seed, self.chain, traj, ctr_p, 1, self.beta, float(L) * L,
                                                      out["dH"].data_ptr(), out["accept"].data_ptr(), out["H"].data_ptr(),
                                                      out["plaq"].data_ptr(), out["topo"].data_ptr(), sp),
                           "u1_metropolis_sums")
            self._call(lib.u1_slab_select_state(st[0, 0, k].data_ptr(), nxt[0, 0, k].data_ptr(), plane, theta_own.data_ptr(),
                                                Lx, L, out["accept"].data_ptr(), s6[0:2].data_ptr(), s6[3:5].data_ptr(), sp),
                       "u1_slab_select_state")
        if ctr is None:
            self.traj_counter += 1
        return out
