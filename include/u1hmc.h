/* u1hmc.h — C ABI of the B200 (sm_100a) 2D U(1) HMC / FT-HMC hot path.
 *
 * The reference (Greyyy-HJC/hmc_ft) is pure Python/torch and has no FFI layer; the seam it
 * offers is the Python class surface (HMC_U1, HMC_U1_FT, FieldTransformation, utils.*).
 * Each entry point below replaces the torch-op sequence of one reference method and cites it
 * (paths relative to 2d_u1_cluster/).  hmc_ft_b200/ binds these with ctypes.
 *
 * Conventions
 *  - All field pointers are DEVICE pointers owned by the caller: contiguous fp64, row-major
 *    theta[b][mu][x][y], b<B chains, mu<2, x,y<L (y fastest) — the reference's [2,L,L] with a
 *    leading chain axis.  L must be even for the FT entry points (utils.py:21-115 masks).
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant and
 *    allocation-free: scratch comes in through (ws, ws_bytes); query with *_workspace_bytes().
 *  - Return value: 0 ok, <0 argument error (U1_ERR_*), >0 a cudaError_t.  Nothing throws.
 *  - No device-global / constant-memory state: FT weights live in an opaque per-handle host object and
 *    travel to the kernels as launch parameters; the handle's device table belongs to the device that was
 *    current at u1ft_create and every u1ft_* call must run on that device (U1_ERR_HANDLE otherwise).
 *    Host-side process state is limited to (a) tuning switches read once from the environment (U1_*,
 *    U1FT_*) plus u1ft_set_chunk_sites, and (b) per-DEVICE caches of occupancy / function attributes,
 *    indexed by the current device and updated atomically, so one process may drive several GPUs from
 *    several threads.
 *  - Field pointers must be 16-byte aligned (the kernels move double2 / TMA bulk copies): U1_ERR_ARG
 *    otherwise.  Batches of any size are accepted (launches are sliced at 65535 chains internally).
 */
#ifndef U1HMC_H
#define U1HMC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define U1_ABI_VERSION 1
#define U1_OK 0
#define U1_ERR_ARG (-1)         /* null pointer / non-positive size            */
#define U1_ERR_SHAPE (-2)       /* unsupported L (odd L on an FT entry point)  */
#define U1_ERR_WORKSPACE (-3)   /* workspace too small or misaligned           */
#define U1_ERR_HANDLE (-4)      /* bad FT handle / weight blob                 */
#define U1_ERR_UNSUPPORTED (-5)

typedef void* u1_stream_t;

int u1_abi_version(void);
const char* u1_status_string(int status);
/* SM count and compute capability of the current device (fails loudly off sm_100). */
int u1_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------- L1a lattice primitives (utils.py) ---------------- */
/* utils.py:118-127,143-150  plaq_from_field[_batch]: P[b][x][y]. */
int u1_plaquette(const double* theta, double* plaq, int B, int L, u1_stream_t stream);
/* utils.py:129-141  rect_from_field_batch: R[b][2][x][y]. */
int u1_rectangle(const double* theta, double* rect, int B, int L, u1_stream_t stream);
/* utils.py:182-187  regularize, elementwise over n doubles (in may equal out). */
int u1_regularize(const double* in, double* out, size_t n, u1_stream_t stream);
/* utils.py:173-180,189-196  plaq_mean_from_field / topo_from_field: per-chain outputs. */
int u1_observables(const double* theta, double* plaq_mean, double* topo, int B, int L,
                   void* ws, size_t ws_bytes, u1_stream_t stream);

/* utils.py:224-298  auto_from_chi / auto_by_def for B chains at once.  topo: [T][B] topological-charge
 * histories (rounded to the nearest integer like the reference).  gamma_chi[d][b] =
 * 1 - <(Q_t-Q_{t+d})^2> / two_v_chi with two_v_chi = 2*V*chi_inf(beta) (utils.py:207-222, a host
 * constant); gamma_def[d][b] = autocovariance / variance (all ones if the variance is zero).
 * Outputs are [max_lag+1][B]; either may be NULL.  ws: 2*B doubles. */
int u1_autocorr(const double* topo, int T, int B, int max_lag, double two_v_chi,
                double* gamma_chi, double* gamma_def, void* ws, size_t ws_bytes, u1_stream_t stream);

/* ---------------- HMC_U1 (hmc_u1.py) ---------------- */
size_t u1_workspace_bytes(int B, int L);
/* hmc_u1.py:56-63  action(theta) -> S[b]. */
int u1_action(const double* theta, double* S, int B, int L, double beta,
              void* ws, size_t ws_bytes, u1_stream_t stream);
/* hmc_u1.py:65-69  force(theta) (autograd of the action) -> F[b][2][x][y]. */
int u1_force(const double* theta, double* force, int B, int L, double beta, u1_stream_t stream);
/* hmc_u1.py:71-80  leapfrog(theta, pi): n_steps force evaluations, half drifts at both ends,
 * theta wrapped at the end.  One fused drift+force+kick kernel per step (HBM streaming).
 * In-place is allowed (theta_out==theta_in, pi_out==pi_in) only through the workspace path:
 * ws must hold u1_workspace_bytes(B,L). */
int u1_leapfrog(const double* theta_in, const double* pi_in, double* theta_out, double* pi_out,
                int B, int L, double beta, double dt, int n_steps,
                void* ws, size_t ws_bytes, u1_stream_t stream);
/* One leapfrog step as a single kernel: theta' = theta + drift*pi ; pi' = pi - kick*F(theta').
 * (The body of hmc_u1.py:75-77.)  Out-of-place only. */
int u1_leapfrog_step(const double* theta_in, const double* pi_in, double* theta_out, double* pi_out,
                     int B, int L, double beta, double drift, double kick, u1_stream_t stream);
/* Philox4x32-10 momenta pi[b] ~ N(0,1) for global chain ids chain0..chain0+B-1 and trajectory
 * index `traj` (replaces torch.randn_like, hmc_u1.py:83). */
int u1_momenta(double* pi, uint64_t seed, uint64_t chain0, uint64_t traj, int B, int L,
               u1_stream_t stream);
/* hmc_u1.py:82-97 metropolis_step, `n_traj` times back to back, plus the stored-step
 * observables of run() (hmc_u1.py:219-225).  theta is updated in place.
 *   pi_inject  [n_traj][B][2][L][L] or NULL (NULL: Philox momenta on device)
 *   u_inject   [n_traj][B]           or NULL (NULL: Philox uniforms on device)
 * Outputs are [n_traj][B]: dH, accept (int32 0/1), H (H_new if accepted else H_old, as the
 * reference returns), plaq and topo of the state after the step.  Any output may be NULL.
 * path: 0 auto, 1 force the streaming (one kernel per leapfrog step) path,
 *       2 force the SM-resident whole-trajectory kernel (L in {8,16,32,64}). */
int u1_trajectory(double* theta, const double* pi_inject, const double* u_inject,
                  uint64_t seed, uint64_t chain0, uint64_t traj0, int n_traj,
                  int B, int L, double beta, double dt, int n_steps,
                  double* out_dH, int32_t* out_accept, double* out_H, double* out_plaq,
                  double* out_topo, int path, void* ws, size_t ws_bytes, u1_stream_t stream);
/* Number of kernel launches the last u1_* / u1ft_* call on this host thread issued
 * (bench.py reports it as gpu_launches). */
long long u1_last_launch_count(void);

/* FP64-FMA throughput probe (the roofline denominator of the FT path, which MEASURED_PEAKS.json
 * does not hold): launches `blocks` CTAs of 256 threads, each thread issuing iters*32 dependent-
 * chain-free DFMAs.  out: device [blocks*256] doubles (sink).  flops = blocks*256*iters*32*2. */
int u1_fp64_probe(double* out, int blocks, int iters, u1_stream_t stream);

/* ---------------- slab (domain-decomposed) variants for one large lattice ----------------
 * The lattice is split along x into slabs of Lx rows (global width L).  The caller supplies the
 * halo rows it received from its ring neighbours (NCCL send/recv through torch.distributed):
 *   theta_up  [2][L]  : last row of the upper neighbour  (x = -1), values BEFORE this step's drift
 *   pi_up     [2][L]
 *   theta_dn  [2][L]  : first row of the lower neighbour (x = Lx)
 *   pi_dn     [2][L]
 * Since drift is local, halos of (theta, pi) before the step are enough to advance one step. */
int u1_slab_leapfrog_step(const double* theta_in, const double* pi_in,
                          const double* theta_up, const double* pi_up,
                          const double* theta_dn, const double* pi_dn,
                          double* theta_out, double* pi_out,
                          int Lx, int L, double beta, double drift, double kick,
                          u1_stream_t stream);
/* Communication-avoiding variant: the caller keeps an extended slab of Lx rows (own rows plus k
 * halo rows on each side, refreshed every k steps) and advances it as a periodic Lx x L lattice;
 * rows within s of the slab edge are invalid after s steps and are never read by the caller.
 * Same kernel as u1_leapfrog_step with a rectangular lattice. */
int u1_leapfrog_step_rect(const double* theta_in, const double* pi_in, double* theta_out, double* pi_out,
                          int B, int Lx, int L, double beta, double drift, double kick,
                          u1_stream_t stream);
/* n consecutive leapfrog steps of a periodic Lx x L lattice in one call, ping-ponging between
 * (thetaA,piA) (the input) and (thetaB,piB).  The first step drifts by first_drift, the others by dt;
 * every kick is dt.  When L % 4 == 0 and Lx, L >= 64 the steps run in groups of up to four inside the
 * temporally blocked kernel (64 x 64 register-resident windows, one HBM pass per group; bit-identical
 * to single steps), otherwise one k_leapfrog_step launch per step.  *result_in_B (host int) tells
 * which buffer pair holds the result. */
int u1_leapfrog_steps_rect(double* thetaA, double* piA, double* thetaB, double* piB,
                           int B, int Lx, int L, double beta, double first_drift, double dt, int n,
                           int* result_in_B, u1_stream_t stream);
/* u1_leapfrog_steps_rect with a read-only input pair: the steps start from (theta_in, pi_in), which are left
 * untouched, and ping-pong between (thetaA, piA) (first output) and (thetaB, piB).  The slab path keeps the accepted
 * state in place this way (no copy into a work buffer at the start of a trajectory). */
int u1_leapfrog_steps_rect3(const double* theta_in, const double* pi_in, double* thetaA, double* piA, double* thetaB,
                            double* piB, int B, int Lx, int L, double beta, double first_drift, double dt, int n,
                            int* result_in_B, u1_stream_t stream);
/* u1_slab_sums_ext of theta' = regularize(theta + a*pi) (hmc_u1.py:78-79 fused into the H_new reduction): theta' of
 * the own rows is written to theta_out_own (same plane stride, a different buffer), sums[0..2] as u1_slab_sums_ext.
 * theta AND pi of the row after the last own row are read in place (exchange one (theta, pi) row first).
 * One launch: the block that finishes last adds the block partials in a fixed order.  ws: at least
 * u1_slab_reduce_workspace_bytes(Lx, L) bytes whose first 256 bytes the caller ZEROED ONCE (ticket word; every launch
 * leaves it zero again). */
int u1_slab_drift_sums_ext(const double* theta_own, const double* pi_own, size_t plane_stride, double* theta_out_own,
                           double a, double* sums, int Lx, int L, void* ws, size_t ws_bytes, u1_stream_t stream);
/* workspace of u1_slab_drift_sums_ext and u1_momenta_rows_pi2_dev for buffers of up to `rows` rows of L sites */
size_t u1_slab_reduce_workspace_bytes(int rows, int L);
/* One launch of the temporally blocked kernel restricted to tile rows [tile_row0, tile_row0 + tile_rows) of the
 * u1_leapfrog_multi_tile_rows(Lx) rows of 56-row tiles: ksteps <= 4 leapfrog steps (first drift first_drift, then
 * dt) from (theta_in, pi_in) to (theta_out, pi_out); a tile row reads rows [56 i - 4, 56 i + 60) (periodic) and
 * writes rows [56 i, 56 i + 56).  Lets the slab path advance the interior of a slab while the halo exchange is
 * still in flight and finish the boundary tile rows afterwards.  U1_ERR_UNSUPPORTED unless L % 4 == 0, Lx, L >= 64. */
int u1_leapfrog_multi_tile_rows(int Lx);
int u1_leapfrog_multi_rows(const double* theta_in, const double* pi_in, double* theta_out, double* pi_out,
                           int Lx, int L, double beta, double first_drift, double dt, int ksteps,
                           int tile_row0, int tile_rows, u1_stream_t stream);
/* Same for TWO disjoint ranges of tile rows in one launch: [tile_row0, +tile_rows) and [tile_row1, +tile_rows1) with
 * tile_row1 >= tile_row0 + tile_rows (the slab path finishes the top and the bottom boundary tile rows together). */
int u1_leapfrog_multi_rows2(const double* theta_in, const double* pi_in, double* theta_out, double* pi_out,
                            int Lx, int L, double beta, double first_drift, double dt, int ksteps,
                            int tile_row0, int tile_rows, int tile_row1, int tile_rows1, u1_stream_t stream);
/* Philox momenta for rows x_first .. x_first+Lx-1 (taken modulo Lx_global) of ONE global lattice of
 * Lx_global x L sites: pi[2][Lx][L].  Keyed by the global site index, so the field does not depend
 * on the decomposition. */
int u1_momenta_rows(double* pi, uint64_t seed, uint64_t chain, uint64_t traj,
                    long long x_first, int Lx, int L, int Lx_global, u1_stream_t stream);
/* Same, with the trajectory index taken as traj + *traj_dev (traj_dev: device pointer to a uint64 counter,
 * may be NULL): lets a captured CUDA graph of a slab trajectory be replayed with a counter that the graph
 * itself increments, instead of a launch parameter frozen at capture time. */
int u1_momenta_rows_dev(double* pi, uint64_t seed, uint64_t chain, uint64_t traj, const uint64_t* traj_dev,
                        long long x_first, int Lx, int L, int Lx_global, u1_stream_t stream);
/* u1_momenta_rows_dev that also returns pi2_out[0] = sum of pi^2 over the rows [own_first, own_first + own_rows) of
 * the buffer (the kinetic term of H_old, hmc_u1.py:86, without a second pass over pi).  ws as for
 * u1_slab_drift_sums_ext (its own workspace: the two calls may be in flight on different streams). */
int u1_momenta_rows_pi2_dev(double* pi, uint64_t seed, uint64_t chain, uint64_t traj, const uint64_t* traj_dev,
                            long long x_first, int Lx, int L, int Lx_global, int own_first, int own_rows, double* pi2_out,
                            void* ws, size_t ws_bytes, u1_stream_t stream);
/* out = regularize(theta + a*pi), elementwise over n doubles (last half drift, hmc_u1.py:78-79). */
int u1_drift_wrap(const double* theta, const double* pi, double* out, double a, size_t n,
                  u1_stream_t stream);
/* Metropolis decision from Hamiltonian sums (hmc_u1.py:84-97): sums_* are [B][3] =
 * {sum cos(wrap P), sum wrap P, sum pi^2}; H = -beta*s0 + s2/2; accept iff u < exp(-dH) with u
 * injected ([B]) or drawn from Philox(seed, chain0+b, traj).  out_accept (int32 [B]) is required,
 * the other outputs may be NULL.  plaq/topo are those of the state kept. */
int u1_metropolis_sums(const double* sums_old, const double* sums_new, const double* u_inject,
                       uint64_t seed, uint64_t chain0, uint64_t traj, int B, double beta, double volume,
                       double* out_dH, int32_t* out_accept, double* out_H, double* out_plaq,
                       double* out_topo, u1_stream_t stream);
/* Same, with the trajectory index taken as traj + *traj_dev (see u1_momenta_rows_dev). */
int u1_metropolis_sums_dev(const double* sums_old, const double* sums_new, const double* u_inject,
                           uint64_t seed, uint64_t chain0, uint64_t traj, const uint64_t* traj_dev, int B,
                           double beta, double volume, double* out_dH, int32_t* out_accept, double* out_H,
                           double* out_plaq, double* out_topo, u1_stream_t stream);
/* theta[b] = theta_new[b] for every chain with accept[b] != 0 (per_chain doubles per chain). */
int u1_select_accepted(double* theta, const double* theta_new, const int32_t* accept,
                       size_t per_chain, int B, u1_stream_t stream);
/* Partial sums of one slab: sums[0]=sum cos(wrap P), sums[1]=sum wrap P, sums[2]=sum pi^2
 * (pi may be NULL).  theta_dn as above (first row of the lower neighbour). */
int u1_slab_sums(const double* theta, const double* theta_dn, const double* pi, double* sums,
                 int Lx, int L, void* ws, size_t ws_bytes, u1_stream_t stream);
/* Same sums taken directly from a halo-extended slab buffer (no packing copies): theta_own / pi_own point
 * at the first OWN row of the mu=0 plane, the mu=1 plane lies plane_stride doubles further, and the row
 * after the last own row (the lower neighbour's first row, already exchanged) is read in place.
 * plane_stride >= (Lx+1)*L.  pi_own may be NULL. */
int u1_slab_sums_ext(const double* theta_own, const double* pi_own, size_t plane_stride, double* sums,
                     int Lx, int L, void* ws, size_t ws_bytes, u1_stream_t stream);

/* ---------------- NVLink peer path of the slab decomposition (csrc/u1_slab.cu) ----------------
 * The reference has no multi-GPU path; these entry points replace what a torch.distributed
 * implementation of SURVEY §8(e) would do with ncclSend/ncclRecv + ncclAllReduce per exchange: the kernels
 * store halo rows and Hamiltonian sums directly into the neighbours' memory over NVLink and synchronise
 * through epoch flags, so a trajectory needs no host-driven collective and can be captured in one CUDA graph.
 *
 * Peer region: u1_slab_peer_region_bytes(k, L) bytes of zero-initialised device memory per rank, allocated by
 * u1_peer_alloc (cudaMalloc; the only entry points that allocate - the memory must be IPC-exportable, which a
 * caching allocator's sub-blocks are not).  ipc_handle_out (may be NULL) receives u1_ipc_handle_bytes() bytes
 * that another process of the same box passes to u1_peer_open to map the region (cudaIpcOpenMemHandle, peer
 * access enabled lazily).  Within one process the pointers can be used directly. */
size_t u1_ipc_handle_bytes(void);
size_t u1_slab_peer_region_bytes(int k, int L);
int u1_peer_alloc(void** dev_ptr, size_t bytes, void* ipc_handle_out);
int u1_peer_open(const void* ipc_handle, void** dev_ptr);
int u1_peer_close(void* dev_ptr);
int u1_peer_free(void* dev_ptr);
/* Loads the code of every kernel a slab trajectory launches (CUDA loads kernels lazily at first launch, and such
 * a load may wait for a kernel that is spinning on a peer flag).  Call once per device before the first exchange. */
int u1_slab_preload(void);
int u1_plain_preload_slab_kernels(void);
/* First 8 control words of a region copied to the host (synchronous): [0] top flag, [1] bottom flag,
 * [2] exchange epoch, [5] Metropolis epoch, [6] error word (non-zero: a bounded wait on a peer timed out). */
int u1_slab_peer_status(const void* region_self, unsigned long long* host_words8);
/* Ring halo exchange of an extended slab ext[2][Lx+2k][L] (own rows k..k+Lx-1): my first rows_up own rows become
 * the bottom halo of rank-1 (region_up), my last rows_dn own rows the top halo of rank+1 (region_dn); the same
 * call on the neighbours fills my halos.  rows_* <= k, adjacent to the own rows.  ext_pi may be NULL (theta only).
 * One kernel launch: push over NVLink, st.release.sys epoch flag, bounded wait, unpack.  World size 1: pass the
 * own region three times (periodic wrap). */
int u1_slab_exchange(double* ext_theta, double* ext_pi, int Lx, int k, int L, int rows_up, int rows_dn,
                     void* region_self, void* region_up, void* region_dn, u1_stream_t stream);
/* hmc_u1.py:84-97 for a decomposed lattice: all-gather of every rank's Hamiltonian sums ([3] old, [3] new, as
 * u1_slab_sums_ext produces them) through the peer regions (regions: HOST array of `world` device pointers,
 * regions[rank] is the caller's own), summed in rank order so that every rank forms bit-identical totals, then
 * the Metropolis decision with the shared draw Philox(seed, chain, traj + *traj_dev) or u_inject[0].
 * Outputs are single elements; out_accept is required. */
int u1_slab_metropolis_peer(const double* sums_old, const double* sums_new, const double* u_inject,
                            uint64_t seed, uint64_t chain, uint64_t traj, const uint64_t* traj_dev,
                            double beta, double volume, int rank, int world, void* const* regions,
                            double* out_dH, int32_t* out_accept, double* out_H, double* out_plaq, double* out_topo,
                            u1_stream_t stream);
/* In-place helpers on the own rows of an extended buffer (both mu planes in one launch; *_own point at the first
 * own row of the mu=0 plane, the mu=1 plane lies plane_stride doubles further):
 *   drift_wrap: theta = regularize(theta + a*pi)   (hmc_u1.py:78-79)
 *   copy_in:    ext = theta_own [2][Lx][L]          select: theta_own = ext where *accept != 0 */
int u1_slab_drift_wrap(double* ext_theta_own, const double* ext_pi_own, size_t plane_stride, int Lx, int L, double a,
                       u1_stream_t stream);
int u1_slab_copy_in(const double* theta_own, double* ext_theta_own, size_t plane_stride, int Lx, int L, u1_stream_t stream);
/* select into a state kept in extended layout: state_own = src_own where *accept != 0 (both plane-strided), and the
 * caller's contiguous theta_own [2][Lx][L] as well when it is not NULL.  sums_state / sums_new (both or neither): on
 * accept sums_state[0..1] = sums_new[0..1], i.e. the plaquette sums of the kept state travel with it, so the next
 * trajectory's H_old needs no pass over theta. */
int u1_slab_select_state(double* state_own, const double* src_own, size_t plane_stride, double* theta_own, int Lx, int L,
                         const int32_t* accept, double* sums_state, const double* sums_new, u1_stream_t stream);
int u1_slab_select(double* theta_own, const double* ext_theta_own, size_t plane_stride, int Lx, int L,
                   const int32_t* accept, u1_stream_t stream);

/* Training-loss reduction (field_trans_opt.py:452-458): out4[k] = sum_i |a[i]-b[i]|^(2k+2), k = 0..3, over n
 * doubles (deterministic two-phase reduction; out4 is device memory).  The loss is
 * sum_k out4[k]^(1/p_k) / vol^(1/p_k), p_k = 2,4,6,8, formed by the caller. */
size_t u1_diff_power_sums_workspace_bytes(void);
int u1_diff_power_sums(const double* a, const double* b, size_t n, double* out4, void* ws, size_t ws_bytes,
                       u1_stream_t stream);

/* ---------------- FieldTransformation / HMC_U1_FT ---------------- */
typedef struct u1ft_handle_s* u1ft_handle_t;
/* field_trans_opt.py:654-692 (_load_best_model) + cnn_model_opt.py:6-247.
 * n_res_blocks: 0 = jointCNN_simple, 1 = jointCNN_rnet1, 3 = jointCNN_rnet3 (any 0..8 accepted).
 * The handle owns a 14 KB device table (GELU Taylor coefficients) on the current device.
 * weights: HOST fp64 blob, for subset s=0..7, for each conv layer in forward order:
 *   W[o][c][3][3] (torch Conv2d layout) then b[o].  First layer 6->12, the rest 12->12. */
int u1ft_create(u1ft_handle_t* out, int n_res_blocks, const double* weights, size_t n_doubles);
int u1ft_destroy(u1ft_handle_t h);
size_t u1ft_weight_count(int n_res_blocks);
/* Tunable: lattice sites (chains x L^2) processed per kernel launch by the FT entry points.  By default
 * (sites <= 0, or never called and no U1FT_CHUNK_SITES in the environment) the library uses as many sites
 * as fit a quarter of the device memory at the rnet3 workspace cost (~7.4 KB per site; all 1024 chains of
 * the L=64 benchmark in one 30 GB chunk on a 180 GB B200).  The chains are cut into equal chunks of at most
 * that many sites.  Larger chunks amortise kernel-boundary ramps; the workspace grows linearly.  Returns
 * the previous explicit value (0 = automatic), so passing the return value back restores the behaviour. */
long long u1ft_set_chunk_sites(long long sites);
/* u1ft_leapfrog / u1ft_trajectory replay cached CUDA graphs of their building blocks (one FT force + kick, one
 * forward + action pass; U1FT_GRAPH=0 disables).  u1_last_launch_count() keeps counting KERNELS; this returns the
 * number of launch API calls the host issued in the last u1ft_* call of this thread (a graph replay counts once). */
long long u1ft_last_host_launch_count(void);
/* Workspace for B chains.  with_backward!=0 sizes it for u1ft_force / u1ft_trajectory. */
size_t u1ft_workspace_bytes(u1ft_handle_t h, int B, int L, int with_backward);
/* field_trans_opt.py:219-243 forward / field_transformation and :284-383 compute_jac_logdet in
 * one pass (one CNN evaluation per subset).  theta_ori and logdet may each be NULL. */
int u1ft_forward(u1ft_handle_t h, const double* theta, double* theta_ori, double* logdet,
                 int B, int L, void* ws, size_t ws_bytes, u1_stream_t stream);
/* field_trans_opt.py:137-217  ft_phase(theta, index) -> Delta[b][2][x][y] (diagnostic surface). */
int u1ft_phase(u1ft_handle_t h, const double* theta, int index, double* delta,
               int B, int L, void* ws, size_t ws_bytes, u1_stream_t stream);
/* field_trans_opt.py:110-135  compute_K0_K1 of subset `index`: K[b][12][x][y] (channels 0..3 = K0,
 * 4..11 = K1).  Only the entries the transformation reads are defined: the active sites of the subset
 * and the 6 channels of its direction (the rest is multiplied by the field mask in ft_phase and is
 * left untouched here). */
int u1ft_coefficients(u1ft_handle_t h, const double* theta, int index, double* K,
                      int B, int L, void* ws, size_t ws_bytes, u1_stream_t stream);
/* One fixed-point iteration of field_trans_opt.py:245-282 (inverse) for subset `index` with the
 * coefficients K of u1ft_coefficients (they do not depend on the subset's own links, so one CNN
 * evaluation serves every iteration):  theta_next = theta_curr - ft_phase(theta_iter, index), and
 * norms[0] = |theta_next - theta_iter|^2, norms[1] = |theta_iter|^2 summed over the whole batch
 * (device doubles; the caller takes the square roots and tests the tolerance, as the reference
 * does on the host).  ws: u1ft_inverse_workspace_bytes(B, L). */
size_t u1ft_inverse_workspace_bytes(int B, int L);
int u1ft_inverse_iter(const double* theta_curr, const double* theta_iter, const double* K, int index,
                      double* theta_next, double* norms, int B, int L,
                      void* ws, size_t ws_bytes, u1_stream_t stream);
/* hmc_u1_ft.py:101-125  new_action = S(F(theta)) - logdet.  Optional extra outputs. */
int u1ft_action(u1ft_handle_t h, const double* theta, double* S_ft, double* theta_ori,
                double* logdet, int B, int L, double beta,
                void* ws, size_t ws_bytes, u1_stream_t stream);
/* hmc_u1_ft.py:128-152  new_force: hand-derived reverse pass (no autograd). S_ft may be NULL. */
int u1ft_force(u1ft_handle_t h, const double* theta, double* force, double* S_ft,
               int B, int L, double beta, void* ws, size_t ws_bytes, u1_stream_t stream);
/* hmc_u1_ft.py:154-179 leapfrog with new_force. */
int u1ft_leapfrog(u1ft_handle_t h, const double* theta_in, const double* pi_in,
                  double* theta_out, double* pi_out, int B, int L, double beta, double dt,
                  int n_steps, void* ws, size_t ws_bytes, u1_stream_t stream);
/* hmc_u1_ft.py:181-209 metropolis_step x n_traj + stored-step observables of run()
 * (hmc_u1_ft.py:348-358: plaq/topo of regularize(F(theta))).  Same conventions as u1_trajectory;
 * theta_ori_out [B][2][L][L] (may be NULL) receives F(theta) of the final state. */
int u1ft_trajectory(u1ft_handle_t h, double* theta, const double* pi_inject, const double* u_inject,
                    uint64_t seed, uint64_t chain0, uint64_t traj0, int n_traj,
                    int B, int L, double beta, double dt, int n_steps,
                    double* out_dH, int32_t* out_accept, double* out_H, double* out_plaq,
                    double* out_topo, double* theta_ori_out,
                    void* ws, size_t ws_bytes, u1_stream_t stream);

/* ---------------- training gradient of the field transformation (csrc/u1_ft_train.cuh) ----------------
 * field_trans_opt.py:443-480 train_step = loss_fn -> loss.backward() -> AdamW.  The reference gets d loss / d W from
 * autograd through the unrolled fixed-point inverse (:245-282) and through compute_force(create_graph=True)
 * (:400-441); here it is hand-derived.  The caller (hmc_ft_b200/field_trans.py loss_and_grad) composes:
 *   theta_n = inverse(theta_o), recording each stage's input and iteration count;  d = F_FT(theta_n) - F_plain(theta_n;1)
 *   v = u1_loss_grad(d);  (Hv, gradW) = u1ft_train_force_grad(theta_n, v);  ubar = Hv - u1_force_hvp(theta_n, v; 1)
 *   for the stages in reverse order of execution: ubar = u1ft_train_inverse_stage_bwd(stage, ubar)  (+= gradW)
 * gradW: device [u1ft_weight_count] doubles in the layout of the u1ft_create blob, ACCUMULATED (zero it first). */
size_t u1ft_train_workspace_bytes(u1ft_handle_t h, int B, int L);
/* Hessian-vector product of the plain action (hmc_u1.py:56-69 differentiated twice): out = d/d eps force(theta + eps w). */
int u1_force_hvp(const double* theta, const double* w, double* out, int B, int L, double beta, u1_stream_t stream);
/* v = d loss / d (a - b) for loss = sum_{p=2,4,6,8} ||a-b||_p / vol^(1/p) (field_trans_opt.py:452-458); sums4 as
 * produced by u1_diff_power_sums (device). */
int u1_loss_grad(const double* a, const double* b, const double* sums4, double vol, size_t n, double* v, u1_stream_t stream);
/* phi = v . F_FT(theta; beta, W):  Hv = d phi / d theta (Hessian-vector product of S_FT), gradW += d phi / d W.
 * force_check (may be NULL) receives d phi / d v, which must equal u1ft_force(theta) (self check).  phi: reserved, NULL. */
int u1ft_train_force_grad(u1ft_handle_t h, const double* theta, const double* v, int B, int L, double beta,
                          double* phi, double* Hv, double* force_check, double* gradW, size_t n_weights,
                          void* ws, size_t ws_bytes, u1_stream_t stream);
/* Adjoint of ONE stage of the inverse (subset `index`, n_iter fixed-point iterations from the stage input theta_curr;
 * n_iter = 0: the stage was skipped): ubar_out = d loss / d theta_curr given ubar_in = d loss / d (stage output);
 * gradW += contribution of the stage's CNN.  ubar_in != ubar_out. */
int u1ft_train_inverse_stage_bwd(u1ft_handle_t h, int index, const double* theta_curr, int n_iter, const double* ubar_in,
                                 double* ubar_out, double* gradW, size_t n_weights, int B, int L,
                                 void* ws, size_t ws_bytes, u1_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* U1HMC_H */
