/* b200nb.h — C ABI of libb200nb: the B200-native all-pairs gravity engine behind MUrB's `--im gpu+b200`.
 *
 * The reference (albtad01/NBody-EuroHPC) has no FFI; its seam is the C++ virtual class
 * SimulationNBodyInterface<T> (src/common/core/SimulationNBodyInterface.hpp:15-88) selected by the `--im` string in
 * createImplem() (src/murb/main.cpp:205-270).  The glue class SimulationNBodyB200 (nbody-eurohpc_b200/glue/) derives
 * from that interface and forwards to the functions below; see INTEGRATION.md for the one-branch patch a maintainer
 * adds to main.cpp.  Every entry point names the reference code it replaces.
 *
 * Conventions
 *  - plain C, plain pointers and sizes, no C++/torch types, no exceptions across the boundary;
 *  - return 0 (B200NB_OK) on success, a B200NB_E* code otherwise; b200nb_last_error() gives the text.  The library
 *    never calls exit() (the reference's CUDA_CHECK does: SimulationNBodyCUDATileFullDevice.cu:10-17) — the glue maps a
 *    failure onto the reference convention (print + exit);
 *  - host pointers are borrowed for the duration of the call; the context owns all device memory, streams, events
 *    and NCCL communicators; the caller's current CUDA device is restored before returning;
 *  - a context is used from one host thread at a time (the reference never calls an implementation from more than
 *    one thread); several contexts may exist one after another or side by side;
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with B200NB_ECUDA.
 *
 * Units follow the reference: SI, fp32 state, G = 6.67384e-11f (SimulationNBodyInterface.hpp:18).
 */
#ifndef B200NB_H_
#define B200NB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NB_VERSION 100

enum {
    B200NB_OK = 0,
    B200NB_EINVAL = 1, /* bad argument */
    B200NB_ECUDA = 2,  /* CUDA runtime / driver error, or no device */
    B200NB_ENCCL = 3,  /* NCCL missing or failed */
    B200NB_ESTATE = 4  /* call not valid in the current state (e.g. step before upload) */
};

/* Integrators.  0 is MUrB's scheme, Bodies<T>::updatePositionAndVelocity (src/common/core/Bodies.cpp:259-278) ==
 * devUpdatePositionsAndVelocities (src/common/core/CUDABodies.cu:125-153):  q += (v + a*dt/2)*dt ; v += a*dt.
 * 1 is the kick-drift-kick leapfrog the reference documents but does not implement correctly
 * (spec: src/common/core/CUDABodies.cu:172-211):  v+ = v + a(x)*dt/2 ; x' = x + v+*dt ; v' = v+ + a(x')*dt/2. */
enum { B200NB_INTEGRATOR_MURB = 0, B200NB_INTEGRATOR_LEAPFROG = 1 };

typedef struct b200nb_ctx b200nb_ctx;

/* ---- lifetime ------------------------------------------------------------------------------------------------
 * Replaces: CUDABodies ctor/dtor (src/common/core/CUDABodies.cu:3-49,379-399) + SimulationNBodyCUDATileFullDevice
 * ctor/dtor (src/murb/implem/SimulationNBodyCUDATileFullDevice.cu:155-201,238-244). */

/* One process driving n_gpus devices (0 = every visible device, 1 = the current device).  Targets are sharded
 * contiguously over the devices; for n_gpus > 1 positions are exchanged with ncclAllGather each step. */
int b200nb_create(b200nb_ctx **out, uint64_t n_bodies, int n_gpus, float G, float soft);

/* One process driving n_shards shards with an explicit placement: shard i (targets [i*L, (i+1)*L)) lives on CUDA device
 * devices[i].  A device may be listed more than once ("virtual shards"): the slice arithmetic, chunk rotation, double
 * buffering and exchange of the multi-GPU path then run on a single GPU, which is how the sharded path is tested on a
 * one-GPU box.  Shards that share a device exchange positions with the p2p-push path (see b200nb_exchange_name; NCCL
 * cannot hold the same GPU twice in a communicator).  n_shards <= 16.
 * Replaces: the rank layout of SimulationNBodyMultiNode.cpp:62-91 (MPI ranks -> shards). */
int b200nb_create_sharded(b200nb_ctx **out, uint64_t n_bodies, int n_shards, const int *devices, float G, float soft);

/* One process per GPU (torchrun style): this process owns shard `rank` of `n_ranks` on CUDA device `device`.
 * nccl_id is the 128-byte ncclUniqueId made by b200nb_comm_unique_id() on rank 0 and broadcast by the caller
 * (ignored, may be NULL, when n_ranks == 1); like any ncclUniqueId it is good for ONE context — make a fresh one
 * for every b200nb_create_rank.  Collective over all ranks when n_ranks > 1.
 * Replaces: MPI_Init + buildCountsDispls in SimulationNBodyMultiNode.cpp:62-91. */
int b200nb_create_rank(b200nb_ctx **out, uint64_t n_bodies, float G, float soft, int rank, int n_ranks, int device,
                       const void *nccl_id);
int b200nb_comm_unique_id(void *id128);

void b200nb_destroy(b200nb_ctx *ctx);
const char *b200nb_last_error(const b200nb_ctx *ctx); /* ctx may be NULL: error of the last failed create */

/* ---- state ---------------------------------------------------------------------------------------------------
 * Host SoA arrays of n_bodies floats, the layout of dataSoA_t (src/common/core/Bodies.hpp:15-24).
 * upload replaces CUDABodies::memcpyBuffersOnDevice (CUDABodies.cu:31-49) and devInitializeDevGM
 * (SimulationNBodyCUDATileFullDevice.cu:41-45): G*m is folded into the device layout once.  Every rank passes the
 * full arrays (the reference's MPI variant replicates state the same way, SimulationNBodyMultiNode.cpp:93-117), but
 * each shard copies only its own slice (and reads the last body, where the padding sits) across the host link; the
 * position exchange of the step path then replicates the slices on every GPU.  Collective when n_ranks > 1. */
int b200nb_upload(b200nb_ctx *ctx, const float *qx, const float *qy, const float *qz, const float *m, const float *vx,
                  const float *vy, const float *vz);

/* Joins all devices, then copies positions and velocities of all n_bodies back (any pointer may be NULL).
 * Replaces the lazy D2H in CUDABodies::getDataSoA (CUDABodies.cu:63-93).  Collective when n_ranks > 1. */
int b200nb_download_state(b200nb_ctx *ctx, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz);

/* The same for the local shards' own bodies only: entries [first, first + count) of each array are written (global
 * indexing, see b200nb_slice_bounds), the rest is left untouched.  With one rank per process this is the D2H side of
 * a sharded driver (every rank keeps the host copy of its own targets, the analogue of the send buffers of
 * SimulationNBodyMultiNode.cpp:119-148) and moves 24 B per local body instead of 24 B per body per rank.  Not
 * collective. */
int b200nb_download_slice(b200nb_ctx *ctx, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz);

/* Accelerations of the last force pass (b200nb_accel or the last step).  Replaces getAccSoA()
 * (SimulationNBodyCUDAPropertyTracking.cu:308-319).  Collective when n_ranks > 1. */
int b200nb_download_accel(b200nb_ctx *ctx, float *ax, float *ay, float *az);

/* ---- the hot path ---------------------------------------------------------------------------------------------
 * b200nb_step == n_steps x computeOneIteration() (SimulationNBodyCUDATileFullDevice.cu:203-236; oracle:
 * SimulationNBodyNaive.cpp:56-61): force pass over all N^2 ordered pairs, self included, then the integrator.
 * Asynchronous: returns after enqueueing; b200nb_sync / download_* join. */
int b200nb_step(b200nb_ctx *ctx, float dt, int integrator, int n_steps);

/* Force pass only: computeBodiesAcceleration() (SimulationNBodyNaive.cpp:34-53) on the current positions. */
int b200nb_accel(b200nb_ctx *ctx);

/* Integrator only, with caller-supplied accelerations (host SoA, n_bodies each): the shape of
 * Bodies<T>::updatePositionsAndVelocities(const accSoA_t&, T&) (Bodies.cpp:280-288; CUDABodies.cu:353-371), which
 * test_CUDABodies.cpp:42-75 exercises.  MUrB integrator only. */
int b200nb_integrate_host_accel(b200nb_ctx *ctx, const float *ax, const float *ay, const float *az, float dt);

/* Total energy  sum_i [ m_i |v_i|^2 / 2  -  (1/2) sum_{j != i} G m_i m_j / sqrt(|r_ij|^2 + soft^2) ],  accumulated in
 * fp64.  Definition: devComputeBodiesMetrics (SimulationNBodyCUDAPropertyTracking.cu:217-304), which subtracts the
 * self term the same way.  Joins the devices.  Collective when n_ranks > 1. */
int b200nb_energy(b200nb_ctx *ctx, double *total);

/* Every per-iteration metric of the reference's history in one pass (SimulationHistory.hpp:12-15; CSV columns
 * "iteration,energy,ang_momentum,density_center_x,density_center_y,density_center_z", SimulationHistory.hpp:45 and
 * SimulationHistory.cpp:103-122).  Upstream only ever fills in the energy (SimulationNBodyCUDAPropertyTracking.cu:
 * 217-304); the angular momentum and density centre are declared and exported but never computed, so their
 * definitions are fixed here:  L = sum_i m_i (r_i x v_i) about the origin (ang_momentum = |L|);  centre of mass;
 * density centre = sum_i w_i r_i / sum_i w_i with w_i = m_i * sum_{j != i} G m_j / sqrt(|r_ij|^2 + soft^2), the
 * potential-weighted centre.  out[B200NB_N_METRICS], fp64.  Joins the devices.  Collective when n_ranks > 1. */
enum {
    B200NB_METRIC_ENERGY = 0,
    B200NB_METRIC_ANG_X = 1, B200NB_METRIC_ANG_Y = 2, B200NB_METRIC_ANG_Z = 3,
    B200NB_METRIC_MASS = 4,
    B200NB_METRIC_COM_X = 5, B200NB_METRIC_COM_Y = 6, B200NB_METRIC_COM_Z = 7,
    B200NB_METRIC_DENSITY_X = 8, B200NB_METRIC_DENSITY_Y = 9, B200NB_METRIC_DENSITY_Z = 10,
    B200NB_N_METRICS = 11
};
int b200nb_metrics(b200nb_ctx *ctx, double *out);

/* Joins every device and communication stream.  main.cpp:353-371 only synchronises the current device, so the glue
 * calls this when more than one device is in use. */
int b200nb_sync(b200nb_ctx *ctx);

/* ---- introspection / measurement ------------------------------------------------------------------------------ */
/* Targets per rank L (host only): rank r owns global bodies [r*L, min((r+1)*L, n)).  Contiguous, balanced to the
 * 256-body granularity (two 128-body AoSoA blocks; the launch rounds up to whole target tiles on its own); the
 * analogue of buildCountsDispls (SimulationNBodyMultiNode.cpp:76-91). */
uint64_t b200nb_slice_length(uint64_t n_bodies, int n_ranks);
/* Global body range [first, first + count) owned by the local_shard-th shard of this context (0 <= local_shard <
 * b200nb_n_local_gpus); count is 0 for a shard that holds only padding. */
int b200nb_slice_bounds(const b200nb_ctx *ctx, int local_shard, uint64_t *first, uint64_t *count);
uint64_t b200nb_n_bodies(const b200nb_ctx *ctx);
int b200nb_n_local_gpus(const b200nb_ctx *ctx);
uint64_t b200nb_allocated_bytes(const b200nb_ctx *ctx); /* device bytes, all local GPUs */
uint64_t b200nb_launch_count(const b200nb_ctx *ctx);    /* kernels of this library launched so far */
const char *b200nb_kernel_name(const b200nb_ctx *ctx);  /* force-kernel variant in use */
/* How positions travel between GPUs each step (the role of the MPI_Allgatherv calls of
 * SimulationNBodyMultiNode.cpp:93-148): "none" (one GPU); "nccl-allgather" (default) - in-place ncclAllGather on a
 * communication stream, overlapped with the own-slice force launch; "p2p-push" - in-process multi-GPU with NVLink peer
 * access, chosen with B200NB_EXCHANGE=p2p or when libnccl cannot be loaded: the integrator kernel stores the new
 * positions of its slice into every GPU's double-buffered body array, one event per GPU is the only synchronisation.
 * The two produce bit-identical results. */
const char *b200nb_exchange_name(const b200nb_ctx *ctx);

/* CUDA-event stopwatch on the library's own compute stream(s) (torch.cuda.Event cannot see them).
 * slot in [0,8); elapsed is the max over the local devices. */
int b200nb_event_record(b200nb_ctx *ctx, int slot);
int b200nb_event_elapsed_ms(b200nb_ctx *ctx, int slot_start, int slot_stop, float *ms);

/* When enabled every force-kernel launch is bracketed by events; get returns the accumulated device time of the
 * force kernels and their launch count since the last enable (joins the devices). */
int b200nb_profile_enable(b200nb_ctx *ctx, int on);
int b200nb_profile_get(b200nb_ctx *ctx, double *force_ms_total, uint64_t *force_launches);

/* Evicts the L2 between timed steps: overwrites a 256 MiB scratch buffer (larger than the 126 MB L2) on the compute
 * stream(s) with cudaMemsetAsync.  Measurement hygiene only; the hot path never calls it. */
int b200nb_flush_l2(b200nb_ctx *ctx);

/* Page-locked host buffers for the end-to-end path (cudaHostAlloc / cudaFreeHost). */
int b200nb_host_alloc(void **ptr, uint64_t bytes);
int b200nb_host_free(void *ptr);

/* ---- initial conditions (host only, no GPU needed) ------------------------------------------------------------
 * Bit-identical restatement of Bodies<float>::initGalaxy / initRandomly (src/common/core/Bodies.cpp:158-257) for
 * drivers that do not link the reference (bench.py).  scheme: 0 = "galaxy", 1 = "random"; arrays of n floats;
 * uses glibc srand(seed)/rand() in the reference's call order.  Not thread-safe (libc rand state). */
int b200nb_init_bodies(int scheme, uint64_t n, unsigned seed, float *qx, float *qy, float *qz, float *vx, float *vy,
                       float *vz, float *m, float *r);

/* The reference's file-based scheme, Bodies<T>::initMilkyWayAndromeda (src/common/core/Bodies.cpp:82-153): a text file
 * with one body per non-empty line, "mass x y z vx vy vz" in galaxy units, rescaled per galaxy component.  Count the
 * bodies first, then load into arrays of that many floats.  Host only. */
int b200nb_tab_count(const char *path, uint64_t *n_bodies);
int b200nb_tab_load(const char *path, uint64_t n, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz,
                    float *m, float *r);

#ifdef __cplusplus
}
#endif
#endif /* B200NB_H_ */
