// libb200nb: context, sharding, streams/events, NCCL exchange and the C ABI declared in include/b200nb.h.
//
// One context drives one or more "shards" (one per GPU).  A shard owns a contiguous range of L targets and keeps
//   bodies  : the FULL system as AoSoA blocks (positions + G*m), 16 B/body, replicated on every GPU
//   vel/acc : SoA [3][L] for its own targets,  mass [L],  partial [S][3][L] chunk partial sums
// State replication + target sharding is the reference's only distributed strategy
// (src/murb/implem/SimulationNBodyMultiNode.cpp:76-148: MPI_Allgatherv of qx,qy,qz,m then of ax,ay,az); here the
// exchange is ONE in-place ncclAllGather of the blocked slice per step on a communication stream, overlapped with
// the force pass over the rank's own (already resident) source chunks, and no acceleration gather at all because
// every GPU integrates only its own targets.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200nb.h"
#include "integrate_sm100.cuh"
#include "plan.hpp"
// The default force-kernel variant once more, as a cubin whose hot loop has been re-ordered after ptxas by
// tools/sass_resched.py (same instructions, same registers, same arithmetic: bit-identical results, ~5 % faster; the
// order was found with the GPU as the cost function and is kept in csrc/resched/*.order.json).  Generated at build time;
// the array is empty when this compiler's loop is not the one the order was derived from.
#include "generated/force_resched_cubin.inc"

using namespace b200nb;

namespace {

// ------------------------------------------------------------------------------------------------ NCCL (lazy)
// Loaded with dlopen only when more than one rank is requested, so single-GPU use has no NCCL dependency.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load()
    {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
#define LOADSYM(field, sym)                                                                                            \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, sym));                                                     \
    if (!field) { err = std::string("missing NCCL symbol ") + sym; return false; }
        LOADSYM(GetUniqueId, "ncclGetUniqueId");
        LOADSYM(CommInitRank, "ncclCommInitRank");
        LOADSYM(CommDestroy, "ncclCommDestroy");
        LOADSYM(GroupStart, "ncclGroupStart");
        LOADSYM(GroupEnd, "ncclGroupEnd");
        LOADSYM(AllGather, "ncclAllGather");
        LOADSYM(AllReduce, "ncclAllReduce");
        LOADSYM(GetErrorString, "ncclGetErrorString");
#undef LOADSYM
        return true;
    }
};
NcclApi g_nccl;
thread_local std::string g_create_error;

// ------------------------------------------------------------------------------------------------ driver API (lazy)
// Only used to load and launch the re-ordered cubin; resolved with dlopen so the library links against cudart alone.
struct DriverApi {
    void *handle = nullptr;
    bool tried = false;
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **) = nullptr;
    bool load()
    {
        if (tried) return handle != nullptr;
        tried = true;
        void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return false;
#define LOADDRV(field, sym)                                                                                            \
    field = reinterpret_cast<decltype(field)>(dlsym(h, sym));                                                          \
    if (!field) { dlclose(h); return false; }
        LOADDRV(ModuleLoadData, "cuModuleLoadData");
        LOADDRV(ModuleUnload, "cuModuleUnload");
        LOADDRV(ModuleGetFunction, "cuModuleGetFunction");
        LOADDRV(FuncSetAttribute, "cuFuncSetAttribute");
        LOADDRV(LaunchKernel, "cuLaunchKernel");
#undef LOADDRV
        handle = h;
        return true;
    }
};
DriverApi g_drv;

// ------------------------------------------------------------------------------------------------ kernel table
typedef void (*ForceKernelFn)(const ForceArgs);
typedef void (*ForceKernelSkFn)(const ForceArgsSK);
struct KernelVariant {
    const char *name;
    ForceKernelFn fn;
    int threads, r, tjb, st;
    size_t smem;
    double int_per_clk_sm; // measured steady-state rate (profiles/), used to choose between variants for small N
    ForceKernelSkFn fn_sk; // stream-K decomposition of the same inner loop (nullptr: not built for this variant)
    size_t smem_sk;
};
#define VARIANT(NAME, THREADS, R, TJB, ST, MATH, WP, U, MINB, RATE)                                                  \
    KernelVariant { NAME, force_kernel<THREADS, R, TJB, ST, MATH, WP, U, MINB>, THREADS, R, TJB, ST,                 \
                    force_smem_bytes<THREADS, R, TJB, ST, WP>(), RATE, nullptr, 0 }
#define VARIANT_SK(NAME, THREADS, R, TJB, ST, U, MINB, RATE)                                                         \
    KernelVariant { NAME, force_kernel<THREADS, R, TJB, ST, 1, false, U, MINB>, THREADS, R, TJB, ST,                 \
                    force_smem_bytes<THREADS, R, TJB, ST, false>(), RATE, force_kernel_sk<THREADS, R, TJB, ST, U, MINB>, \
                    force_sk_smem_bytes<THREADS, R, TJB, ST>() }
const KernelVariant g_variants[] = {
    // default first; chosen from the B200 sweep in profiles/ (tools/kbench)
    // 8 targets per thread and only 2 warps per scheduler: the operand-reuse cache keeps hitting while one warp keeps
    // issuing, which is what gets the accumulate FFMA2 triples back to 2 cycles
    VARIANT_SK("pk_t128_r8_tj2_st3_cta_u1_mb2", 128, 8, 2, 3, 1, 2, 9.53),
    // small systems: 256-target tiles and 1-block stages give enough CTAs to fill 148 SMs below N ~ 30k
    VARIANT_SK("pk_t128_r2_tj1_st3_cta_u2_mb4", 128, 2, 1, 3, 2, 4, 9.2),
    // one-warp CTAs, 8 per SM: the R = 8 inner loop on 256-target tiles.  Same rate as the default at large N (9.53),
    // ahead of it when a rank has few tiles (sharded runs: 25 088 targets x 200k sources 9.40 vs 9.34 vs 9.19 for the
    // R = 2 variant, profiles/r02_kbench_cluster_smalltiles.txt)
    VARIANT("pk_t32_r8_tj2_st2_cta_u1_mb8", 32, 8, 2, 2, 1, false, 1, 8, 9.53),
    VARIANT("pk_t128_r8_tj4_st2_cta_u1_mb2", 128, 8, 4, 2, 1, false, 1, 2, 9.5),
    VARIANT("pk_t256_r8_tj2_st3_cta_u1_mb1", 256, 8, 2, 3, 1, false, 1, 1, 9.5),
    VARIANT("pk_t256_r2_tj2_st3_cta_u2_mb3", 256, 2, 2, 3, 1, false, 2, 3, 9.0),
    VARIANT("sc_t256_r4_tj2_st3_cta_u1_mb2", 256, 4, 2, 3, 0, false, 1, 2, 8.2),
};
constexpr int N_VARIANTS = sizeof(g_variants) / sizeof(g_variants[0]);
constexpr uint64_t SLICE_ALIGN = 256;  // slice length granularity: a multiple of BLK; target tiles round up on their own (Lp)
constexpr uint32_t MAX_ROWS = 256; // upper bound on partial rows (the 1 GiB cap in max_rows_for usually binds first for big N)
constexpr int N_TIMER_SLOTS = 8;

struct Shard {
    int rank = 0, device = 0;
    cudaStream_t s_compute = nullptr, s_comm = nullptr;
    cudaEvent_t ev_integrated = nullptr, ev_gathered = nullptr;
    cudaEvent_t ev_pushed = nullptr; // p2p exchange: this shard's integrator has stored its new positions on every GPU
    float *bodies_next = nullptr;    // p2p exchange: the other half of the double-buffered body array
    cudaEvent_t timer[N_TIMER_SLOTS] = {};
    float *bodies = nullptr, *vel = nullptr, *acc = nullptr, *mass = nullptr, *partial = nullptr;
    float *stage = nullptr;      // 7 host-layout SoA slices of L floats (H2D / D2H staging of the shard's own bodies)
    float *stage_full = nullptr; // one rank per process only: 3 x stage_stride floats for a full-system position download (lazy)
    double *energy_blocks = nullptr, *energy_out = nullptr;
    void *l2_scratch = nullptr;
    ncclComm_t comm = nullptr;
    CUmodule resched_mod = nullptr;   // the re-ordered cubin of the default variant on this device (nullptr: not in use)
    CUfunction resched_fn = nullptr;
    int n_sms = 0, occ = 0, occ_sk = 0;
    uint32_t n_local = 0; // real bodies in the slice
    // profiling mode: event pairs (start, stop) around force launches.  In-flight pairs are folded into prof_ms and
    // recycled once PROF_RING of them are pending, so a long profiled run neither grows nor creates events per launch.
    std::vector<cudaEvent_t> prof, prof_free;
    double prof_ms = 0.0;
    uint64_t prof_n = 0;
    uint64_t bytes = 0;
};

} // namespace

// A captured run of `steps` consecutive iterations (force + integrator launches) for one (integrator, dt).
struct StepGraph {
    int integrator = 0;
    float dt = 0.f;
    int steps = 0;
    uint64_t launches = 0; // kernels inside the graph
    cudaGraphExec_t exec = nullptr;
};
constexpr int GRAPH_STEPS = 16;     // iterations per graph launch
constexpr int GRAPH_MIN_STEPS = 32; // only worth capturing for longer runs
constexpr size_t GRAPH_CACHE_MAX = 4; // executable graphs kept per context (oldest evicted)

struct b200nb_ctx {
    std::vector<StepGraph> graphs;
    uint64_t n = 0;
    int n_ranks = 1;
    uint64_t L = 0, total_pad = 0, stage_stride = 0, total_pad_hint = 0;
    uint64_t Lp = 0; // L rounded up to whole target tiles of the chosen variant: grid size and stride of the partial rows
    uint32_t nblk_total = 0;
    float G = 0.f, soft = 0.f, soft2 = 0.f;
    const KernelVariant *kv = nullptr;
    uint32_t k_per_slice = 1, rows = 1; // S = k_per_slice * n_ranks
    std::string kname;                  // variant name (+ "+sk" in stream-K mode)
    bool stream_k = false;              // stream-K decomposition instead of the (tile x chunk) grid
    bool p2p = false;                   // exchange = peer stores from the integrator (in-process multi-GPU) instead of NCCL
    bool merge_launches = true;         // one force launch when the exchange has already landed at enqueue time (B200NB_SPLIT_LAUNCHES=1: never)
    uint32_t sk_grid = 0;               // CTAs per stream-K launch (resident slots)
    uint32_t sk_rows_own = 0, sk_rows_rem = 0;
    std::vector<Shard> shards;
    bool uploaded = false, acc_valid = false, profiling = false;
    uint64_t launches = 0;
    std::string err;
};

namespace {

struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) saved = -1; }
    ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

int fail(b200nb_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(c, call)                                                                                                    \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
            return fail(c, B200NB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);  \
    } while (0)
#define NC(c, call)                                                                                                    \
    do {                                                                                                               \
        ncclResult_t r_ = (call);                                                                                      \
        if (r_ != ncclSuccess)                                                                                         \
            return fail(c, B200NB_ENCCL, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

int alloc_shard(b200nb_ctx *c, Shard &s)
{
    CU(c, cudaSetDevice(s.device));
    cudaDeviceProp prop;
    CU(c, cudaGetDeviceProperties(&prop, s.device));
    if (prop.major < 10)
        return fail(c, B200NB_ECUDA, "device %d is sm_%d%d; libb200nb is built for sm_100a only", s.device, prop.major,
                    prop.minor);
    s.n_sms = prop.multiProcessorCount;
    CU(c, cudaStreamCreateWithFlags(&s.s_compute, cudaStreamNonBlocking));
    CU(c, cudaStreamCreateWithFlags(&s.s_comm, cudaStreamNonBlocking));
    CU(c, cudaEventCreateWithFlags(&s.ev_integrated, cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&s.ev_gathered, cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&s.ev_pushed, cudaEventDisableTiming));
    for (auto &t : s.timer) CU(c, cudaEventCreate(&t));
    CU(c, cudaFuncSetAttribute(c->kv->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->kv->smem));
    CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ, c->kv->fn, c->kv->threads, c->kv->smem));
    if (s.occ < 1) return fail(c, B200NB_ECUDA, "force kernel %s cannot be resident on device %d", c->kv->name, s.device);
    if (c->kv->fn_sk) {
        CU(c, cudaFuncSetAttribute(c->kv->fn_sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->kv->smem_sk));
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s.occ_sk, c->kv->fn_sk, c->kv->threads, c->kv->smem_sk));
    }
    const uint64_t first = (uint64_t)s.rank * c->L;
    s.n_local = first >= c->n ? 0u : (uint32_t)std::min<uint64_t>(c->L, c->n - first);
    // the post-ptxas re-ordered build of the default variant (B200NB_NO_RESCHED=1: launch what ptxas scheduled)
    if (c->kv == &g_variants[2] && force_resched_cubin_size > 0 && !getenv("B200NB_NO_RESCHED") && g_drv.load()) {
        CU(c, cudaFree(nullptr)); // make sure the primary context exists and is current
        CUmodule mod = nullptr;
        CUfunction fn = nullptr;
        if (g_drv.ModuleLoadData(&mod, force_resched_cubin) == CUDA_SUCCESS && g_drv.ModuleGetFunction(&fn, mod, force_resched_kernel) == CUDA_SUCCESS &&
            g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)c->kv->smem) == CUDA_SUCCESS) {
            s.resched_mod = mod;
            s.resched_fn = fn;
        } else {
            if (mod) g_drv.ModuleUnload(mod);
            fprintf(stderr, "libb200nb: the re-ordered force kernel could not be loaded on device %d; using the ptxas-scheduled one\n", s.device);
        }
    }
    return B200NB_OK;
}

int alloc_buffers(b200nb_ctx *c, Shard &s)
{
    CU(c, cudaSetDevice(s.device));
    const size_t L = c->L;
    auto dmalloc = [&](void **p, size_t bytes) -> cudaError_t { s.bytes += bytes; return cudaMalloc(p, bytes); };
    CU(c, dmalloc((void **)&s.bodies, c->total_pad * 16));
    if (c->p2p) CU(c, dmalloc((void **)&s.bodies_next, c->total_pad * 16));
    CU(c, dmalloc((void **)&s.vel, 3 * L * 4));
    CU(c, dmalloc((void **)&s.acc, 3 * L * 4));
    CU(c, dmalloc((void **)&s.mass, L * 4));
    CU(c, dmalloc((void **)&s.partial, (size_t)c->rows * 3 * c->Lp * 4));
    CU(c, dmalloc((void **)&s.stage, 7 * L * 4));
    const size_t eb = (L + ENERGY_THREADS - 1) / ENERGY_THREADS;
    CU(c, dmalloc((void **)&s.energy_blocks, eb * 8 * MR_COUNT));
    CU(c, dmalloc((void **)&s.energy_out, 8 * MR_COUNT));
    CU(c, cudaMemsetAsync(s.acc, 0, 3 * L * 4, s.s_compute));
    CU(c, cudaMemsetAsync(s.vel, 0, 3 * L * 4, s.s_compute));
    CU(c, cudaEventRecord(s.ev_gathered, s.s_comm)); // "positions are current" before the first step
    return B200NB_OK;
}

const KernelVariant *pick_variant()
{
    const char *e = getenv("B200NB_VARIANT");
    if (e && *e) {
        for (int i = 0; i < N_VARIANTS; ++i)
            if (!strcmp(e, g_variants[i].name)) return &g_variants[i];
        const long idx = strtol(e, nullptr, 10);
        if (idx >= 0 && idx < N_VARIANTS && e[0] >= '0' && e[0] <= '9') return &g_variants[idx];
        fprintf(stderr, "libb200nb: unknown B200NB_VARIANT '%s', choosing automatically\n", e);
    }
    return nullptr; // choose by the planner's time estimate (choose_variant)
}

uint32_t max_rows_for(uint64_t L, int n_ranks)
{
    // rows of partial sums cost 12*L bytes each: never more than MAX_ROWS, never more than 1 GiB in total
    return (uint32_t)std::max<uint64_t>((uint64_t)n_ranks, std::min<uint64_t>(MAX_ROWS, (1ull << 30) / (12ull * L)));
}

// Source blocks the force pass has to visit.  The slice length is rounded up to the launch granularity, so the tail of
// the (last) slice is padding with G*m = 0; on one GPU everything past the last real body's block is skipped outright
// (0.3 % of the pass at N = 200k).  With several ranks every slice keeps its full length so that slices stay congruent.
uint32_t source_blocks(uint64_t n, uint64_t L, int n_ranks)
{
    return n_ranks == 1 ? (uint32_t)((n + BLK - 1) / BLK) : (uint32_t)(L * (uint64_t)n_ranks / BLK);
}

ChunkPlan plan_for(const KernelVariant &kv, uint64_t n, uint64_t L, int n_sms, int occ, int n_ranks)
{
    const uint32_t ti = kv.threads * kv.r;
    return plan_chunks((uint32_t)((L + ti - 1) / ti), source_blocks(n, L, n_ranks) / (uint32_t)n_ranks, (uint32_t)(n_sms * occ),
                       (uint32_t)n_ranks, max_rows_for((L + ti - 1) / ti * ti, n_ranks), (uint32_t)(2 * kv.tjb));
}

// Between the one-warp R = 8 variant (256-target tiles), the 4-warp R = 8 variant (1024-target tiles) and the small R = 2
// variant, take the one the planner expects to finish first (plan.hpp: choose_variant_index).  The one-warp variant is
// listed first: it matches the 4-warp one at large N (N = 200k inside bench.py: 14.52 vs 14.56 ms per launch, with 56
// instead of 74 partial rows; profiles/r02_ncu_force_kernel_*.txt) and its finer tiles win when a rank has few of them.
int choose_variant(b200nb_ctx *c, int device, const KernelVariant **out)
{
    cudaDeviceProp prop;
    CU(c, cudaSetDevice(device)); // the occupancy query below answers for the current device
    CU(c, cudaGetDeviceProperties(&prop, device));
    const KernelVariant *cands[3] = {&g_variants[2], &g_variants[0], &g_variants[1]};
    *out = cands[0];
    const char *mode = getenv("B200NB_MODE");
    if (mode && !strcmp(mode, "sk")) {
        // stream-K: every CTA gets ceil(U/G) (tile x block) units; a unit is TI x 128 interactions at SM rate / occ
        double best_t = 1e300;
        for (const KernelVariant *kv : {&g_variants[0], &g_variants[1]}) {
            int occ_sk = 0;
            CU(c, cudaFuncSetAttribute(kv->fn_sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kv->smem_sk));
            CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sk, kv->fn_sk, kv->threads, kv->smem_sk));
            if (occ_sk < 1) continue;
            const uint64_t G = (uint64_t)prop.multiProcessorCount * occ_sk;
            const uint64_t ti = (uint64_t)kv->threads * kv->r;
            const uint64_t U = ((c->L + ti - 1) / ti) * (c->total_pad_hint / BLK);
            const double t = (double)((U + G - 1) / G) * (double)ti * (double)occ_sk / kv->int_per_clk_sm;
            if (t < best_t) { best_t = t; *out = kv; }
        }
        return B200NB_OK;
    }
    VariantShape shapes[3];
    uint32_t max_rows[3];
    for (int i = 0; i < 3; ++i) {
        const KernelVariant *kv = cands[i];
        int occ = 0;
        CU(c, cudaFuncSetAttribute(kv->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kv->smem));
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kv->fn, kv->threads, kv->smem));
        shapes[i] = VariantShape{(uint32_t)kv->threads, (uint32_t)kv->r, (uint32_t)kv->tjb, (uint32_t)std::max(occ, 0), kv->int_per_clk_sm};
        const uint64_t ti = (uint64_t)kv->threads * kv->r;
        max_rows[i] = max_rows_for((c->L + ti - 1) / ti * ti, c->n_ranks);
    }
    const int pick = choose_variant_index(shapes, 3, c->L, source_blocks(c->n, c->L, c->n_ranks) / (uint32_t)c->n_ranks,
                                          (uint32_t)prop.multiProcessorCount, (uint32_t)c->n_ranks, max_rows);
    if (pick >= 0) *out = cands[pick];
    return B200NB_OK;
}

int enqueue_gather(b200nb_ctx *c);
int sync_all(b200nb_ctx *c);

constexpr size_t PROF_RING = 512; // pending (start, stop) pairs per shard before they are folded

// fold every pending pair into the running total (joins the stop events) and recycle the events
int prof_drain(b200nb_ctx *c, Shard &s)
{
    for (size_t i = 0; i + 1 < s.prof.size(); i += 2) {
        float ms = 0.f;
        CU(c, cudaEventSynchronize(s.prof[i + 1]));
        CU(c, cudaEventElapsedTime(&ms, s.prof[i], s.prof[i + 1]));
        s.prof_ms += ms;
        s.prof_n++;
    }
    s.prof_free.insert(s.prof_free.end(), s.prof.begin(), s.prof.end());
    s.prof.clear();
    return B200NB_OK;
}

// start event of a profiled launch (recorded on the compute stream); the matching prof_stop records the other one
int prof_start(b200nb_ctx *c, Shard &s)
{
    if (s.prof.size() >= 2 * PROF_RING)
        if (int rc = prof_drain(c, s)) return rc;
    while (s.prof_free.size() < 2) {
        cudaEvent_t ev;
        CU(c, cudaEventCreate(&ev));
        s.prof_free.push_back(ev);
    }
    cudaEvent_t e[2];
    for (auto &ev : e) { ev = s.prof_free.back(); s.prof_free.pop_back(); }
    s.prof.push_back(e[0]);
    s.prof.push_back(e[1]);
    CU(c, cudaEventRecord(e[0], s.s_compute));
    return B200NB_OK;
}
int prof_stop(b200nb_ctx *c, Shard &s)
{
    CU(c, cudaEventRecord(s.prof.back(), s.s_compute));
    return B200NB_OK;
}

// In-process multi-GPU has two exchange steps.  Default: in-place ncclAllGather on a communication stream, which runs
// entirely beside the own-slice force launch.  Alternative (B200NB_EXCHANGE=p2p, and automatically when libnccl cannot
// be loaded): every GPU can address every other one (NVLink / NVSwitch peers), so the integrator stores the new
// positions of its slice straight into the body array of every GPU (integrate_sm100.cuh: IntegrateArgs::out) and one
// event per GPU is the only synchronisation; the body array is double-buffered so that a fast GPU never overwrites
// positions a slower one is still reading.  Measured on 2 B200s the fused form is 0.1-1.4 % *slower*
// (profiles/r01_exchange_p2p_vs_nccl.txt): the peer stores must be flushed before the integrator kernel can retire,
// which puts an NVLink round trip on the compute stream's critical path, while the separate all-gather hides all of
// its latency behind the own-slice launch.  Both paths produce bit-identical results (tested).
int enable_p2p(b200nb_ctx *c)
{
    const char *e = getenv("B200NB_EXCHANGE");
    // several shards on one device ("virtual shards", b200nb_create_sharded): NCCL refuses a communicator with the
    // same GPU twice, and the peer path needs no peer at all there
    bool shared_device = false;
    for (size_t i = 0; i < c->shards.size(); ++i)
        for (size_t j = i + 1; j < c->shards.size(); ++j) shared_device |= c->shards[i].device == c->shards[j].device;
    if (shared_device) {
        if (e && !strcmp(e, "nccl"))
            return fail(c, B200NB_EINVAL, "B200NB_EXCHANGE=nccl cannot be used when several shards share one device");
        e = "p2p";
    }
    if (e && !strcmp(e, "nccl")) return B200NB_OK;
    if (e && *e && strcmp(e, "p2p"))
        return fail(c, B200NB_EINVAL, "B200NB_EXCHANGE must be 'p2p' or 'nccl', not '%s'", e);
    if (!(e && *e)) {
        if (g_nccl.load()) return B200NB_OK; // default
        e = nullptr;                         // no NCCL on this machine: try the peer path before giving up
    }
    if (c->shards.size() > (size_t)MAX_PUSH_TARGETS) {
        if (e) return fail(c, B200NB_EINVAL, "B200NB_EXCHANGE=p2p supports at most %d GPUs", MAX_PUSH_TARGETS);
        return B200NB_OK;
    }
    for (auto &a : c->shards)
        for (auto &b : c->shards) {
            if (a.device == b.device) continue;
            int ok = 0;
            CU(c, cudaDeviceCanAccessPeer(&ok, a.device, b.device));
            if (!ok) {
                if (e) return fail(c, B200NB_ECUDA, "B200NB_EXCHANGE=p2p: device %d cannot access device %d", a.device, b.device);
                return B200NB_OK; // no peer path: NCCL
            }
        }
    for (auto &a : c->shards) {
        CU(c, cudaSetDevice(a.device));
        for (auto &b : c->shards) {
            if (a.device == b.device) continue;
            const cudaError_t r = cudaDeviceEnablePeerAccess(b.device, 0);
            if (r == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); continue; }
            CU(c, r);
        }
    }
    c->p2p = true;
    return B200NB_OK;
}

int create_common(b200nb_ctx **out, uint64_t n, float G, float soft, int n_ranks, const std::vector<int> &ranks,
                  const std::vector<int> &devices, const void *nccl_id)
{
    if (!out) return fail(nullptr, B200NB_EINVAL, "out is NULL");
    *out = nullptr;
    if (n == 0 || n > (1ull << 31)) return fail(nullptr, B200NB_EINVAL, "n_bodies must be in [1, 2^31]");
    if (!(soft != 0.f)) return fail(nullptr, B200NB_EINVAL, "softening factor can't be equal to 0 (main.cpp:150-155)");
    DeviceGuard guard;
    b200nb_ctx *c = new b200nb_ctx();
    c->n = n; c->n_ranks = n_ranks; c->G = G; c->soft = soft; c->soft2 = soft * soft;
    c->kv = pick_variant();
    c->L = b200nb_slice_length(n, n_ranks);
    c->total_pad_hint = c->L * n_ranks;
    if (!c->kv) {
        if (int rc = choose_variant(c, devices[0], &c->kv)) { g_create_error = c->err; delete c; return rc; }
    }
    c->total_pad = c->L * n_ranks;
    c->nblk_total = (uint32_t)(c->total_pad / BLK);
    c->stage_stride = (n + 3) / 4 * 4;
    c->shards.resize(ranks.size());
    auto bail = [&](int code) {
        g_create_error = c->err;
        b200nb_destroy(c);
        return code;
    };
    for (size_t i = 0; i < ranks.size(); ++i) {
        c->shards[i].rank = ranks[i];
        c->shards[i].device = devices[i];
        if (int rc = alloc_shard(c, c->shards[i])) return bail(rc);
    }
    const Shard &s0 = c->shards[0];
    {
        const uint64_t ti = (uint64_t)c->kv->threads * c->kv->r;
        c->Lp = (c->L + ti - 1) / ti * ti;
    }
    c->k_per_slice = plan_for(*c->kv, c->n, c->L, s0.n_sms, s0.occ, n_ranks).n_chunks;
    c->rows = c->k_per_slice * n_ranks;
    {
        // "grid" (default): dynamic (tile x chunk) grid; "sk": static stream-K split, measured 3-8 % slower on B200
        // (profiles/r01_streamk_vs_grid.txt) and kept as a tested alternative
        const char *mode = getenv("B200NB_MODE");
        const bool want_sk = mode && !strcmp(mode, "sk");
        c->stream_k = want_sk && c->kv->fn_sk && s0.occ_sk >= 1;
        if (want_sk && !c->stream_k) { c->err = "B200NB_MODE=sk: variant has no stream-K kernel"; return bail(B200NB_EINVAL); }
    }
    c->kname = std::string(c->kv->name) + (c->stream_k ? "+sk" : "") + (!c->stream_k && c->shards[0].resched_fn ? "+resched" : "");
    c->merge_launches = getenv("B200NB_SPLIT_LAUNCHES") == nullptr;
    if (c->stream_k) {
        const uint32_t ti = c->kv->threads * c->kv->r;
        const uint32_t n_itiles = (uint32_t)(c->Lp / ti), nbs = (uint32_t)(c->L / BLK);
        const char *w = getenv("B200NB_SK_WAVES"); // experiment: CTAs = waves x resident slots
        c->sk_grid = (uint32_t)(s0.n_sms * s0.occ_sk) * (uint32_t)std::max(1, w ? atoi(w) : 1);
        auto max_rows = [&](uint32_t nb) {
            uint32_t m = 0;
            if (nb == 0) return m;
            const uint64_t U = (uint64_t)n_itiles * nb;
            for (uint32_t t = 0; t < n_itiles; ++t) m = std::max(m, sk_rows_of_tile(t, nb, U, c->sk_grid));
            return m;
        };
        c->sk_rows_own = max_rows(nbs);
        c->sk_rows_rem = max_rows(c->nblk_total - nbs);
        c->rows = c->sk_rows_own + c->sk_rows_rem;
    }
    if (n_ranks > 1 && ranks.size() == (size_t)n_ranks) {
        if (int rc = enable_p2p(c)) return bail(rc);
    }
    for (auto &s : c->shards)
        if (int rc = alloc_buffers(c, s)) return bail(rc);

    if (n_ranks > 1 && !c->p2p) {
        if (!g_nccl.load()) { c->err = g_nccl.err; return bail(B200NB_ENCCL); }
        ncclUniqueId id;
        if (nccl_id) memcpy(&id, nccl_id, sizeof id);
        else if (ranks.size() == (size_t)n_ranks) {
            ncclResult_t r = g_nccl.GetUniqueId(&id);
            if (r != ncclSuccess) { c->err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return bail(B200NB_ENCCL); }
        } else { c->err = "nccl_id is required when this process does not own every rank"; return bail(B200NB_EINVAL); }
        ncclResult_t r = g_nccl.GroupStart();
        for (auto &s : c->shards) {
            if (r != ncclSuccess) break;
            if (cudaSetDevice(s.device) != cudaSuccess) { r = ncclUnhandledCudaError; break; }
            r = g_nccl.CommInitRank(&s.comm, n_ranks, id, s.rank);
        }
        ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) { c->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); return bail(B200NB_ENCCL); }
        // NCCL builds its channels lazily on the first collective (hundreds of ms): pay that here, not in step 1
        // of the caller's timed loop (main.cpp:348-371 times every iteration).  The buffer content is irrelevant.
        if (int rc = enqueue_gather(c)) return bail(rc);
        if (int rc = sync_all(c)) return bail(rc);
    }
    *out = c;
    return B200NB_OK;
}

// Have the remote slices of the current positions already landed on shard `s` (host-side query, no waiting)?  True when
// the caller steps synchronously (the CLI joins the device after every iteration, main.cpp:356-368): the exchange of
// the previous step is long over by the time the next force pass is enqueued, there is nothing left to overlap, and
// the pass can be ONE launch over all chunks instead of own-slice + remote launches with a tail each.
bool remote_positions_landed(b200nb_ctx *c, Shard &s)
{
    bool ready = true;
    if (!c->p2p) {
        ready = cudaEventQuery(s.ev_gathered) == cudaSuccess;
    } else {
        for (auto &o : c->shards) {
            if (&o == &s || !ready) continue;
            if (o.device != s.device) (void)cudaSetDevice(o.device); // query an event on the device that owns it
            ready = cudaEventQuery(o.ev_pushed) == cudaSuccess;
        }
        (void)cudaSetDevice(s.device);
    }
    if (!ready) (void)cudaGetLastError(); // cudaErrorNotReady is an answer, not a fault
    return ready;
}

// the compute stream of `s` may not read remote slices before they have landed
int wait_remote_positions(b200nb_ctx *c, Shard &s)
{
    if (!c->p2p) {
        CU(c, cudaStreamWaitEvent(s.s_compute, s.ev_gathered, 0));
        return B200NB_OK;
    }
    for (auto &o : c->shards)
        if (&o != &s) CU(c, cudaStreamWaitEvent(s.s_compute, o.ev_pushed, 0));
    return B200NB_OK;
}

// stream-K launches: one for the rank's own (already resident) source blocks, one for the gathered remote blocks
int enqueue_force_sk(b200nb_ctx *c)
{
    const KernelVariant &kv = *c->kv;
    const uint32_t ti = kv.threads * kv.r;
    const uint32_t nbs = (uint32_t)(c->L / BLK);
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        ForceArgsSK a{};
        a.src = s.bodies; a.tgt = s.bodies; a.partial = s.partial;
        a.tgt_blk0 = (uint32_t)((uint64_t)s.rank * c->L / BLK);
        a.tgt_stride = (uint32_t)c->Lp; a.tgt_count = (uint32_t)c->L;
        a.src_nblk_total = c->nblk_total;
        a.blk_rot = nbs * (uint32_t)s.rank;
        a.n_itiles = (uint32_t)(c->Lp / ti);
        a.soft2 = c->soft2;
        auto launch = [&](uint32_t lb0, uint32_t nb, uint32_t row0) -> int {
            a.lb0 = lb0; a.nb = nb; a.row0 = row0;
            if (c->profiling) if (int rc = prof_start(c, s)) return rc;
            kv.fn_sk<<<c->sk_grid, kv.threads, kv.smem_sk, s.s_compute>>>(a);
            if (c->profiling) if (int rc = prof_stop(c, s)) return rc;
            c->launches++;
            CU(c, cudaGetLastError());
            return B200NB_OK;
        };
        if (int rc = launch(0, nbs, 0)) return rc; // own slice: resident, overlaps the all-gather
        if (c->n_ranks > 1) {
            if (int rc = wait_remote_positions(c, s)) return rc;
            if (int rc = launch(nbs, c->nblk_total - nbs, c->sk_rows_own)) return rc;
        }
    }
    return B200NB_OK;
}

// enqueue the force pass (own chunks, wait for the gather, remote chunks) on every shard's compute stream
int enqueue_force(b200nb_ctx *c)
{
    if (c->stream_k) return enqueue_force_sk(c);
    const KernelVariant &kv = *c->kv;
    const uint32_t ti = kv.threads * kv.r;
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        ForceArgs a{};
        a.src = s.bodies; a.tgt = s.bodies; a.partial = s.partial;
        a.tgt_blk0 = (uint32_t)((uint64_t)s.rank * c->L / BLK);
        a.tgt_stride = (uint32_t)c->Lp; a.tgt_count = (uint32_t)c->L;
        a.src_nblk_total = source_blocks(c->n, c->L, c->n_ranks);
        a.n_chunks_total = c->rows;
        a.chunk_rot = c->k_per_slice * (uint32_t)s.rank;
        a.soft2 = c->soft2;
        a.dbg = nullptr;
        const uint32_t n_itiles = (uint32_t)(c->Lp / ti);
        auto launch = [&](uint32_t first, uint32_t count) -> int {
            a.chunk_first = first;
            if (c->profiling) if (int rc = prof_start(c, s)) return rc;
            if (s.resched_fn) {
                void *params[1] = {&a};
                const CUresult r = g_drv.LaunchKernel(s.resched_fn, n_itiles, count, 1, (unsigned)kv.threads, 1, 1, (unsigned)kv.smem,
                                                      (CUstream)s.s_compute, params, nullptr);
                if (r != CUDA_SUCCESS) return fail(c, B200NB_ECUDA, "cuLaunchKernel(re-ordered force kernel) failed: %d", (int)r);
            } else {
                kv.fn<<<dim3(n_itiles, count), kv.threads, kv.smem, s.s_compute>>>(a);
            }
            if (c->profiling) if (int rc = prof_stop(c, s)) return rc;
            c->launches++;
            CU(c, cudaGetLastError());
            return B200NB_OK;
        };
        if (c->n_ranks == 1) {
            if (int rc = launch(0, c->rows)) return rc;
        } else if (c->merge_launches && remote_positions_landed(c, s)) {
            if (int rc = wait_remote_positions(c, s)) return rc; // already complete: orders the stream, costs nothing
            if (int rc = launch(0, c->rows)) return rc;          // same chunks, same rows: bit-identical to the split form
        } else {
            if (int rc = launch(0, c->k_per_slice)) return rc; // own slice: resident, overlaps the exchange
            if (int rc = wait_remote_positions(c, s)) return rc;
            if (int rc = launch(c->k_per_slice, c->rows - c->k_per_slice)) return rc;
        }
    }
    return B200NB_OK;
}

int enqueue_integrate(b200nb_ctx *c, int mode, float dt)
{
    const bool moves = mode == IM_MURB || mode == IM_MURB_STORED || mode == IM_LF_KICK_DRIFT;
    const bool push = c->p2p && moves;
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        if (push) {
            // the buffers about to be overwritten were last read one position update ago: every GPU's kernels of
            // that time precede its previous push in stream order, so waiting for those pushes is enough
            if (int rc = wait_remote_positions(c, s)) return rc;
        }
        if (s.n_local == 0) continue;
        IntegrateArgs a{};
        a.bodies = s.bodies; a.vel = s.vel; a.acc = s.acc; a.partial = s.partial;
        if (push) {
            a.n_out = 0;
            for (auto &o : c->shards) a.out[a.n_out++] = o.bodies_next;
        } else {
            a.out[0] = s.bodies; a.n_out = 1;
        }
        a.rows = c->rows; a.L = (uint32_t)c->L; a.pstride = (uint32_t)c->Lp; a.n_local = s.n_local;
        a.first = (uint64_t)s.rank * c->L; a.dt = dt; a.mode = mode;
        if (c->stream_k) {
            const uint32_t ti = c->kv->threads * c->kv->r, n_itiles = (uint32_t)(c->Lp / ti), nbs = (uint32_t)(c->L / BLK);
            a.sk_ti = ti;
            a.sk[0] = SkRows{(uint64_t)n_itiles * nbs, c->sk_grid, nbs, 0};
            const uint32_t nbr = c->nblk_total - nbs;
            a.sk[1] = SkRows{(uint64_t)n_itiles * nbr, c->sk_grid, nbr, c->sk_rows_own};
        }
        integrate_kernel<<<(s.n_local + 255) / 256, 256, 0, s.s_compute>>>(a);
        c->launches++;
        CU(c, cudaGetLastError());
    }
    if (push) {
        for (auto &s : c->shards) {
            CU(c, cudaSetDevice(s.device));
            CU(c, cudaEventRecord(s.ev_pushed, s.s_compute));
            std::swap(s.bodies, s.bodies_next);
        }
    }
    return B200NB_OK;
}

// positions of the own slice changed: publish them to every other rank (in place, on the comm stream)
int enqueue_gather(b200nb_ctx *c)
{
    if (c->n_ranks == 1 || c->p2p) return B200NB_OK; // p2p: the integrator already stored the positions everywhere
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        CU(c, cudaEventRecord(s.ev_integrated, s.s_compute));
        CU(c, cudaStreamWaitEvent(s.s_comm, s.ev_integrated, 0));
    }
    NC(c, g_nccl.GroupStart());
    ncclResult_t r = ncclSuccess;
    for (auto &s : c->shards) {
        if (cudaSetDevice(s.device) != cudaSuccess) { r = ncclUnhandledCudaError; break; }
        const size_t count = (size_t)c->L * 4; // floats in one blocked slice
        r = g_nccl.AllGather(s.bodies + (size_t)s.rank * count, s.bodies, count, ncclFloat, s.comm, s.s_comm);
        if (r != ncclSuccess) break;
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return fail(c, B200NB_ENCCL, "ncclAllGather failed: %s", g_nccl.GetErrorString(r));
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        CU(c, cudaEventRecord(s.ev_gathered, s.s_comm));
    }
    return B200NB_OK;
}

int sync_all(b200nb_ctx *c)
{
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        CU(c, cudaStreamSynchronize(s.s_compute));
        CU(c, cudaStreamSynchronize(s.s_comm));
    }
    return B200NB_OK;
}

// gather a [3][L] per-rank device array (vel or acc) into host SoA arrays of n floats
int download_sliced(b200nb_ctx *c, float *Shard::*field, float *hx, float *hy, float *hz)
{
    float *dst[3] = {hx, hy, hz};
    const bool all_local = c->shards.size() == (size_t)c->n_ranks;
    if (all_local) {
        for (auto &s : c->shards) {
            if (s.n_local == 0) continue;
            CU(c, cudaSetDevice(s.device));
            for (int k = 0; k < 3; ++k)
                if (dst[k])
                    CU(c, cudaMemcpyAsync(dst[k] + (size_t)s.rank * c->L, (s.*field) + (size_t)k * c->L,
                                          (size_t)s.n_local * 4, cudaMemcpyDeviceToHost, s.s_compute));
        }
        return sync_all(c);
    }
    // one rank per process: all-gather the slices through a temporary [P][3][L] device buffer
    Shard &s = c->shards[0];
    CU(c, cudaSetDevice(s.device));
    struct Scratch { // freed on every exit path
        float *p = nullptr;
        ~Scratch() { if (p) cudaFree(p); }
    } tmp;
    const size_t count = 3 * (size_t)c->L;
    CU(c, cudaMalloc((void **)&tmp.p, count * c->n_ranks * 4));
    CU(c, cudaMemcpyAsync(tmp.p + (size_t)s.rank * count, s.*field, count * 4, cudaMemcpyDeviceToDevice, s.s_compute));
    NC(c, g_nccl.AllGather(tmp.p + (size_t)s.rank * count, tmp.p, count, ncclFloat, s.comm, s.s_compute));
    for (int rk = 0; rk < c->n_ranks; ++rk) {
        const uint64_t first = (uint64_t)rk * c->L;
        if (first >= c->n) break;
        const size_t cnt = (size_t)std::min<uint64_t>(c->L, c->n - first);
        for (int k = 0; k < 3; ++k)
            if (dst[k])
                CU(c, cudaMemcpyAsync(dst[k] + first, tmp.p + (size_t)rk * count + (size_t)k * c->L, cnt * 4,
                                      cudaMemcpyDeviceToHost, s.s_compute));
    }
    CU(c, cudaStreamSynchronize(s.s_compute));
    return B200NB_OK;
}

} // namespace

// =================================================================================================== C ABI
extern "C" {

int b200nb_create(b200nb_ctx **out, uint64_t n_bodies, int n_gpus, float G, float soft)
{
    int cur = 0, visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1)
        return fail(nullptr, B200NB_ECUDA, "no CUDA device visible (libb200nb has no CPU fallback)");
    if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
    if (n_gpus < 0 || n_gpus > visible) return fail(nullptr, B200NB_EINVAL, "n_gpus=%d but %d device(s) visible", n_gpus, visible);
    if (n_gpus == 0) n_gpus = visible;
    std::vector<int> ranks, devices;
    for (int i = 0; i < n_gpus; ++i) {
        ranks.push_back(i);
        devices.push_back(n_gpus == 1 ? cur : i);
    }
    return create_common(out, n_bodies, G, soft, n_gpus, ranks, devices, nullptr);
}

int b200nb_create_sharded(b200nb_ctx **out, uint64_t n_bodies, int n_shards, const int *devices, float G, float soft)
{
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1)
        return fail(nullptr, B200NB_ECUDA, "no CUDA device visible (libb200nb has no CPU fallback)");
    if (n_shards < 1 || n_shards > MAX_PUSH_TARGETS || !devices)
        return fail(nullptr, B200NB_EINVAL, "n_shards must be in [1, %d] with a device list", MAX_PUSH_TARGETS);
    std::vector<int> ranks, devs;
    for (int i = 0; i < n_shards; ++i) {
        if (devices[i] < 0 || devices[i] >= visible) return fail(nullptr, B200NB_EINVAL, "device %d not visible", devices[i]);
        ranks.push_back(i);
        devs.push_back(devices[i]);
    }
    return create_common(out, n_bodies, G, soft, n_shards, ranks, devs, nullptr);
}

int b200nb_create_rank(b200nb_ctx **out, uint64_t n_bodies, float G, float soft, int rank, int n_ranks, int device,
                       const void *nccl_id)
{
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(nullptr, B200NB_EINVAL, "bad rank %d of %d", rank, n_ranks);
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1)
        return fail(nullptr, B200NB_ECUDA, "no CUDA device visible (libb200nb has no CPU fallback)");
    if (device < 0 || device >= visible) return fail(nullptr, B200NB_EINVAL, "device %d not visible", device);
    if (n_ranks > 1 && !nccl_id) return fail(nullptr, B200NB_EINVAL, "nccl_id is required for n_ranks > 1");
    return create_common(out, n_bodies, G, soft, n_ranks, {rank}, {device}, nccl_id);
}

int b200nb_comm_unique_id(void *id128)
{
    if (!id128) return fail(nullptr, B200NB_EINVAL, "id128 is NULL");
    if (!g_nccl.load()) return fail(nullptr, B200NB_ENCCL, "%s", g_nccl.err.c_str());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, B200NB_ENCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    memcpy(id128, &id, sizeof id);
    return B200NB_OK;
}

void b200nb_destroy(b200nb_ctx *c)
{
    if (!c) return;
    DeviceGuard guard;
    for (auto &s : c->shards) { // join before anything the streams may still be using goes away
        if (cudaSetDevice(s.device) != cudaSuccess) continue;
        if (s.s_compute) cudaStreamSynchronize(s.s_compute);
        if (s.s_comm) cudaStreamSynchronize(s.s_comm);
    }
    for (auto &g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto &s : c->shards) {
        if (cudaSetDevice(s.device) != cudaSuccess) continue;
        if (s.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s.comm);
        if (s.resched_mod) g_drv.ModuleUnload(s.resched_mod);
        for (auto e : s.prof) cudaEventDestroy(e);
        for (auto e : s.prof_free) cudaEventDestroy(e);
        cudaFree(s.bodies); cudaFree(s.bodies_next); cudaFree(s.vel); cudaFree(s.acc); cudaFree(s.mass); cudaFree(s.partial); cudaFree(s.stage); cudaFree(s.stage_full);
        cudaFree(s.energy_blocks); cudaFree(s.energy_out); cudaFree(s.l2_scratch);
        for (auto t : s.timer) if (t) cudaEventDestroy(t);
        if (s.ev_integrated) cudaEventDestroy(s.ev_integrated);
        if (s.ev_gathered) cudaEventDestroy(s.ev_gathered);
        if (s.ev_pushed) cudaEventDestroy(s.ev_pushed);
        if (s.s_compute) cudaStreamDestroy(s.s_compute);
        if (s.s_comm) cudaStreamDestroy(s.s_comm);
    }
    delete c;
}

const char *b200nb_last_error(const b200nb_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

// After an upload every shard holds only its own slice of the blocked array: replicate it everywhere (the same exchange
// a step performs).  p2p: each shard copies its slice into both halves of every shard's double buffer (G*m and the
// padding bodies live in the buffers too); NCCL: the in-place all-gather.
static int publish_slices(b200nb_ctx *c)
{
    if (c->n_ranks == 1) return B200NB_OK;
    if (!c->p2p) return enqueue_gather(c);
    const size_t slice_floats = (size_t)c->L * 4;
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        const float *mine = s.bodies + (size_t)s.rank * slice_floats;
        for (auto &o : c->shards) {
            if (&o != &s)
                CU(c, cudaMemcpyAsync(o.bodies + (size_t)s.rank * slice_floats, mine, slice_floats * 4, cudaMemcpyDefault, s.s_compute));
            CU(c, cudaMemcpyAsync(o.bodies_next + (size_t)s.rank * slice_floats, mine, slice_floats * 4, cudaMemcpyDefault, s.s_compute));
        }
        CU(c, cudaEventRecord(s.ev_pushed, s.s_compute));
    }
    return B200NB_OK;
}

int b200nb_upload(b200nb_ctx *c, const float *qx, const float *qy, const float *qz, const float *m, const float *vx,
                  const float *vy, const float *vz)
{
    if (!c) return B200NB_EINVAL;
    if (!qx || !qy || !qz || !m || !vx || !vy || !vz) return fail(c, B200NB_EINVAL, "upload: NULL array");
    DeviceGuard guard;
    const float *src[7] = {qx, qy, qz, m, vx, vy, vz};
    if (c->p2p) { // peers may still be storing positions into this GPU's buffers
        if (int rc = sync_all(c)) return rc;
    }
    const float px = qx[c->n - 1], py = qy[c->n - 1], pz = qz[c->n - 1]; // where the padding bodies sit
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        // the previous step's gather may still be writing remote slices of `bodies`
        CU(c, cudaStreamWaitEvent(s.s_compute, s.ev_gathered, 0));
        // only the shard's own bodies cross the host link (7 x 4 B each); the exchange below replicates the positions
        const size_t first = (size_t)s.rank * c->L;
        if (s.n_local > 0)
            for (int k = 0; k < 7; ++k)
                CU(c, cudaMemcpyAsync(s.stage + (size_t)k * c->L, src[k] + first, (size_t)s.n_local * 4, cudaMemcpyHostToDevice, s.s_compute));
        const int grid = (int)std::min<uint64_t>((c->L + 255) / 256, (uint64_t)s.n_sms * 8);
        pack_slice_kernel<<<grid, 256, 0, s.s_compute>>>(s.stage, (uint32_t)c->L, s.n_local, first, px, py, pz, c->G, s.bodies,
                                                        s.vel, s.mass);
        c->launches++;
        CU(c, cudaGetLastError());
    }
    if (int rc = publish_slices(c)) return rc;
    // host pointers are only borrowed for the call: the copies must have left them (pinned memory is truly async)
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        CU(c, cudaStreamSynchronize(s.s_compute));
    }
    c->uploaded = true;
    c->acc_valid = false;
    return B200NB_OK;
}

// positions and velocities of the local shards' own bodies into the global-index host arrays (other entries untouched)
static int download_local(b200nb_ctx *c, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz)
{
    float *q[3] = {qx, qy, qz}, *v[3] = {vx, vy, vz};
    for (auto &s : c->shards) {
        if (s.n_local == 0) continue;
        CU(c, cudaSetDevice(s.device));
        const size_t first = (size_t)s.rank * c->L;
        if (qx || qy || qz) {
            const int grid = (int)std::min<uint64_t>((s.n_local + 255) / 256, (uint64_t)s.n_sms * 8);
            unpack_positions_kernel<<<grid, 256, 0, s.s_compute>>>(s.bodies, first, s.n_local, s.stage, c->L);
            c->launches++;
            CU(c, cudaGetLastError());
        }
        for (int k = 0; k < 3; ++k) {
            if (q[k])
                CU(c, cudaMemcpyAsync(q[k] + first, s.stage + (size_t)k * c->L, (size_t)s.n_local * 4, cudaMemcpyDeviceToHost, s.s_compute));
            if (v[k])
                CU(c, cudaMemcpyAsync(v[k] + first, s.vel + (size_t)k * c->L, (size_t)s.n_local * 4, cudaMemcpyDeviceToHost, s.s_compute));
        }
    }
    return sync_all(c);
}

int b200nb_download_state(b200nb_ctx *c, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz)
{
    if (!c) return B200NB_EINVAL;
    if (!c->uploaded) return fail(c, B200NB_ESTATE, "download_state before upload");
    DeviceGuard guard;
    if (int rc = sync_all(c)) return rc;
    if (c->shards.size() == (size_t)c->n_ranks) // every shard is local: each one returns its own slice over its own link
        return download_local(c, qx, qy, qz, vx, vy, vz);
    // one rank per process: positions are replicated, so this rank's device has all of them after the gather
    if (qx || qy || qz) {
        Shard &s = c->shards[0];
        CU(c, cudaSetDevice(s.device));
        if (!s.stage_full) {
            CU(c, cudaMalloc((void **)&s.stage_full, 3 * c->stage_stride * 4));
            s.bytes += 3 * c->stage_stride * 4;
        }
        const int grid = (int)std::min<uint64_t>((c->n + 255) / 256, (uint64_t)s.n_sms * 8);
        unpack_positions_kernel<<<grid, 256, 0, s.s_compute>>>(s.bodies, 0, c->n, s.stage_full, c->stage_stride);
        c->launches++;
        CU(c, cudaGetLastError());
        float *dst[3] = {qx, qy, qz};
        for (int k = 0; k < 3; ++k)
            if (dst[k])
                CU(c, cudaMemcpyAsync(dst[k], s.stage_full + (size_t)k * c->stage_stride, c->n * 4, cudaMemcpyDeviceToHost, s.s_compute));
        CU(c, cudaStreamSynchronize(s.s_compute));
    }
    if (vx || vy || vz) return download_sliced(c, &Shard::vel, vx, vy, vz);
    return B200NB_OK;
}

int b200nb_download_slice(b200nb_ctx *c, float *qx, float *qy, float *qz, float *vx, float *vy, float *vz)
{
    if (!c) return B200NB_EINVAL;
    if (!c->uploaded) return fail(c, B200NB_ESTATE, "download_slice before upload");
    DeviceGuard guard;
    if (int rc = sync_all(c)) return rc;
    return download_local(c, qx, qy, qz, vx, vy, vz);
}

int b200nb_slice_bounds(const b200nb_ctx *c, int local_shard, uint64_t *first, uint64_t *count)
{
    if (!c || local_shard < 0 || (size_t)local_shard >= c->shards.size()) return B200NB_EINVAL;
    const Shard &s = c->shards[local_shard];
    if (first) *first = std::min<uint64_t>((uint64_t)s.rank * c->L, c->n);
    if (count) *count = s.n_local;
    return B200NB_OK;
}

int b200nb_download_accel(b200nb_ctx *c, float *ax, float *ay, float *az)
{
    if (!c) return B200NB_EINVAL;
    if (!c->uploaded) return fail(c, B200NB_ESTATE, "download_accel before upload");
    DeviceGuard guard;
    if (int rc = sync_all(c)) return rc;
    return download_sliced(c, &Shard::acc, ax, ay, az);
}

// one iteration, enqueued on the compute stream(s)
static int enqueue_one_step(b200nb_ctx *c, float dt, int integrator)
{
    if (integrator == B200NB_INTEGRATOR_MURB) {
        if (int rc = enqueue_force(c)) return rc;
        if (int rc = enqueue_integrate(c, IM_MURB, dt)) return rc;
        if (int rc = enqueue_gather(c)) return rc;
        c->acc_valid = false; // acc belongs to the positions before the update
    } else {
        if (!c->acc_valid) { // a(x_0): once after an upload
            if (int rc = enqueue_force(c)) return rc;
            if (int rc = enqueue_integrate(c, IM_REDUCE_ONLY, dt)) return rc;
        }
        if (int rc = enqueue_integrate(c, IM_LF_KICK_DRIFT, dt)) return rc;
        if (int rc = enqueue_gather(c)) return rc;
        if (int rc = enqueue_force(c)) return rc;
        if (int rc = enqueue_integrate(c, IM_LF_KICK, dt)) return rc;
        c->acc_valid = true; // acc == a(x_{n+1})
    }
    return B200NB_OK;
}

// Long single-GPU runs replay a captured CUDA graph of GRAPH_STEPS iterations: at murb-test sizes (N ~ 2k) an
// iteration is a few microseconds of GPU work behind two launches, so launch latency is the whole cost.
static int get_step_graph(b200nb_ctx *c, float dt, int integrator, StepGraph **out)
{
    for (auto &g : c->graphs)
        if (g.integrator == integrator && g.dt == dt && g.steps == GRAPH_STEPS) { *out = &g; return B200NB_OK; }
    Shard &s = c->shards[0];
    CU(c, cudaSetDevice(s.device));
    const uint64_t before = c->launches;
    const bool acc_valid = c->acc_valid;
    CU(c, cudaStreamBeginCapture(s.s_compute, cudaStreamCaptureModeThreadLocal));
    int rc = B200NB_OK;
    for (int i = 0; i < GRAPH_STEPS && rc == B200NB_OK; ++i) rc = enqueue_one_step(c, dt, integrator);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s.s_compute, &graph);
    StepGraph g;
    g.integrator = integrator; g.dt = dt; g.steps = GRAPH_STEPS; g.launches = c->launches - before;
    c->launches = before;      // nothing ran yet
    c->acc_valid = acc_valid;
    if (rc != B200NB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(c, B200NB_ECUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) return fail(c, B200NB_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ei));
    // one executable graph per (integrator, dt): callers that change dt all the time would otherwise grow this
    // list without bound.  Eviction is rare: join the stream first so no launch of the evicted graph is in flight.
    while (c->graphs.size() >= GRAPH_CACHE_MAX) {
        CU(c, cudaStreamSynchronize(s.s_compute));
        cudaGraphExecDestroy(c->graphs.front().exec);
        c->graphs.erase(c->graphs.begin());
    }
    c->graphs.push_back(g);
    *out = &c->graphs.back();
    return B200NB_OK;
}

int b200nb_step(b200nb_ctx *c, float dt, int integrator, int n_steps)
{
    if (!c) return B200NB_EINVAL;
    if (!c->uploaded) return fail(c, B200NB_ESTATE, "step before upload");
    if (integrator != B200NB_INTEGRATOR_MURB && integrator != B200NB_INTEGRATOR_LEAPFROG)
        return fail(c, B200NB_EINVAL, "unknown integrator %d", integrator);
    if (n_steps < 0) return fail(c, B200NB_EINVAL, "n_steps < 0");
    DeviceGuard guard;
    int it = 0;
    static const bool no_graph = getenv("B200NB_NO_GRAPH") != nullptr;
    if (c->n_ranks == 1 && !c->profiling && !no_graph && n_steps >= GRAPH_MIN_STEPS) {
        if (integrator == B200NB_INTEGRATOR_LEAPFROG && !c->acc_valid) { // the one-off a(x_0) pass stays outside the graph
            if (int rc = enqueue_one_step(c, dt, integrator)) return rc;
            ++it;
        }
        StepGraph *g = nullptr;
        if (int rc = get_step_graph(c, dt, integrator, &g)) return rc;
        Shard &s = c->shards[0];
        CU(c, cudaSetDevice(s.device));
        for (; it + GRAPH_STEPS <= n_steps; it += GRAPH_STEPS) {
            CU(c, cudaGraphLaunch(g->exec, s.s_compute));
            c->launches += g->launches;
        }
        c->acc_valid = integrator == B200NB_INTEGRATOR_LEAPFROG;
    }
    for (; it < n_steps; ++it)
        if (int rc = enqueue_one_step(c, dt, integrator)) return rc;
    return B200NB_OK;
}

int b200nb_accel(b200nb_ctx *c)
{
    if (!c) return B200NB_EINVAL;
    if (!c->uploaded) return fail(c, B200NB_ESTATE, "accel before upload");
    DeviceGuard guard;
    if (int rc = enqueue_force(c)) return rc;
    if (int rc = enqueue_integrate(c, IM_REDUCE_ONLY, 0.f)) return rc;
    c->acc_valid = true;
    return B200NB_OK;
}

int b200nb_integrate_host_accel(b200nb_ctx *c, const float *ax, const float *ay, const float *az, float dt)
{
    if (!c) return B200NB_EINVAL;
    if (!c->uploaded) return fail(c, B200NB_ESTATE, "integrate before upload");
    if (!ax || !ay || !az) return fail(c, B200NB_EINVAL, "integrate_host_accel: NULL array");
    DeviceGuard guard;
    const float *src[3] = {ax, ay, az};
    for (auto &s : c->shards) {
        if (s.n_local == 0) continue;
        CU(c, cudaSetDevice(s.device));
        // an in-place all-gather of a preceding asynchronous step may still be sending this slice from the comm stream
        if (c->n_ranks > 1 && !c->p2p) CU(c, cudaStreamWaitEvent(s.s_compute, s.ev_gathered, 0));
        for (int k = 0; k < 3; ++k)
            CU(c, cudaMemcpyAsync(s.stage + (size_t)k * c->L, src[k] + (size_t)s.rank * c->L, (size_t)s.n_local * 4,
                                  cudaMemcpyHostToDevice, s.s_compute));
        load_acc_kernel<<<(s.n_local + 255) / 256, 256, 0, s.s_compute>>>(s.stage, c->L, s.n_local, s.acc, c->L);
        c->launches++;
        CU(c, cudaGetLastError());
    }
    if (int rc = enqueue_integrate(c, IM_MURB_STORED, dt)) return rc;
    if (int rc = enqueue_gather(c)) return rc;
    c->acc_valid = false;
    return sync_all(c); // borrowed host pointers (see upload)
}

// raw[MR_COUNT]: the metric rows summed over every body of every rank
static int metric_sums(b200nb_ctx *c, double *raw)
{
    if (!c->uploaded) return fail(c, B200NB_ESTATE, "metrics before upload");
    DeviceGuard guard;
    if (int rc = sync_all(c)) return rc; // positions gathered, velocities final
    for (int k = 0; k < MR_COUNT; ++k) raw[k] = 0.0;
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        const uint32_t nb = (s.n_local + ENERGY_THREADS - 1) / ENERGY_THREADS;
        double part[MR_COUNT];
        if (nb > 0) {
            energy_kernel<<<nb, ENERGY_THREADS, 0, s.s_compute>>>(s.bodies, s.vel, s.mass, (uint32_t)c->L, s.n_local,
                                                                 (uint64_t)s.rank * c->L, c->nblk_total, c->soft2,
                                                                 s.energy_blocks);
            energy_final_kernel<<<MR_COUNT, 256, 0, s.s_compute>>>(s.energy_blocks, nb, s.energy_out);
            c->launches += 2;
            CU(c, cudaGetLastError());
        } else {
            CU(c, cudaMemsetAsync(s.energy_out, 0, 8 * MR_COUNT, s.s_compute));
        }
        if (c->shards.size() != (size_t)c->n_ranks) // one rank per process: sum over ranks on the device
            NC(c, g_nccl.AllReduce(s.energy_out, s.energy_out, MR_COUNT, ncclDouble, ncclSum, s.comm, s.s_compute));
        CU(c, cudaMemcpyAsync(part, s.energy_out, 8 * MR_COUNT, cudaMemcpyDeviceToHost, s.s_compute));
        CU(c, cudaStreamSynchronize(s.s_compute));
        for (int k = 0; k < MR_COUNT; ++k) raw[k] += part[k];
    }
    return B200NB_OK;
}

int b200nb_energy(b200nb_ctx *c, double *total)
{
    if (!c || !total) return B200NB_EINVAL;
    double raw[MR_COUNT];
    if (int rc = metric_sums(c, raw)) return rc;
    *total = raw[MR_ENERGY];
    return B200NB_OK;
}

int b200nb_metrics(b200nb_ctx *c, double *out)
{
    if (!c || !out) return B200NB_EINVAL;
    double raw[MR_COUNT];
    if (int rc = metric_sums(c, raw)) return rc;
    out[B200NB_METRIC_ENERGY] = raw[MR_ENERGY];
    out[B200NB_METRIC_ANG_X] = raw[MR_LX];
    out[B200NB_METRIC_ANG_Y] = raw[MR_LY];
    out[B200NB_METRIC_ANG_Z] = raw[MR_LZ];
    out[B200NB_METRIC_MASS] = raw[MR_M];
    for (int k = 0; k < 3; ++k) {
        out[B200NB_METRIC_COM_X + k] = raw[MR_M] != 0.0 ? raw[MR_MX + k] / raw[MR_M] : 0.0;
        out[B200NB_METRIC_DENSITY_X + k] = raw[MR_W] != 0.0 ? raw[MR_WX + k] / raw[MR_W] : out[B200NB_METRIC_COM_X + k];
    }
    return B200NB_OK;
}

int b200nb_sync(b200nb_ctx *c)
{
    if (!c) return B200NB_EINVAL;
    DeviceGuard guard;
    return sync_all(c);
}

uint64_t b200nb_slice_length(uint64_t n_bodies, int n_ranks)
{
    if (n_ranks < 1) return 0;
    const uint64_t per = (n_bodies + n_ranks - 1) / n_ranks;
    return (per + SLICE_ALIGN - 1) / SLICE_ALIGN * SLICE_ALIGN;
}
uint64_t b200nb_n_bodies(const b200nb_ctx *c) { return c ? c->n : 0; }
int b200nb_n_local_gpus(const b200nb_ctx *c) { return c ? (int)c->shards.size() : 0; }
uint64_t b200nb_allocated_bytes(const b200nb_ctx *c)
{
    uint64_t b = 0;
    if (c) for (auto &s : c->shards) b += s.bytes;
    return b;
}
uint64_t b200nb_launch_count(const b200nb_ctx *c) { return c ? c->launches : 0; }
const char *b200nb_kernel_name(const b200nb_ctx *c) { return c ? c->kname.c_str() : ""; }
const char *b200nb_exchange_name(const b200nb_ctx *c)
{
    if (!c) return "";
    return c->n_ranks == 1 ? "none" : (c->p2p ? "p2p-push" : "nccl-allgather");
}

int b200nb_event_record(b200nb_ctx *c, int slot)
{
    if (!c || slot < 0 || slot >= N_TIMER_SLOTS) return B200NB_EINVAL;
    DeviceGuard guard;
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        CU(c, cudaEventRecord(s.timer[slot], s.s_compute));
    }
    return B200NB_OK;
}

int b200nb_event_elapsed_ms(b200nb_ctx *c, int a, int b, float *ms)
{
    if (!c || !ms || a < 0 || b < 0 || a >= N_TIMER_SLOTS || b >= N_TIMER_SLOTS) return B200NB_EINVAL;
    DeviceGuard guard;
    float mx = 0.f;
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        CU(c, cudaEventSynchronize(s.timer[b]));
        float t = 0.f;
        CU(c, cudaEventElapsedTime(&t, s.timer[a], s.timer[b]));
        mx = std::max(mx, t);
    }
    *ms = mx;
    return B200NB_OK;
}

int b200nb_profile_enable(b200nb_ctx *c, int on)
{
    if (!c) return B200NB_EINVAL;
    DeviceGuard guard;
    if (int rc = sync_all(c)) return rc;
    for (auto &s : c->shards) { // pending pairs are dropped, their events kept for reuse
        s.prof_free.insert(s.prof_free.end(), s.prof.begin(), s.prof.end());
        s.prof.clear();
        s.prof_ms = 0.0;
        s.prof_n = 0;
    }
    c->profiling = on != 0;
    return B200NB_OK;
}

int b200nb_profile_get(b200nb_ctx *c, double *force_ms_total, uint64_t *force_launches)
{
    if (!c) return B200NB_EINVAL;
    DeviceGuard guard;
    if (int rc = sync_all(c)) return rc;
    double mx = 0.0;
    uint64_t n = 0;
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        if (int rc = prof_drain(c, s)) return rc;
        mx = std::max(mx, s.prof_ms);
        n = std::max<uint64_t>(n, s.prof_n);
    }
    if (force_ms_total) *force_ms_total = mx;
    if (force_launches) *force_launches = n;
    return B200NB_OK;
}

int b200nb_flush_l2(b200nb_ctx *c)
{
    if (!c) return B200NB_EINVAL;
    DeviceGuard guard;
    constexpr size_t bytes = 256ull << 20;
    for (auto &s : c->shards) {
        CU(c, cudaSetDevice(s.device));
        if (!s.l2_scratch) CU(c, cudaMalloc(&s.l2_scratch, bytes));
        CU(c, cudaMemsetAsync(s.l2_scratch, 0x5a, bytes, s.s_compute));
    }
    return B200NB_OK;
}

int b200nb_host_alloc(void **ptr, uint64_t bytes)
{
    if (!ptr) return B200NB_EINVAL;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(nullptr, B200NB_ECUDA, "cudaHostAlloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
    return B200NB_OK;
}

int b200nb_host_free(void *ptr)
{
    if (!ptr) return B200NB_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? B200NB_OK : B200NB_ECUDA;
}

} // extern "C"
