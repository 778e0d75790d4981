// All-pairs softened-gravity force pass for sm_100a (B200).
//
// Replaces, from scratch, the reference's GPU force kernels
//   src/murb/implem/SimulationNBodyCUDATileFullDevice.cu:53-153   (gpu+tile+full)
//   src/murb/implem/SimulationNBodyCUDATileFullDevice200k.cu:102-175
// and computes the same law as the oracle path
//   src/murb/implem/SimulationNBodyNaive.cpp:34-53
//     a_i = sum_j G m_j (q_j - q_i) / (|q_j - q_i|^2 + soft^2)^(3/2),  self term included (contributes 0).
//
// Design (B200-first, not a port):
//  * Bodies live in HBM as AoSoA blocks of BLK=128 bodies: [x[128] | y[128] | z[128] | G*m[128]] = 2 KiB.
//    One block is one contiguous TMA bulk copy (cp.async.bulk -> SASS UBLKCP) and, once in shared memory,
//    is read as LDS.128 of four consecutive x / y / z / Gm values: exactly the operand shape of the
//    Blackwell-only packed FP32 instructions (FFMA2/FADD2/FMUL2, PTX *.f32x2).
//  * Two consecutive SOURCES (j, j+1) are processed per packed instruction against one target whose
//    coordinates ptxas folds into the broadcast `.F32` operand form: 14 issue slots per 2 interactions
//    (3 FADD2 + 6 FFMA2 + 3 FMUL2 + 2 MUFU.RSQ) instead of 26 scalar ones.
//  * R targets are register-blocked per thread so one LDS.128 feeds 4*R interactions.
//  * Source tiles are multi-buffered through shared memory by an mbarrier/TMA pipeline; the buffers are
//    either shared by the CTA (one __syncthreads per tile) or private per warp (no CTA barrier at all).
//  * The launch is a 2-D grid (target tiles x source chunks). Every CTA writes its partial sums to its
//    own row of `partial`; the integrator kernel adds the rows in a fixed order, so results are
//    deterministic and the grid can be sized to the 148 SMs independently of N.
//  * Accuracy: per-tile accumulators (<= TJB*64 terms per packed half) are folded into a second-level
//    accumulator after every tile, and chunk partials are summed in fp64 by the integrator, so the
//    fp32 summation error does not grow with N (north-star bound: max |da|/|a| <= 1e-5 vs fp64).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "plan.hpp"

namespace b200nb {

constexpr int BLK = 128;                 // bodies per AoSoA block
constexpr int BLK_FLOATS = 4 * BLK;      // x | y | z | gm
constexpr int BLK_BYTES = BLK_FLOATS * 4;

// float index of component `comp` (0=x,1=y,2=z,3=G*m) of body i inside a blocked array
__host__ __device__ __forceinline__ size_t blk_index(size_t i, int comp)
{
    return (i / BLK) * (size_t)BLK_FLOATS + (size_t)comp * BLK + (i % BLK);
}

// ------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint64_t pk2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok;
}
// Bounded wait: a pipeline bug traps (sticky launch error, reported by the next b200nb call as B200NB_ECUDA) instead of
// hanging the GPU.  The bound is wall time on %globaltimer, looked at every 4096 polls, not a poll count: a slow but
// legitimate completion (compute-sanitizer, cuda-gdb, time-slicing, preemption) must not kill the context.
#ifndef B200NB_MBAR_TIMEOUT_NS
#define B200NB_MBAR_TIMEOUT_NS 20000000000ull // 20 s; define as 0 to wait forever
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (B200NB_MBAR_TIMEOUT_NS != 0 && (++spins & 0xfffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > B200NB_MBAR_TIMEOUT_NS) __trap();
        }
    }
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// thread-block cluster helpers (distributed shared memory)
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// read a float from the shared memory of CTA `rank` of this cluster at the same offset as local address `addr`
__device__ __forceinline__ float ld_dsmem_f32(uint32_t addr, uint32_t rank)
{
    uint32_t remote;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------- launch args
struct ForceArgs {
    const float *src;            // blocked bodies the sources are read from (full replicated array)
    const float *tgt;            // blocked bodies the targets are read from (same array on one GPU)
    float *partial;              // [rows][3][tgt_stride] partial accelerations
    uint32_t tgt_blk0;           // first target block inside `tgt`
    uint32_t tgt_stride;         // floats per component row of `partial`: the local target count rounded up to whole tiles
    uint32_t tgt_count;          // local targets (the slice length); the last tile may reach past it: those threads
                                 // re-read the last target and their partial sums are never used
    uint32_t src_nblk_total;     // AoSoA blocks in the whole padded system (sources)
    uint32_t n_chunks_total;     // S: source chunks the system is cut into (a multiple of the rank count)
    uint32_t chunk_first;        // first LOGICAL chunk of this launch; gridDim.y chunks are processed
    uint32_t chunk_rot;          // logical -> physical chunk rotation, (logical + rot) % S.  rot = k*rank makes
                                 // logical chunks [0,k) the rank's own slice (resident before the all-gather)
    const uint32_t *chunk_tab;   // optional: {first block, end block} per LOGICAL chunk (rotation already applied);
                                 // overrides the balanced split above, so chunks may have different lengths
    float soft2;
    unsigned long long *dbg;     // optional: per-CTA {clock64 start,end, globaltimer start,end}
};

template <int THREADS, int R, int TJB, int ST, bool WARP_PRIVATE>
constexpr size_t force_smem_bytes()
{
    return (size_t)(WARP_PRIVATE ? THREADS / 32 : 1) * ST * (TJB * BLK_BYTES + 8);
}

// ------------------------------------------------------------------------------------------- inner math
// acc += f * d on a packed pair.  SCALAR_ACC issues it as two scalar FFMA: an FFMA2 whose three 64-bit operands are
// all distinct needs three register-file reads per lane pair, and the B200 sweep (profiles/) shows those
// accumulate instructions are where the FMA pipe loses cycles; two FFMA cost the same pipe time (2 x 1 clk) and the
// spare issue slots are free in the packed regime.
template <bool SCALAR_ACC>
__device__ __forceinline__ uint64_t acc_fma(uint64_t f, uint64_t d, uint64_t acc)
{
    if (!SCALAR_ACC) return fma2(f, d, acc);
    float f0, f1, d0, d1, a0, a1;
    upk2(f, f0, f1);
    upk2(d, d0, d1);
    upk2(acc, a0, a1);
    return pk2(fmaf(f0, d0, a0), fmaf(f1, d1, a1));
}

// Compile-time scheduling knobs.  They only permute independent statements (no numerical effect beyond the order
// of three commutative FMA terms) but change how ptxas interleaves the chains, i.e. how many accumulate triples keep
// the force factor in the operand-reuse cache; tools/tune_schedule.py scores every combination on the SASS with
// the measured register-read model and the winners are confirmed on the GPU.
#ifndef B200NB_KNOB_FMUL
#define B200NB_KNOB_FMUL 0  // 0: (inv*inv)*(g*inv), 1: ((inv*inv)*inv)*g
#endif
#ifndef B200NB_KNOB_LOOP
#define B200NB_KNOB_LOOP 0  // 0: source-pair outer / target inner, 1: target outer / source-pair inner
#endif
#ifndef B200NB_KNOB_DORD
#define B200NB_KNOB_DORD 3  // permutation of (x,y,z) in the r^2 chain
#endif
#ifndef B200NB_KNOB_AORD
#define B200NB_KNOB_AORD 1  // permutation of (x,y,z) in the accumulate triple
#endif

template <int P> struct Perm3;
template <> struct Perm3<0> { static constexpr int a = 0, b = 1, c = 2; };
template <> struct Perm3<1> { static constexpr int a = 0, b = 2, c = 1; };
template <> struct Perm3<2> { static constexpr int a = 1, b = 0, c = 2; };
template <> struct Perm3<3> { static constexpr int a = 1, b = 2, c = 0; };
template <> struct Perm3<4> { static constexpr int a = 2, b = 0, c = 1; };
template <> struct Perm3<5> { static constexpr int a = 2, b = 1, c = 0; };

// one target against one packed source pair: 3 FADD2 + 3 FFMA2 + 2 MUFU.RSQ + 3 FMUL2 + 3 FFMA2
template <bool SCALAR_ACC>
__device__ __forceinline__ void pair_interaction(const uint64_t (&sj)[3], uint64_t gj, float xi, float yi, float zi,
                                                 uint64_t soft2p, uint64_t (&acc)[3])
{
    uint64_t d3[3];
    d3[0] = sub2(sj[0], pk2(xi, xi));
    d3[1] = sub2(sj[1], pk2(yi, yi));
    d3[2] = sub2(sj[2], pk2(zi, zi));
    using D = Perm3<B200NB_KNOB_DORD>;
    uint64_t d = fma2(d3[D::a], d3[D::a], soft2p);
    d = fma2(d3[D::b], d3[D::b], d);
    d = fma2(d3[D::c], d3[D::c], d);
    float d0, d1;
    upk2(d, d0, d1);
    const uint64_t inv = pk2(rsqrt_approx(d0), rsqrt_approx(d1));
#if B200NB_KNOB_FMUL == 0
    const uint64_t gi = mul2(gj, inv);
    const uint64_t i2 = mul2(inv, inv);
    const uint64_t f = mul2(i2, gi);
#else
    const uint64_t i2 = mul2(inv, inv);
    const uint64_t i3 = mul2(i2, inv);
    const uint64_t f = mul2(i3, gj);
#endif
    using A = Perm3<B200NB_KNOB_AORD>;
    acc[A::a] = acc_fma<SCALAR_ACC>(f, d3[A::a], acc[A::a]);
    acc[A::b] = acc_fma<SCALAR_ACC>(f, d3[A::b], acc[A::b]);
    acc[A::c] = acc_fma<SCALAR_ACC>(f, d3[A::c], acc[A::c]);
}

// One AoSoA block (128 sources) against R register-blocked targets, packed f32x2 along the sources.
template <int R, int U, bool SCALAR_ACC>
__device__ __forceinline__ void block_packed(const float *__restrict__ sb, const float (&xi)[R], const float (&yi)[R],
                                             const float (&zi)[R], uint64_t soft2p, uint64_t (&ax)[R],
                                             uint64_t (&ay)[R], uint64_t (&az)[R])
{
#pragma unroll U
    for (int j = 0; j < BLK; j += 4) {
        const float4 xv = *reinterpret_cast<const float4 *>(sb + j);
        const float4 yv = *reinterpret_cast<const float4 *>(sb + BLK + j);
        const float4 zv = *reinterpret_cast<const float4 *>(sb + 2 * BLK + j);
        const float4 gv = *reinterpret_cast<const float4 *>(sb + 3 * BLK + j);
        const uint64_t s0[3] = {pk2(xv.x, xv.y), pk2(yv.x, yv.y), pk2(zv.x, zv.y)};
        const uint64_t s1[3] = {pk2(xv.z, xv.w), pk2(yv.z, yv.w), pk2(zv.z, zv.w)};
        const uint64_t g0 = pk2(gv.x, gv.y), g1 = pk2(gv.z, gv.w);
#if B200NB_KNOB_LOOP == 0
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
#else
#pragma unroll
        for (int k = 0; k < R; ++k) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#endif
                uint64_t acc[3] = {ax[k], ay[k], az[k]};
                pair_interaction<SCALAR_ACC>(h ? s1 : s0, h ? g1 : g0, xi[k], yi[k], zi[k], soft2p, acc);
                ax[k] = acc[0]; ay[k] = acc[1]; az[k] = acc[2];
            }
        }
    }
}

// The same block with the sources broadcast by warp shuffles instead of same-address shared-memory loads: every lane
// keeps one source of a 32-source group in registers (one conflict-free LDS per component) and each source pair costs
// 8 SHFL.IDX where block_packed needs 2 LDS.128.  Comparator only (MATH = 3): the arithmetic is identical, the
// shuffles add issue slots and 4 live registers and buy nothing, see profiles/r01_shuffle_vs_lds.txt.
template <int R>
__device__ __forceinline__ void block_packed_shfl(const float *__restrict__ sb, const float (&xi)[R], const float (&yi)[R],
                                                  const float (&zi)[R], uint64_t soft2p, uint64_t (&ax)[R],
                                                  uint64_t (&ay)[R], uint64_t (&az)[R])
{
    const int lane = threadIdx.x & 31;
    for (int g = 0; g < BLK; g += 32) {
        const float xr = sb[g + lane], yr = sb[BLK + g + lane], zr = sb[2 * BLK + g + lane], gr = sb[3 * BLK + g + lane];
#pragma unroll 2
        for (int j = 0; j < 32; j += 2) {
            const uint64_t sj[3] = {pk2(__shfl_sync(0xffffffffu, xr, j), __shfl_sync(0xffffffffu, xr, j + 1)),
                                    pk2(__shfl_sync(0xffffffffu, yr, j), __shfl_sync(0xffffffffu, yr, j + 1)),
                                    pk2(__shfl_sync(0xffffffffu, zr, j), __shfl_sync(0xffffffffu, zr, j + 1))};
            const uint64_t gj = pk2(__shfl_sync(0xffffffffu, gr, j), __shfl_sync(0xffffffffu, gr, j + 1));
#pragma unroll
            for (int k = 0; k < R; ++k) {
                uint64_t acc[3] = {ax[k], ay[k], az[k]};
                pair_interaction<false>(sj, gj, xi[k], yi[k], zi[k], soft2p, acc);
                ax[k] = acc[0]; ay[k] = acc[1]; az[k] = acc[2];
            }
        }
    }
}

// Scalar FFMA/FMUL/FADD version of the same block (13 issue slots per interaction).
template <int R, int U>
__device__ __forceinline__ void block_scalar(const float *__restrict__ sb, const float (&xi)[R], const float (&yi)[R],
                                             const float (&zi)[R], float soft2, float (&ax)[R], float (&ay)[R],
                                             float (&az)[R])
{
#pragma unroll U
    for (int j = 0; j < BLK; j += 4) {
        const float4 xv = *reinterpret_cast<const float4 *>(sb + j);
        const float4 yv = *reinterpret_cast<const float4 *>(sb + BLK + j);
        const float4 zv = *reinterpret_cast<const float4 *>(sb + 2 * BLK + j);
        const float4 gv = *reinterpret_cast<const float4 *>(sb + 3 * BLK + j);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        const float ys[4] = {yv.x, yv.y, yv.z, yv.w};
        const float zs[4] = {zv.x, zv.y, zv.z, zv.w};
        const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const float dx = xs[q] - xi[k];
                const float dy = ys[q] - yi[k];
                const float dz = zs[q] - zi[k];
                float d = fmaf(dx, dx, soft2);
                d = fmaf(dy, dy, d);
                d = fmaf(dz, dz, d);
                const float inv = rsqrt_approx(d);
                const float gi = gs[q] * inv;
                const float i2 = inv * inv;
                const float f = i2 * gi;
                ax[k] = fmaf(f, dx, ax[k]);
                ay[k] = fmaf(f, dy, ay[k]);
                az[k] = fmaf(f, dz, az[k]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- the kernel
// THREADS      threads per CTA (multiple of 32; THREADS*R must be a multiple of BLK)
// R            targets per thread (register blocking)
// TJB          AoSoA blocks (of 128 sources) per pipeline stage
// ST           pipeline stages
// MATH         0 = scalar FP32, 1 = packed f32x2, 2 = packed f32x2 with scalar-FFMA accumulation,
//              3 = packed f32x2 with shuffle-broadcast sources (comparator)
// WARP_PRIVATE each warp owns its stages and barriers (no CTA-wide barrier in the loop)
// U            unroll of the 4-source inner step
// MINB         min resident CTAs per SM for __launch_bounds__
// CL           thread-block cluster size along the chunk axis (launch with cluster dims (1, CL, 1); gridDim.y % CL == 0).
//              CL > 1: the CL CTAs of a cluster work on CL consecutive chunks of the SAME target tile and add their
//              partial sums over distributed shared memory in a fixed order (fp64, rounded once), so the launch writes
//              one partial row per cluster instead of one per CTA: CL times less HBM/L2 traffic for the partial rows
//              and CL times fewer rows for the integrator to add, still deterministic.
template <int THREADS, int R, int TJB, int ST, int MATH, bool WARP_PRIVATE, int U, int MINB, int CL = 1>
__global__ void __launch_bounds__(THREADS, MINB) force_kernel(const ForceArgs a)
{
    constexpr int TI = THREADS * R;
    static_assert(TI % BLK == 0, "target tile must cover whole blocks");
    static_assert(CL == 1 || (!WARP_PRIVATE && (size_t)ST * TJB * BLK_BYTES >= (size_t)3 * TI * 4 && (3 * TI) % CL == 0),
                  "cluster reduction reuses the drained pipeline stages as a [3][TI] float buffer");
    constexpr int GROUPS = WARP_PRIVATE ? THREADS / 32 : 1;
    constexpr int STAGE_FLOATS = TJB * BLK_FLOATS;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *stages = reinterpret_cast<float *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)GROUPS * ST * STAGE_FLOATS * 4);

    const int grp = WARP_PRIVATE ? (int)(threadIdx.x >> 5) : 0;
    const bool leader = WARP_PRIVATE ? ((threadIdx.x & 31) == 0) : (threadIdx.x == 0);
    float *my_stages = stages + (size_t)grp * ST * STAGE_FLOATS;
    const uint32_t my_stages_u32 = smem_u32(my_stages);
    const uint32_t bar0 = smem_u32(bars + grp * ST);

    unsigned long long t_clk0 = 0, t_ns0 = 0;
    if (a.dbg != nullptr && threadIdx.x == 0) {
        t_clk0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ns0));
    }

    // balanced split of the source blocks over the S chunks; partial row = logical chunk index
    const uint32_t c = a.chunk_first + blockIdx.y;
    const uint32_t pc = (c + a.chunk_rot) % a.n_chunks_total;
    uint32_t b_begin, b_end;
    if (a.chunk_tab != nullptr) {
        b_begin = __ldg(a.chunk_tab + 2 * c);
        b_end = __ldg(a.chunk_tab + 2 * c + 1);
    } else {
        b_begin = (uint32_t)(((uint64_t)a.src_nblk_total * pc) / a.n_chunks_total);
        b_end = (uint32_t)(((uint64_t)a.src_nblk_total * (pc + 1)) / a.n_chunks_total);
    }
    const uint32_t nblk = b_end - b_begin;
    const uint32_t ntiles = (nblk + TJB - 1) / TJB;
    const float *chunk_src = a.src + (size_t)b_begin * BLK_FLOATS;

    if (leader) {
#pragma unroll
        for (int s = 0; s < ST; ++s) mbar_init(bar0 + 8 * s, 1);
        fence_mbar_init();
    }
    if (WARP_PRIVATE) __syncwarp(); else __syncthreads();

    auto issue = [&](uint32_t tile, uint32_t stage) {
        const uint32_t nb = min((uint32_t)TJB, nblk - tile * TJB);
        const uint32_t bytes = nb * BLK_BYTES;
        const uint32_t bar = bar0 + 8 * stage;
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(my_stages_u32 + stage * (STAGE_FLOATS * 4), chunk_src + (size_t)tile * STAGE_FLOATS, bytes, bar);
    };
    if (leader) {
#pragma unroll
        for (int s = 0; s < ST; ++s)
            if ((uint32_t)s < ntiles) issue(s, s);
    }

    // register-blocked targets: thread owns local targets  blockIdx.x*TI + k*THREADS + threadIdx.x
    float xi[R], yi[R], zi[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const size_t il = min((size_t)blockIdx.x * TI + (size_t)k * THREADS + threadIdx.x, (size_t)a.tgt_count - 1);
        const size_t ig = (size_t)a.tgt_blk0 * BLK + il;
        xi[k] = __ldg(a.tgt + blk_index(ig, 0));
        yi[k] = __ldg(a.tgt + blk_index(ig, 1));
        zi[k] = __ldg(a.tgt + blk_index(ig, 2));
    }

    float mx[R], my[R], mz[R]; // second-level accumulators
#pragma unroll
    for (int k = 0; k < R; ++k) mx[k] = my[k] = mz[k] = 0.f;

    const uint64_t soft2p = pk2(a.soft2, a.soft2);

    for (uint32_t t = 0; t < ntiles; ++t) {
        const uint32_t s = t % ST;
        const uint32_t parity = (t / ST) & 1u;
        mbar_wait(bar0 + 8 * s, parity);
        const float *sb = my_stages + (size_t)s * STAGE_FLOATS;
        const uint32_t nb = min((uint32_t)TJB, nblk - t * TJB);

        if (MATH != 0) {
            uint64_t ax[R], ay[R], az[R];
#pragma unroll
            for (int k = 0; k < R; ++k) ax[k] = ay[k] = az[k] = 0ull;
            for (uint32_t b = 0; b < nb; ++b) {
                if (MATH == 3) block_packed_shfl<R>(sb + b * BLK_FLOATS, xi, yi, zi, soft2p, ax, ay, az);
                else block_packed<R, U, MATH == 2>(sb + b * BLK_FLOATS, xi, yi, zi, soft2p, ax, ay, az);
            }
#pragma unroll
            for (int k = 0; k < R; ++k) {
                float lo, hi;
                upk2(ax[k], lo, hi); mx[k] += lo + hi;
                upk2(ay[k], lo, hi); my[k] += lo + hi;
                upk2(az[k], lo, hi); mz[k] += lo + hi;
            }
        } else {
            float ax[R], ay[R], az[R];
#pragma unroll
            for (int k = 0; k < R; ++k) ax[k] = ay[k] = az[k] = 0.f;
            for (uint32_t b = 0; b < nb; ++b) block_scalar<R, U>(sb + b * BLK_FLOATS, xi, yi, zi, a.soft2, ax, ay, az);
#pragma unroll
            for (int k = 0; k < R; ++k) { mx[k] += ax[k]; my[k] += ay[k]; mz[k] += az[k]; }
        }

        if (t + ST < ntiles) { // refill the stage we just drained
            if (WARP_PRIVATE) __syncwarp(); else __syncthreads();
            if (leader) issue(t + ST, s);
        }
    }

    if constexpr (CL == 1) {
        const size_t row = (size_t)c * 3;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const size_t il = (size_t)blockIdx.x * TI + (size_t)k * THREADS + threadIdx.x;
            a.partial[(row + 0) * a.tgt_stride + il] = mx[k];
            a.partial[(row + 1) * a.tgt_stride + il] = my[k];
            a.partial[(row + 2) * a.tgt_stride + il] = mz[k];
        }
    } else {
        // cluster reduction over distributed shared memory.  red[comp][off], off = target offset inside the tile.
        __syncthreads(); // every thread is done reading the pipeline stages
        float *red = stages;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int off = k * THREADS + (int)threadIdx.x;
            red[off] = mx[k];
            red[TI + off] = my[k];
            red[2 * TI + off] = mz[k];
        }
        cluster_sync_all(); // all CL buffers are complete and visible cluster-wide
        uint32_t crank;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
        constexpr int COLS = 3 * TI / CL; // columns of the [3][TI] buffer this CTA finishes
        const size_t row = (size_t)(c / CL) * 3;
        const uint32_t red_u32 = smem_u32(red);
        for (int idx = (int)threadIdx.x; idx < COLS; idx += THREADS) {
            const int col = (int)crank * COLS + idx;
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < CL; ++q) s += (double)ld_dsmem_f32(red_u32 + 4u * (uint32_t)col, (uint32_t)q); // fixed order
            const int comp = col / TI, off = col - comp * TI;
            a.partial[(row + comp) * a.tgt_stride + (size_t)blockIdx.x * TI + off] = (float)s;
        }
        cluster_sync_all(); // nobody leaves while a peer may still read its shared memory
    }

    if (a.dbg != nullptr && threadIdx.x == 0) {
        unsigned long long t_ns1;
        const unsigned long long t_clk1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ns1));
        unsigned long long *d = a.dbg + 4ull * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
        d[0] = t_clk0; d[1] = t_clk1; d[2] = t_ns0; d[3] = t_ns1;
    }
}


// =========================================================================================== stream-K decomposition
// The same inner loop, different work decomposition.  Work unit = (target tile, one 128-source block); a launch has
// U = n_itiles * nb units and exactly G CTAs (one per resident slot: SMs x CTAs/SM); CTA c takes the contiguous unit
// range [c*U/G, (c+1)*U/G).  All CTAs are resident at once and get equal work (+-1 unit), so there is no wave
// quantisation and no tail, for any N.  A CTA's range cuts into at most a few (tile, block-range) segments; each
// segment writes ONE partial row (row = ordinal of the segment inside its tile, computed in closed form), so a tile has
// ~2-3 rows instead of one per chunk.  Everything is static: results stay deterministic.
// Accuracy: segments are long (thousands of blocks at N=1M), so the fp32 second-level accumulator is folded into a
// per-CTA fp64 accumulator in shared memory every SK_FLUSH_TILES tiles (FP64 pipe, off the FMA pipe; no registers).
struct ForceArgsSK {
    const float *src;          // blocked bodies the sources are read from
    const float *tgt;          // blocked bodies the targets are read from
    float *partial;            // [rows][3][tgt_stride]
    uint32_t tgt_blk0;         // first target block inside `tgt`
    uint32_t tgt_stride;       // floats per component row of `partial` (whole tiles)
    uint32_t tgt_count;        // local targets; threads of the last tile past it re-read the last target
    uint32_t src_nblk_total;   // NB: blocks of the whole padded system
    uint32_t blk_rot;          // logical -> physical block rotation (rank's first block): phys = (lb + rot) % NB
    uint32_t lb0, nb;          // logical block range [lb0, lb0 + nb) of this launch
    uint32_t n_itiles;         // target tiles of this rank
    uint32_t row0;             // first partial row of this launch
    float soft2;
};

constexpr int SK_FLUSH_TILES = 32;

template <int THREADS, int R, int TJB, int ST>
constexpr size_t force_sk_smem_bytes()
{
    return (size_t)ST * (TJB * BLK_BYTES + 8) + (size_t)3 * THREADS * R * sizeof(double);
}

template <int THREADS, int R, int TJB, int ST, int U, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) force_kernel_sk(const ForceArgsSK a)
{
    constexpr int TI = THREADS * R;
    static_assert(TI % BLK == 0, "target tile must cover whole blocks");
    constexpr int STAGE_FLOATS = TJB * BLK_FLOATS;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *stages = reinterpret_cast<float *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)ST * STAGE_FLOATS * 4);
    double *sacc = reinterpret_cast<double *>(smem_raw + (size_t)ST * (STAGE_FLOATS * 4 + 8)); // [3][TI]

    const bool leader = threadIdx.x == 0;
    const uint32_t stages_u32 = smem_u32(stages);
    const uint32_t bar0 = smem_u32(bars);

    const uint64_t units = (uint64_t)a.n_itiles * a.nb;
    const uint32_t G = gridDim.x, c = blockIdx.x;
    uint64_t u = units * c / G;
    const uint64_t u_end = units * (c + 1) / G;
    if (u >= u_end) return;

    if (leader) {
#pragma unroll
        for (int s = 0; s < ST; ++s) mbar_init(bar0 + 8 * s, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const uint64_t soft2p = pk2(a.soft2, a.soft2);
    uint32_t tcount = 0; // pipeline tiles consumed so far by this CTA (stage and phase bookkeeping across runs)

    while (u < u_end) {
        const uint32_t t = (uint32_t)(u / a.nb);
        const uint32_t off = (uint32_t)(u - (uint64_t)t * a.nb);
        const uint32_t len = (uint32_t)min((uint64_t)(a.nb - off), u_end - u);
        const uint32_t row = a.row0 + (c - sk_cta_of((uint64_t)t * a.nb, units, G));

        float xi[R], yi[R], zi[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const size_t il = min((size_t)t * TI + (size_t)k * THREADS + threadIdx.x, (size_t)a.tgt_count - 1);
            const size_t ig = (size_t)a.tgt_blk0 * BLK + il;
            xi[k] = __ldg(a.tgt + blk_index(ig, 0));
            yi[k] = __ldg(a.tgt + blk_index(ig, 1));
            zi[k] = __ldg(a.tgt + blk_index(ig, 2));
            sacc[(0 * R + k) * THREADS + threadIdx.x] = 0.0;
            sacc[(1 * R + k) * THREADS + threadIdx.x] = 0.0;
            sacc[(2 * R + k) * THREADS + threadIdx.x] = 0.0;
        }
        float mx[R], my[R], mz[R];
#pragma unroll
        for (int k = 0; k < R; ++k) mx[k] = my[k] = mz[k] = 0.f;

        auto flush64 = [&]() { // fold the fp32 second level into the CTA's fp64 accumulators (own slots: no sync needed)
#pragma unroll
            for (int k = 0; k < R; ++k) {
                sacc[(0 * R + k) * THREADS + threadIdx.x] += (double)mx[k];
                sacc[(1 * R + k) * THREADS + threadIdx.x] += (double)my[k];
                sacc[(2 * R + k) * THREADS + threadIdx.x] += (double)mz[k];
                mx[k] = my[k] = mz[k] = 0.f;
            }
        };

        // the segment's logical blocks are contiguous; physically they may wrap around the end of the array once
        const uint32_t p0 = (a.lb0 + off + a.blk_rot) % a.src_nblk_total;
        const uint32_t len0 = min(len, a.src_nblk_total - p0);
        for (int part = 0; part < 2; ++part) {
            const uint32_t pb = part == 0 ? p0 : 0u;
            const uint32_t nblk = part == 0 ? len0 : len - len0;
            if (nblk == 0) continue;
            const float *run_src = a.src + (size_t)pb * BLK_FLOATS;
            const uint32_t ntiles = (nblk + TJB - 1) / TJB;
            auto issue = [&](uint32_t tile) {
                const uint32_t stage = (tcount + tile) % ST;
                const uint32_t nbk = min((uint32_t)TJB, nblk - tile * TJB);
                const uint32_t bytes = nbk * BLK_BYTES;
                const uint32_t bar = bar0 + 8 * stage;
                mbar_arrive_expect_tx(bar, bytes);
                bulk_g2s(stages_u32 + stage * (STAGE_FLOATS * 4), run_src + (size_t)tile * STAGE_FLOATS, bytes, bar);
            };
            __syncthreads(); // every thread is done with the stages of the previous run
            if (leader)
                for (uint32_t s = 0; s < (uint32_t)ST && s < ntiles; ++s) issue(s);

            for (uint32_t tl = 0; tl < ntiles; ++tl) {
                const uint32_t tc = tcount + tl;
                const uint32_t s = tc % ST;
                mbar_wait(bar0 + 8 * s, (tc / ST) & 1u);
                const float *sb = stages + (size_t)s * STAGE_FLOATS;
                const uint32_t nbk = min((uint32_t)TJB, nblk - tl * TJB);
                uint64_t ax[R], ay[R], az[R];
#pragma unroll
                for (int k = 0; k < R; ++k) ax[k] = ay[k] = az[k] = 0ull;
                for (uint32_t b = 0; b < nbk; ++b) block_packed<R, U, false>(sb + b * BLK_FLOATS, xi, yi, zi, soft2p, ax, ay, az);
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    float lo, hi;
                    upk2(ax[k], lo, hi); mx[k] += lo + hi;
                    upk2(ay[k], lo, hi); my[k] += lo + hi;
                    upk2(az[k], lo, hi); mz[k] += lo + hi;
                }
                if ((tl % SK_FLUSH_TILES) == SK_FLUSH_TILES - 1) flush64();
                if (tl + ST < ntiles) { // refill the stage we just drained
                    __syncthreads();
                    if (leader) issue(tl + ST);
                }
            }
            tcount += ntiles;
        }
        flush64();

        const size_t prow = (size_t)row * 3;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const size_t il = (size_t)t * TI + (size_t)k * THREADS + threadIdx.x;
            a.partial[(prow + 0) * a.tgt_stride + il] = (float)sacc[(0 * R + k) * THREADS + threadIdx.x];
            a.partial[(prow + 1) * a.tgt_stride + il] = (float)sacc[(1 * R + k) * THREADS + threadIdx.x];
            a.partial[(prow + 2) * a.tgt_stride + il] = (float)sacc[(2 * R + k) * THREADS + threadIdx.x];
        }
        u += len;
    }
}

} // namespace b200nb
