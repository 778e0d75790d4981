// Host-side launch planning for the 2-D (target tiles x source chunks) force grid.
// The reference sizes its grid as ceil(N/1024) CTAs (SimulationNBodyCUDATileFullDevice.cu:181), which leaves
// 100 of 148 SMs half-loaded at N=200k; here the source axis is cut into chunks so that the CTA count is many whole
// waves over (SM count x resident CTAs per SM), independently of N.
#pragma once
#include <algorithm>
#include <cstdint>

#ifdef __CUDACC__
#define B200NB_HD __host__ __device__ __forceinline__
#else
#define B200NB_HD inline
#endif

namespace b200nb {

// ---- stream-K ownership arithmetic (used by force_kernel_sk, integrate_kernel and the host planner) ----
// CTA that owns unit x when CTA c owns the unit range [c*U/G, (c+1)*U/G)
B200NB_HD uint32_t sk_cta_of(uint64_t x, uint64_t U, uint32_t G) { return (uint32_t)(((x + 1) * G - 1) / U); }
// partial rows tile t receives from a launch with U units, nb blocks per tile, G CTAs
B200NB_HD uint32_t sk_rows_of_tile(uint32_t t, uint32_t nb, uint64_t U, uint32_t G)
{
    const uint64_t first = (uint64_t)t * nb;
    return sk_cta_of(first + nb - 1, U, G) - sk_cta_of(first, U, G) + 1;
}

struct ChunkPlan {
    uint32_t n_chunks;       // chunks per slice (k); the grid has k * n_ranks chunks in total
    uint32_t waves;          // CTA waves of one force pass (both launches when n_ranks > 1)
    double wave_efficiency;  // CTAs / (waves * slots)
    double cta_block_times;  // modelled duration of the pass in units of "one CTA streaming one 128-source block"
};

// Cost model fitted to the B200 chunk sweep (profiles/): one pass takes (waves + 1/2) CTA-times — the CTA scheduler
// refills SMs dynamically, so only about half a CTA of tail is lost — and a CTA costs its blocks plus ~0.3 block of
// prologue/epilogue (pipeline fill, target loads, partial stores).  Many short CTAs (>= ~40 waves) win until the
// per-CTA overhead catches up.
//   n_itiles          target tiles of one rank
//   blocks_per_slice  128-body source blocks in one rank's slice
//   slots             SMs x resident CTAs per SM
//   n_ranks           the own-slice chunks [0,k) and the remote chunks [k, k*n_ranks) are separate launches
inline ChunkPlan plan_chunks(uint32_t n_itiles, uint32_t blocks_per_slice, uint32_t slots, uint32_t n_ranks,
                             uint32_t max_rows, uint32_t min_blocks_per_chunk = 8)
{
    const uint32_t k_hi = std::max(1u, std::min(max_rows / std::max(1u, n_ranks), blocks_per_slice / std::max(1u, min_blocks_per_chunk)));
    ChunkPlan best{1, 1, 0.0, 0.0};
    double best_t = 1e300;
    for (uint32_t k = 1; k <= k_hi; ++k) {
        const uint64_t m_own = (uint64_t)n_itiles * k, m_rem = m_own * (n_ranks - 1);
        const uint64_t w_own = (m_own + slots - 1) / slots, w_rem = (m_rem + slots - 1) / slots;
        const double bpc = (double)blocks_per_slice / k;
        const double tails = n_ranks > 1 ? 1.0 : 0.5; // two launches, two tails
        const double t = ((double)(w_own + w_rem) + tails) * (bpc + 0.3);
        if (t < best_t) {
            best_t = t;
            best = ChunkPlan{k, (uint32_t)(w_own + w_rem), (double)(m_own + m_rem) / (double)((w_own + w_rem) * slots), t};
        }
    }
    return best;
}

// ---- choosing between kernel variants (host only) ----
// Shape of a force-kernel variant as the planner sees it.  rate = measured steady-state interactions / clk / SM.
struct VariantShape {
    uint32_t threads, r, tjb, occ; // occ: resident CTAs per SM (cudaOccupancyMaxActiveBlocksPerMultiprocessor); 0 = cannot run
    double rate;
};

// Modelled duration of one force pass of a rank (arbitrary units, comparable between variants):
// (CTA-block-times of the plan) x (interactions per CTA-block) / (per-CTA rate = SM rate / resident CTAs).
inline double modelled_pass_time(const VariantShape &v, uint64_t L, uint32_t blocks_per_slice, uint32_t n_sms, uint32_t n_ranks,
                                 uint32_t max_rows)
{
    const uint32_t ti = v.threads * v.r;
    const ChunkPlan p = plan_chunks((uint32_t)((L + ti - 1) / ti), blocks_per_slice, n_sms * v.occ, n_ranks, max_rows, 2 * v.tjb);
    return p.cta_block_times * (double)ti * (double)v.occ / v.rate;
}

// Index of the variant expected to finish first.  Candidates are listed in order of preference: a later one replaces
// the incumbent only if the model expects it to be clearly faster - by more than 0.5 % between variants with the same
// inner loop (same R), by more than 2.5 % when the inner loop changes (whole-wave quantisation in the model is worth
// that much; measured: at 25 088 targets x 200k sources the model favours R = 2 by 1.7 % and R = 8 is 2.3 % faster).
inline int choose_variant_index(const VariantShape *v, int count, uint64_t L, uint32_t blocks_per_slice, uint32_t n_sms,
                                uint32_t n_ranks, const uint32_t *max_rows, double *times = nullptr)
{
    int best = -1;
    double best_t = 1e300;
    for (int i = 0; i < count; ++i) {
        if (times) times[i] = -1.0;
        if (v[i].occ < 1) continue;
        const double t = modelled_pass_time(v[i], L, blocks_per_slice, n_sms, n_ranks, max_rows[i]);
        if (times) times[i] = t;
        const double margin = (best >= 0 && v[i].r != v[best].r) ? 0.975 : 0.995;
        if (best < 0 || t < margin * best_t) { best = i; best_t = t; }
    }
    return best;
}

} // namespace b200nb
