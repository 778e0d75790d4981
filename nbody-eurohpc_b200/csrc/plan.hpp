// Host-side launch planning for the 2-D (target tiles x source chunks) force grid.
// The reference sizes its grid as ceil(N/1024) CTAs (SimulationNBodyCUDATileFullDevice.cu:181), which leaves
// 100 of 148 SMs half-loaded at N=200k; here the number of source chunks is chosen so that the CTA count is
// as close as possible to a whole number of waves over (SM count x resident CTAs per SM).
#pragma once
#include <algorithm>
#include <cstdint>

namespace b200nb {

struct ChunkPlan {
    uint32_t n_chunks;
    double wave_efficiency; // CTAs / (waves * slots)
    uint32_t waves;
};

// n_itiles: target tiles; n_src_blocks: AoSoA source blocks; slots: SMs * resident CTAs/SM;
// min_blocks_per_chunk: keep the TMA pipeline busy; max_chunks: bounds the partial-sum buffer.
inline ChunkPlan plan_chunks(uint32_t n_itiles, uint32_t n_src_blocks, uint32_t slots, uint32_t min_blocks_per_chunk,
                             uint32_t max_chunks)
{
    ChunkPlan best{1, 0.0, 1};
    const uint32_t s_hi = std::max(1u, std::min(max_chunks, n_src_blocks / std::max(1u, min_blocks_per_chunk)));
    double best_score = -1.0;
    for (uint32_t s = 1; s <= s_hi; ++s) {
        const uint64_t m = (uint64_t)n_itiles * s;
        const uint64_t waves = (m + slots - 1) / slots;
        const double eff = (double)m / (double)(waves * slots);
        // chunk sizes differ by at most one block: account for the longest chunk
        const double ragged = (double)n_src_blocks / (double)(s * ((n_src_blocks + s - 1) / s));
        // prefer >= 4 waves (dynamic CTA scheduling evens out SM speed differences), then fewer chunks
        const double wave_bonus = waves >= 4 ? 0.0 : -0.02 * (double)(4 - waves);
        const double score = eff * ragged + wave_bonus - 1e-4 * s;
        if (score > best_score) {
            best_score = score;
            best = ChunkPlan{s, eff * ragged, (uint32_t)waves};
        }
    }
    return best;
}

} // namespace b200nb
