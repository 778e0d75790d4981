"""Host-side launch planner (nbody-eurohpc_b200/csrc/plan.hpp) and the stream-K row arithmetic
(force_sm100.cuh: sk_cta_of / sk_rows_of_tile), exercised on the CPU through a tiny g++ harness."""
import json
import os
import subprocess

import pytest

from conftest import REPO

HARNESS = r'''
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "plan.hpp"
int main(int argc, char **argv)
{
    using namespace b200nb;
    if (argv[1][0] == 'p') {  // p n_itiles blocks_per_slice slots n_ranks max_rows min_blocks
        ChunkPlan p = plan_chunks(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]));
        printf("{\"k\": %u, \"waves\": %u, \"eff\": %.6f, \"t\": %.3f}\n", p.n_chunks, p.waves, p.wave_efficiency, p.cta_block_times);
    } else if (argv[1][0] == 'v') {  // v L blocks_per_slice n_ranks  (B200: 148 SMs; candidates in the library's order)
        const uint64_t L = strtoull(argv[2], 0, 10);
        const uint32_t bps = atoi(argv[3]), ranks = atoi(argv[4]);
        VariantShape big{128, 8, 2, 2, 9.53}, warp{32, 8, 2, 8, 9.53}, small{128, 2, 1, 4, 9.2};
        VariantShape v[3] = {warp, big, small};
        uint32_t max_rows[3];
        for (int i = 0; i < 3; ++i) {
            const uint64_t ti = (uint64_t)v[i].threads * v[i].r, lp = (L + ti - 1) / ti * ti;
            max_rows[i] = (uint32_t)std::max<uint64_t>(ranks, std::min<uint64_t>(256, (1ull << 30) / (12ull * lp)));
        }
        double t[3];
        const int pick = choose_variant_index(v, 3, L, bps, 148, ranks, max_rows, t);
        printf("{\"pick\": %d, \"threads\": %u, \"r\": %u, \"t\": [%.1f, %.1f, %.1f]}\n", pick, v[pick].threads, v[pick].r, t[0], t[1], t[2]);
    } else {                  // s n_itiles nb G : every unit is owned by exactly one CTA, rows per tile are consistent
        const uint32_t n_itiles = atoi(argv[2]), nb = atoi(argv[3]), G = atoi(argv[4]);
        const uint64_t U = (uint64_t)n_itiles * nb;
        uint64_t bad = 0; uint32_t max_rows = 0;
        for (uint32_t c = 0; c < G; ++c)
            for (uint64_t u = U * c / G; u < U * (c + 1) / G; ++u) bad += sk_cta_of(u, U, G) != c;
        for (uint32_t t = 0; t < n_itiles; ++t) {
            const uint32_t rows = sk_rows_of_tile(t, nb, U, G);
            if (rows > max_rows) max_rows = rows;
        }
        printf("{\"bad\": %llu, \"max_rows\": %u}\n", (unsigned long long)bad, max_rows);
    }
    return 0;
}
'''


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("plan")
    src = d / "plan_harness.cpp"
    src.write_text(HARNESS)
    exe = d / "plan_harness"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(REPO, "nbody-eurohpc_b200", "csrc"), str(src), "-o", str(exe)])
    return lambda *a: json.loads(subprocess.check_output([str(exe), *map(str, a)]).decode())


@pytest.mark.parametrize("n_itiles,blocks,slots,ranks", [(196, 1568, 296, 1), (984, 7872, 296, 1), (2, 16, 296, 1), (512, 4096, 296, 8),
                                                         (26, 208, 296, 8), (4096, 32768, 296, 1)])
def test_plan_chunks_properties(harness, n_itiles, blocks, slots, ranks):
    max_rows, min_blocks = 256, 4
    p = harness("p", n_itiles, blocks, slots, ranks, max_rows, min_blocks)
    assert 1 <= p["k"] <= max(1, min(max_rows // ranks, blocks // min_blocks))
    assert 0 < p["eff"] <= 1.0 + 1e-9
    ideal = n_itiles * ranks * blocks / slots            # CTA-block-times with perfect packing and no overhead
    assert p["t"] >= ideal * 0.999
    if n_itiles * ranks * (blocks // min_blocks) >= 8 * slots:   # enough work: close to ideal; two launches (own /
        assert p["t"] <= (1.06 if ranks == 1 else 1.15) * ideal, p  # remote) have two tails, so sharded plans get more slack
    # never worse than the reference's one-CTA-per-tile grid (k = 1)
    p1 = harness("p", n_itiles, blocks, slots, ranks, ranks, blocks)
    assert p["t"] <= p1["t"] + 1e-6


def test_plan_matches_the_measured_configurations(harness):
    # N=200k, R=8 variant on 148 SMs x 2 CTAs: many waves (profiles/: 42 waves, 63 chunks)
    p = harness("p", 196, 1568, 296, 1, 256, 4)
    assert 40 <= p["k"] <= 80 and p["waves"] >= 24


@pytest.mark.parametrize("n_itiles,nb,G", [(196, 1568, 296), (2, 16, 296), (984, 7872, 296), (7, 13, 5), (1, 1, 296), (26, 1456, 888)])
def test_stream_k_ownership_and_rows(harness, n_itiles, nb, G):
    r = harness("s", n_itiles, nb, G)
    assert r["bad"] == 0
    assert 1 <= r["max_rows"] <= G
    if n_itiles >= G:
        assert r["max_rows"] <= 2   # a tile is split across at most two CTAs when there are more tiles than CTAs


@pytest.mark.parametrize("n,ranks,expect", [(200000, 1, (32, 8)), (1000000, 1, (32, 8)), (4194304, 1, (32, 8)), (2048, 1, (128, 2)),
                                            (8192, 1, (128, 2)), (16384, 1, (128, 2)), (200000, 8, (32, 8)), (4194304, 8, (32, 8)),
                                            (1000000, 8, (32, 8)), (50000, 2, (32, 8)), (4000, 2, (128, 2)), (100000, 1, (32, 8))])
def test_variant_choice(harness, n, ranks, expect):
    """Large systems and sharded runs take the one-warp R = 8 variant (256-target tiles), murb-test sizes the small R = 2
    tiles (measured: profiles/r02_kbench_cluster_smalltiles.txt, profiles/r02_ncu_force_kernel_200k_t32.txt)."""
    L = (-(-n // ranks) + 255) // 256 * 256
    blocks_per_slice = -(-n // 128) if ranks == 1 else L // 128
    v = harness("v", L, blocks_per_slice, ranks)
    assert (v["threads"], v["r"]) == expect, v
