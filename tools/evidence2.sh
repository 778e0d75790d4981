#!/usr/bin/env bash
# BASELINE configs[0] at full length (one host core, ~15 min) beside GPU-side checks of the latest change.
set -u
O=gpurun_out
( time oracle/_ref/murb_b200 -n 30000 -i 200 --nv --im cpu+naive --gf ) > $O/r02_config0_cpu_naive_full.txt 2>&1 &
NAIVE=$!
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for args in "--n 200000 --reps 5 --filter t32_r8" "--n 200000 --targets 25088 --reps 5 --filter t32_r8" "--n 1000000 --targets 125184 --reps 3 --filter t32_r8"; do echo "## kbench_r2 --micro 0 $args"; timeout 300 build/kbench_r2 --micro 0 --out $O/kb_r2b.jsonl $args; done > $O/r02_kbench_t32_stages.txt 2>&1
for s in 28 37 46 56 66 80 100 120; do build/kbench_r2 --micro 0 --n 200000 --reps 5 --filter t32_r8_tj2_st2 --chunks $s --out $O/kb_r2c.jsonl | grep "^pk_t32_r8_tj2_st2_cta_u1_mb8 "; done > $O/r02_kbench_t32_chunks.txt 2>&1
cat $O/r02_kbench_t32_stages.txt $O/r02_kbench_t32_chunks.txt
python bench.py --steps 20 --warmup 5 > $O/r02_bench_c.json 2> $O/r02_bench_c.err; echo "bench rc=$?"
wait $NAIVE
grep -E "Entire simulation|real" $O/r02_config0_cpu_naive_full.txt
