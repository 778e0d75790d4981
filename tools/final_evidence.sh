#!/usr/bin/env bash
# Round-end evidence on ONE B200 (run under gpurun): every BASELINE.json config once on the final code.
#   configs[1] bench.py default (200 steps) + the reference arm, configs[3] bench.py at N = 1M, the CLI comparator table,
#   configs[2] N = 200k leapfrog x 1000 energy drift, configs[0] `murb -n 30000 -i 200 --im cpu+naive` AT FULL LENGTH
#   (one host core, ~15 min, runs beside the GPU work of configs[2]).
set -u
O=gpurun_out
python bench.py > $O/r02_bench_1gpu_200k_200steps.json 2> $O/r02_bench_200steps.err; echo "bench default rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_ref.err; echo "reference arm rc=$?"
python bench.py --bodies 1000000 --steps 10 --warmup 3 --no-scaling-base > $O/r02_bench_1m.json 2> $O/r02_bench_1m.err; echo "bench 1M rc=$?"
python bench.py --scheme random --steps 20 --warmup 5 --no-cpu --no-scaling-base --no-side-legs > $O/r02_bench_random_200k.json 2>> $O/r02_bench_1m.err; echo "bench random rc=$?"
bash tools/run_comparators.sh > $O/r02_comparators_cli.txt 2>&1; echo "comparators rc=$?"
( oracle/_ref/murb_b200 -n 30000 -i 200 --nv --im cpu+naive --gf > $O/r02_config0_cpu_naive_full.txt 2>&1 ) &
NAIVE=$!
python tools/energy_drift.py --bodies 200000 --iters 1000 --every 50 --out $O/r02_energy_drift_200k.csv > $O/r02_energy_drift.txt 2>&1; echo "energy drift rc=$?"
oracle/_ref/murb_b200 -n 30000 -i 200 --nv --im gpu+b200 --gf 2>&1 | grep "Entire simulation" > $O/r02_config0_gpu_b200.txt
wait $NAIVE
grep -E "Entire simulation|Elapsed|Maximum resident" $O/r02_config0_cpu_naive_full.txt; cat $O/r02_config0_gpu_b200.txt; tail -4 $O/r02_energy_drift.txt
