#!/usr/bin/env bash
# ncu evidence for profiles/ (run under gpurun on ONE GPU): every profiled command first exits 0 without ncu.
#   1. launch list of a short bench.py run (per-launch device times: the kernel's SHARE of a step)
#   2. ncu --set full of the force kernel inside bench.py at N = 200 000 (default variant) and N = 1 000 000
#   3. the same for the one-warp R = 8 variant sharded runs use
set -u
O=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-scaling-base --no-side-legs"
$B > $O/ncu_plain_200k.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_ncu_launches_bench200k.csv $B > $O/ncu_launches.log 2>&1
$B > $O/ncu_plain_200k_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 3 -c 2 -f -o $O/r02_force_200k $B > $O/ncu_full_200k.log 2>&1
B1="$B --bodies 1000000"
$B1 > $O/ncu_plain_1m.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 3 -c 1 -f -o $O/r02_force_1m $B1 > $O/ncu_full_1m.log 2>&1
export B200NB_VARIANT=pk_t32_r8_tj2_st2_cta_u1_mb8
$B > $O/ncu_plain_200k_t32.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 3 -c 1 -f -o $O/r02_force_200k_t32 $B > $O/ncu_full_200k_t32.log 2>&1
ls -la $O/*.ncu-rep
tail -3 $O/ncu_full_200k.log
