#!/usr/bin/env bash
# ncu evidence for the re-ordered default kernel (run under gpurun on ONE GPU; every profiled command first exits 0 plain)
set -u
O=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-scaling-base --no-side-legs"
$B > $O/ncu_plain_resched.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_ncu_launches_resched_bench200k.csv $B > $O/ncu_launches_resched.log 2>&1
$B > $O/ncu_plain_resched_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 3 -c 1 -f -o $O/r02_force_200k_resched $B > $O/ncu_full_resched.log 2>&1
B1="$B --bodies 1000000"
$B1 > $O/ncu_plain_resched_1m.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:force_kernel -s 3 -c 1 -f -o $O/r02_force_1m_resched $B1 > $O/ncu_full_resched_1m.log 2>&1
ls -la $O/*resched*.ncu-rep
