#!/usr/bin/env bash
# On-box comparators through the SAME patched CLI, same ICs, same flags (BASELINE configs[0..1]):
#  * the reference's CPU paths (cpu+naive is BASELINE configs[0]; shortened iteration counts, stated),
#  * the reference's own gpu+tile+full / gpu+tile+full200k kernels recompiled for sm_100a,
#  * gpu+b200 / gpu+b200+leapfrog.
# Needs oracle/_ref/murb_b200 (built where the reference sources are available).
B=oracle/_ref/murb_b200
[ -x $B ] || { echo "no $B"; exit 0; }
export OMP_NUM_THREADS=$(nproc) OMP_DYNAMIC=FALSE OMP_PLACES=cores OMP_PROC_BIND=close OMP_SCHEDULE=static OMP_WAIT_POLICY=ACTIVE
echo "host: $(nproc) threads, $(grep -m1 'model name' /proc/cpuinfo | cut -d: -f2)   (CLI built with the reference's shipped flags: -O3 -ffast-math, no -march)"
for cfg in "cpu+naive 30000 3" "cpu+simd 30000 10" "cpu+omp 30000 100" "gpu+b200 30000 200"; do
  set -- $cfg
  printf "%-20s n=%-8s i=%-4s : " $1 $2 $3
  $B -n $2 -i $3 --nv --im $1 --gf | grep "Entire simulation"
done
for cfg in "200000 200" "1000000 10"; do
  set -- $cfg
  for tag in gpu+tile+full gpu+tile+full200k gpu+b200 gpu+b200+leapfrog; do
    printf "%-20s n=%-8s i=%-4s : " $tag $1 $2
    $B -n $1 -i $2 --nv --im $tag --gf | grep "Entire simulation"
  done
done
