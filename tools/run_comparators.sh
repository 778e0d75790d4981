#!/usr/bin/env bash
# On-box prior-art bar: the reference's own gpu+tile+full / gpu+tile+full200k kernels recompiled for
# sm_100a, driven by the same patched CLI as gpu+b200, same ICs, same flags.  Needs oracle/_ref/murb_b200.
B=oracle/_ref/murb_b200
[ -x $B ] || { echo "no $B"; exit 0; }
for cfg in "200000 200" "1000000 10"; do
  set -- $cfg
  for tag in gpu+tile+full gpu+tile+full200k gpu+b200 gpu+b200+leapfrog; do
    printf "%-20s n=%-8s i=%-4s : " $tag $1 $2
    $B -n $1 -i $2 --nv --im $tag --gf | grep "Entire simulation"
  done
done
