#!/usr/bin/env bash
# 8-GPU measurements for profiles/ (run under gpurun --gpus 8): the strong-scaling target workload (4M) and the two
# small workloads where per-GPU slices get short (1M, 200k), one process per GPU; then the patched CLI driving 8 GPUs
# from one process.
set -u
O=gpurun_out
nvidia-smi -L > $O/r02_scale8_gpus.txt
run() { # name, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 ${@:3} \
      > $O/r02_bench_8gpu_$2.json 2> $O/r02_bench_8gpu_$2.err; echo "$2 rc=$?"
}
run 29601 4m --steps 5 --warmup 3
run 29602 1m --bodies 1000000 --steps 20 --warmup 5
run 29603 200k --bodies 200000 --steps 50 --warmup 10
MURB_B200_NGPUS=8 oracle/_ref/murb_b200 -n 4194304 -i 5 --nv --im gpu+b200 --gf 2>&1 | tail -3 > $O/r02_cli_8gpu_4m.txt
MURB_B200_NGPUS=8 oracle/_ref/murb_b200 -n 200000 -i 200 --nv --im gpu+b200 --gf 2>&1 | tail -2 >> $O/r02_cli_8gpu_4m.txt
cat $O/r02_cli_8gpu_4m.txt
