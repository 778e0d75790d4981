#!/usr/bin/env python3
"""In-process multi-GPU step time with the two exchange paths: positions pushed over NVLink by the integrator
(B200NB_EXCHANGE=p2p, default) against the in-place ncclAllGather (B200NB_EXCHANGE=nccl).  This is the mode the MUrB
CLI uses (MURB_B200_NGPUS); bench.py --gpus N runs one process per GPU and always takes the NCCL path.
    python tools/exchange_bench.py [n_gpus] [N ...]"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "nbody-eurohpc_b200"))
import b200nb  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sizes = [int(x) for x in sys.argv[2:]] or [30000, 200000, 1000000]
for n in sizes:
    d = b200nb.init_bodies("galaxy", n)
    for integ, iname in ((0, "murb"), (1, "leapfrog")):
        for mode in ("p2p", "nccl"):
            os.environ["B200NB_EXCHANGE"] = mode
            with b200nb.Context(n, b200nb.G_F32, 2e8, n_gpus) as c:
                c.upload(d["qx"], d["qy"], d["qz"], d["m"], d["vx"], d["vy"], d["vz"])
                c.step(3600.0, integ, 5)
                c.sync()
                iters = max(5, min(300, int(2e12 * n_gpus / (float(n) * n))))
                best = 1e30
                for _ in range(3):
                    t0 = time.perf_counter()
                    c.step(3600.0, integ, iters)
                    c.sync()
                    best = min(best, (time.perf_counter() - t0) / iters)
                print(f"n={n:8d} gpus={n_gpus} {iname:8s} {c.exchange_name:15s} {c.kernel_name:30s}: {best * 1e6:10.1f} us/step "
                      f"{float(n) * n / best / 1e9:9.1f} G-int/s", flush=True)
