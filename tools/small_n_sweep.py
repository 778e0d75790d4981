#!/usr/bin/env python3
"""Iteration time versus N for the automatic variant choice and for each forced variant (B200NB_VARIANT)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "nbody-eurohpc_b200"))
import b200nb  # noqa: E402

VARIANTS = (None, "pk_t128_r8_tj2_st3_cta_u1_mb2", "pk_t128_r2_tj1_st3_cta_u2_mb4")
for n in (1000, 2048, 8192, 16384, 30000, 50000, 100000, 200000):
    d = b200nb.init_bodies("galaxy", n)
    for var in VARIANTS:
        if var:
            os.environ["B200NB_VARIANT"] = var
        else:
            os.environ.pop("B200NB_VARIANT", None)
        with b200nb.Context(n, b200nb.G_F32, 2e8, 1) as c:
            c.upload(d["qx"], d["qy"], d["qz"], d["m"], d["vx"], d["vy"], d["vz"])
            c.step(3600.0, 0, 3)
            c.sync()
            iters = 200 if n <= 50000 else 20
            t0 = time.perf_counter()
            c.step(3600.0, 0, iters)
            c.sync()
            dt = time.perf_counter() - t0
            print(f"n={n:7d} {(var or 'auto'):32s} -> {c.kernel_name:30s}: {dt / iters * 1e6:9.1f} us/iter {n * n * iters / dt / 1e9:8.1f} G-int/s", flush=True)
