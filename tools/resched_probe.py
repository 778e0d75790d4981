#!/usr/bin/env python3
"""Where does the re-ordered force kernel differ from the ptxas-scheduled one?  (development probe)"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "nbody-eurohpc_b200"))
import b200nb  # noqa: E402


def run(n, devices, scheme, no_resched, steps=0):
    if no_resched:
        os.environ["B200NB_NO_RESCHED"] = "1"
    else:
        os.environ.pop("B200NB_NO_RESCHED", None)
    d = b200nb.init_bodies(scheme, n)
    with b200nb.Context(n, b200nb.G_F32, 2e8, devices=devices) as ctx:
        ctx.upload(*[d[k] for k in ("qx", "qy", "qz", "m", "vx", "vy", "vz")])
        if steps:
            ctx.step(3600.0, 0, steps)
        ctx.accel()
        return ctx.kernel_name, np.stack(ctx.download_accel())


for n, shards, scheme, steps in [(100000, 8, "random", 0), (100000, 8, "random", 3), (100000, 1, "random", 0), (100000, 4, "random", 0),
                                 (50000, 8, "random", 0), (200000, 8, "galaxy", 0), (12544, 1, "random", 0), (25088, 2, "random", 0),
                                 (100000, 8, "random", 0)]:
    ka, a = run(n, [0] * shards, scheme, False, steps)
    kb, b = run(n, [0] * shards, scheme, True, steps)
    diff = a.view(np.uint32) != b.view(np.uint32)
    bad = np.where(diff.any(axis=0))[0]
    rel = np.abs(a - b).max() / np.abs(b).max()
    print(f"n={n:7d} shards={shards} steps={steps} {ka:45s}: {len(bad):6d} targets differ bitwise, max rel {rel:.2e}"
          + (f"; first {bad[:6]}, last {bad[-3:]}, L={b200nb.slice_length(n, shards)}" if len(bad) else ""), flush=True)
