#!/usr/bin/env python3
"""Which targets carry the largest |a - a_fp64| / |a_fp64|?  (development probe for bench.py's `parity` record)"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "nbody-eurohpc_b200"))
sys.path.insert(0, REPO)
import b200nb  # noqa: E402
import bench  # noqa: E402

pyoracle = bench.load_pyoracle()
oracle = pyoracle.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
scheme = sys.argv[2] if len(sys.argv) > 2 else "galaxy"
d = b200nb.init_bodies(scheme, n)
idx = np.unique(np.concatenate([[0, 1, 2, n - 1], np.random.default_rng(11).integers(0, n, 400)])).astype(np.uint64)
ii = idx.astype(np.int64)
with b200nb.Context(n, b200nb.G_F32, 2e8, 1) as ctx:
    ctx.upload(*[d[k] for k in ("qx", "qy", "qz", "m", "vx", "vy", "vz")])
    for steps in (0, 10, 40):
        if steps:
            ctx.step(3600.0, 0, steps)
        ctx.accel()
        st, acc = ctx.download_state(), ctx.download_accel()
        moved = dict(d)
        moved.update({k: st[k] for k in ("qx", "qy", "qz")})
        a64 = np.stack(oracle.accel_f64(moved, idx))
        got = np.stack([a[ii] for a in acc]).astype(np.float64)
        err = np.linalg.norm(got - a64, axis=0) / np.linalg.norm(a64, axis=0)
        order = np.argsort(-err)[:6]
        print(f"after {steps:3d} more steps: max {err.max():.2e}, median {np.median(err):.2e}; worst targets:",
              [(int(idx[k]), f"{err[k]:.1e}", f"|a|={np.linalg.norm(a64[:, k]):.2e}", f"r={np.linalg.norm([moved[c][idx[k]] for c in ('qx','qy','qz')]):.2e}") for k in order])
