#!/usr/bin/env python3
"""Determinism probe: the same multi-shard run twice with each kernel build (development probe)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "nbody-eurohpc_b200"))
import b200nb  # noqa: E402


def run(n, devices, scheme, no_resched, steps, sync_each=False, split=False):
    os.environ.pop("B200NB_NO_RESCHED", None)
    os.environ.pop("B200NB_SPLIT_LAUNCHES", None)
    if no_resched:
        os.environ["B200NB_NO_RESCHED"] = "1"
    if split:
        os.environ["B200NB_SPLIT_LAUNCHES"] = "1"
    d = b200nb.init_bodies(scheme, n)
    with b200nb.Context(n, b200nb.G_F32, 2e8, devices=devices) as ctx:
        ctx.upload(*[d[k] for k in ("qx", "qy", "qz", "m", "vx", "vy", "vz")])
        for _ in range(steps):
            ctx.step(3600.0, 0, 1)
            if sync_each:
                ctx.sync()
        st = ctx.download_state()
        return np.stack([st[k] for k in ("qx", "qy", "qz")])


n, shards = 100000, 8
ref = run(n, [0], "random", True, 3)
for label, kw in [("ptxas   async", dict(no_resched=True)), ("ptxas   async", dict(no_resched=True)), ("resched async", dict(no_resched=False)),
                  ("resched async", dict(no_resched=False)), ("resched sync ", dict(no_resched=False, sync_each=True)),
                  ("resched split", dict(no_resched=False, split=True)), ("ptxas   split", dict(no_resched=True, split=True)),
                  ("resched 1shard", None)]:
    if kw is None:
        out = run(n, [0], "random", False, 3)
    else:
        out = run(n, [0] * shards, "random", steps=3, **kw)
    d = np.abs(out.astype(np.float64) - ref).max(axis=0)
    scale = np.abs(ref).max()
    print(f"{label}: max |dq| / scale vs 1-shard ptxas = {d.max() / scale:.2e}, bodies off by > 1e-6 scale: {(d > 1e-6 * scale).sum()}", flush=True)
