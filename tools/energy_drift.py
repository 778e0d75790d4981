#!/usr/bin/env python3
"""BASELINE.json configs[2]: N=200000, leapfrog, 1000 iterations, energy drift — B200 leapfrog vs B200 MUrB-explicit with
the on-device fp64 energy (b200nb_energy), plus the CPU cross-check at small N against the oracle's fp64-force drivers
(cpu+naive at N=200k x 1000 it would take days).  Writes a CSV per run.

    python tools/energy_drift.py [--bodies 200000] [--iters 1000] [--every 50] [--out profiles/r01_energy_drift_200k.csv]
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "nbody-eurohpc_b200"))
import b200nb  # noqa: E402

SOFT, DT = 2e8, 3600.0


def drift_gpu(n, scheme, iters, every):
    d = b200nb.init_bodies(scheme, n)
    rows = {}
    for name, integ in (("murb_explicit", 0), ("leapfrog_kdk", 1)):
        with b200nb.Context(n, b200nb.G_F32, SOFT, 1) as ctx:
            ctx.upload(d["qx"], d["qy"], d["qz"], d["m"], d["vx"], d["vy"], d["vz"])
            e = [ctx.energy()]
            for _ in range(iters // every):
                ctx.step(DT, integ, every)
                e.append(ctx.energy())
        rows[name] = np.array(e)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bodies", type=int, default=200000)
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--every", type=int, default=50)
    ap.add_argument("--scheme", default="galaxy")
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "energy_drift.csv"))
    ap.add_argument("--cpu-check", type=int, default=2048, help="also compare with the oracle at this N (0 = skip)")
    args = ap.parse_args()
    rows = drift_gpu(args.bodies, args.scheme, args.iters, args.every)
    e0 = rows["murb_explicit"][0]
    with open(args.out, "w") as f:
        f.write("iteration,energy_murb_explicit,energy_leapfrog_kdk,rel_drift_murb_explicit,rel_drift_leapfrog_kdk\n")
        for k in range(len(rows["murb_explicit"])):
            a, b = rows["murb_explicit"][k], rows["leapfrog_kdk"][k]
            f.write(f"{k * args.every},{a:.17g},{b:.17g},{(a - e0) / abs(e0):.6e},{(b - e0) / abs(e0):.6e}\n")
    for name, e in rows.items():
        print(f"N={args.bodies} {args.scheme} {name}: E0={e[0]:.9e}  max|dE/E0| over {args.iters} it = {np.max(np.abs((e - e[0]) / e[0])):.3e}")
    if args.cpu_check:
        import importlib.util
        spec = importlib.util.spec_from_file_location("pyoracle", os.path.join(REPO, "oracle", "pyoracle.py"))
        pyoracle = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(pyoracle)
        oracle = pyoracle.load()
        n = args.cpu_check
        gpu = drift_gpu(n, args.scheme, args.iters, args.every)
        for name, integ in (("murb_explicit", 0), ("leapfrog_kdk", 1)):
            d = oracle.init_bodies(args.scheme, n)
            e = [oracle.energy(d)]
            for _ in range(args.iters // args.every):
                oracle.run_f64force(d, args.every, integ)
                e.append(oracle.energy(d))
            e = np.array(e)
            dg = (gpu[name] - gpu[name][0]) / abs(gpu[name][0])
            dc = (e - e[0]) / abs(e[0])
            print(f"N={n} {name}: max|dE/E0| B200 {np.max(np.abs(dg)):.3e}  oracle(fp64 force) {np.max(np.abs(dc)):.3e}  "
                  f"max |difference of the two drift curves| {np.max(np.abs(dg - dc)):.3e}")


if __name__ == "__main__":
    main()
