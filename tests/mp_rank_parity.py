"""Multi-process parity worker (launched by tests/test_gpu_parity.py::test_torchrun_ranks_match_single_gpu through
`python -m torch.distributed.run --nproc-per-node P`): every rank owns one GPU and one shard (b200nb_create_rank, NCCL id
broadcast over torch.distributed); rank 0 also runs the same problem on a single-GPU context and compares state,
accelerations and energy.  Exercises the ncclAllGather exchange, the collective downloads and the energy all-reduce."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "nbody-eurohpc_b200"))
import b200nb  # noqa: E402
from b200nb import dist as bdist  # noqa: E402

SOFT, DT = 2e8, 3600.0


def main():
    rank, world, local_rank = bdist.env_rank()
    dist = bdist.init_process_group()
    ok = True
    for n, scheme in ((30000, "galaxy"), (4097, "random")):
        d = b200nb.init_bodies(scheme, n)
        for integ in (0, 1):
            # an ncclUniqueId creates exactly one communicator: a fresh one per context
            nccl_id = bdist.broadcast_bytes(dist, b200nb.Context.unique_id() if rank == 0 else None)
            ctx = b200nb.Context(n, b200nb.G_F32, SOFT, rank=rank, n_ranks=world, device=local_rank, nccl_id=nccl_id)
            ctx.upload(d["qx"], d["qy"], d["qz"], d["m"], d["vx"], d["vy"], d["vz"])
            ctx.step(DT, integ, 4)
            state = ctx.download_state()      # collective
            acc = ctx.download_accel()        # collective
            energy = ctx.energy()             # collective
            metrics = ctx.metrics()           # collective (12-row all-reduce)
            ctx.close()
            if rank == 0:
                with b200nb.Context(n, b200nb.G_F32, SOFT, rank=0, n_ranks=1, device=local_rank) as one:
                    one.upload(d["qx"], d["qy"], d["qz"], d["m"], d["vx"], d["vy"], d["vz"])
                    one.step(DT, integ, 4)
                    s1, a1, e1, m1 = one.download_state(), one.download_accel(), one.energy(), one.metrics()
                scale = max(float(np.abs(s1[c]).max()) for c in ("qx", "qy", "qz"))
                for c in ("qx", "qy", "qz"):
                    ok &= bool(np.all(np.abs(state[c].astype(np.float64) - s1[c]) <= 1e-6 * scale))
                for c in ("vx", "vy", "vz"):
                    ok &= bool(np.all(np.abs(state[c].astype(np.float64) - s1[c]) <= 1e-5 * float(np.abs(s1[c]).max())))
                num = np.linalg.norm(np.stack(acc).astype(np.float64) - np.stack(a1), axis=0)
                den = np.linalg.norm(np.stack(a1).astype(np.float64), axis=0)
                ok &= bool(np.max(num / den) <= 2e-6)
                ok &= abs(energy - e1) <= 1e-6 * abs(e1)
                l1 = float(np.linalg.norm([m1[k] for k in ("ang_x", "ang_y", "ang_z")]))
                ok &= all(abs(metrics[k] - m1[k]) <= 1e-5 * l1 for k in ("ang_x", "ang_y", "ang_z"))
                ok &= all(abs(metrics[k] - m1[k]) <= 1e-5 * scale for k in ("com_x", "com_y", "com_z", "density_x", "density_y", "density_z"))
                ok &= abs(metrics["mass"] - m1["mass"]) <= 1e-12 * m1["mass"]
                print(f"n={n} {scheme} integrator={integ}: max|da|/|a| {np.max(num / den):.2e}, dE {abs(energy - e1) / abs(e1):.1e}, ok={ok}", flush=True)
    # large N, where only sampled targets can be checked on the CPU: fp64 all-pairs oracle (tests/conftest.py)
    import importlib.util
    spec = importlib.util.spec_from_file_location("pyoracle", os.path.join(REPO, "oracle", "pyoracle.py"))
    pyoracle = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pyoracle)
    max_rel_err = pyoracle.max_rel_err
    n, scheme = 500000, "random"
    d = b200nb.init_bodies(scheme, n)
    nccl_id = bdist.broadcast_bytes(dist, b200nb.Context.unique_id() if rank == 0 else None)
    ctx = b200nb.Context(n, b200nb.G_F32, SOFT, rank=rank, n_ranks=world, device=local_rank, nccl_id=nccl_id)
    ctx.upload(d["qx"], d["qy"], d["qz"], d["m"], d["vx"], d["vy"], d["vz"])
    ctx.step(DT, 0, 2)          # two exchanges, so the second force pass reads gathered positions
    ctx.accel()
    state, acc = ctx.download_state(), ctx.download_accel()
    ctx.close()
    if rank == 0:
        oracle = pyoracle.load()
        idx = np.unique(np.concatenate([[0, n // world - 1, n // world, n - 1], np.random.default_rng(5).integers(0, n, 60)])).astype(np.uint64)
        moved = dict(d)
        moved.update({k: state[k] for k in ("qx", "qy", "qz")})
        a64 = oracle.accel_f64(moved, idx)
        err = max_rel_err(a64, [a[idx.astype(np.int64)] for a in acc])
        good = err <= 1e-5
        ok &= good
        print(f"n={n} {scheme} sampled fp64 oracle after 2 sharded steps: max|da|/|a| {err:.2e}, ok={good}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
