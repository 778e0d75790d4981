#!/usr/bin/env python3
"""bench.py — MUrB all-pairs gravity hot path on B200: billion body-interactions/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--bodies n] [--scheme galaxy|random]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one computeOneIteration(): the N^2 force pass + the MUrB integrator (+ the position all-gather for N > 1),
i.e. one iteration of `murb -n <bodies> -i <steps> --nv --im gpu+b200 --gf`.
  * --gpus 1 : BASELINE.json configs[1]  (murb -n 200000, galaxy, soft 2e8, dt 3600 — the gpu+tile+full200k shape)
  * --gpus >1: BASELINE.json configs[4]  (n = 4,194,304 strong scaling: targets sharded over the ranks, one in-place
               ncclAllGather of the 16 B/body position blocks per step)
`value` is device-timed (CUDA events on the library's compute stream, state resident in HBM); `e2e` goes through the
public C-ABI with pinned HOST buffers: upload (H2D) + step + download (D2H) inside the timed region, every step.
`--impl reference` times the reference's own CPU path (cpu+omp, compiled from /root/reference into oracle/_ref) on the
host cores, on a bounded sample of the workload.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "nbody-eurohpc_b200"))

SOFT, DT = 2e8, 3600.0
STRONG_SCALING_BODIES = 4194304
N_SMS, FP32_LANES = 148, 128
PIPE_SLOTS_PER_INTERACTION = 12  # 3 FADD + 6 FFMA + 3 FMUL; + 1 MUFU.RSQ on its own pipe
METRIC = "billion body-interactions/s (N^2 ordered pairs per force pass, self included)"


def sm_max_mhz():
    try:
        return float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["sm_max_mhz"]), "MEASURED_PEAKS.json sm_max_mhz"
    except Exception:
        return 1965.0, "fallback clocks.max.sm 1965 MHz (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks (NVML)
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons, self.power = [], set(), []
        self._stop = threading.Event()
        self._thread = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def _run_smi(self):
        # fallback without NVML bindings: the recipe's nvidia-smi line (B200_PROFILING.md), polled
        import subprocess
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(int(float(out[0])))
                self.max_mhz = int(float(out[1]))
                self.power.append(float(out[2]))
                for nm, v in zip(names, out[3:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._thread = threading.Thread(target=self._run if self.nv else self._run_smi, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "power_w_max": (max(self.power) if self.power else None), "samples": len(s)}


# ------------------------------------------------------------------------------------------------ CPU reference
def _cpu_flags():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except Exception:
        pass
    return set()


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def load_reference():
    """Best ISA build of the reference CPU path present in oracle/_ref (built from /root/reference by oracle/build_ref.sh)."""
    if os.environ.get("B200NB_BENCH_NO_REF"):  # force the oracle-port fallback (tests)
        return None, None
    flags = _cpu_flags()
    cands = []
    if {"avx512f", "avx512dq", "avx512bw", "avx512vl"} <= flags:
        cands.append(("libmurbref_v4.so", "-O3 -ffast-math -march=x86-64-v4 (AVX-512 MIPP)"))
    if {"avx2", "fma"} <= flags:
        cands.append(("libmurbref_v3.so", "-O3 -ffast-math -march=x86-64-v3 (AVX2 MIPP)"))
    cands.append(("libmurbref.so", "-O3 -ffast-math as shipped (SSE2 MIPP)"))
    FP = ctypes.POINTER(ctypes.c_float)
    for name, desc in cands:
        p = os.path.join(REPO, "oracle", "_ref", name)
        if os.path.exists(p):
            L = ctypes.CDLL(p)
            L.ref_accel.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_float, ctypes.c_int] + [FP] * 3
            L.ref_accel.restype = ctypes.c_double
            L.ref_run.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_float, ctypes.c_float, ctypes.c_int] + [FP] * 9
            L.ref_run.restype = ctypes.c_double
            return L, desc
    return None, None


def omp_env():
    # the reference's own recipe (README.md:78-89, SimulationNBodyOpenMP.cpp:99-109); must be set before libgomp starts
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every worker unless the user set it; the CPU arm runs on rank 0 alone (the
    # other ranks exit), so it takes every core it is allowed to use.  B200NB_BENCH_OMP_THREADS pins it explicitly.
    pinned = os.environ.get("B200NB_BENCH_OMP_THREADS")
    if pinned:
        os.environ["OMP_NUM_THREADS"] = pinned
    elif "TORCHELASTIC_RUN_ID" in os.environ or "LOCAL_RANK" in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    os.environ.setdefault("OMP_DYNAMIC", "FALSE")
    os.environ.setdefault("OMP_PLACES", "cores")
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_SCHEDULE", "static")
    os.environ.setdefault("OMP_WAIT_POLICY", "ACTIVE")
    return int(os.environ["OMP_NUM_THREADS"])


def cpu_reference_rate(scheme, n_sample, budget_s, tag="cpu+omp"):
    """G-int/s of the reference CPU implementation on a bounded sample: whole iterations at n_sample bodies."""
    cores = omp_env()
    L, desc = load_reference()
    if L is None:
        return oracle_port_rate(scheme, budget_s)
    none = [None] * 9
    L.ref_run(tag.encode(), n_sample, scheme.encode(), SOFT, DT, 1, *none)  # warm-up: thread start, page faults
    ms1 = L.ref_run(tag.encode(), n_sample, scheme.encode(), SOFT, DT, 1, *none)
    iters = int(max(2, min(200, budget_s * 1e3 / max(ms1, 1e-3))))
    ms = L.ref_run(tag.encode(), n_sample, scheme.encode(), SOFT, DT, iters, *none)
    rate = float(n_sample) ** 2 * iters / (ms * 1e-3) / 1e9
    out = {"value": rate, "unit": "G-int/s", "cores": cores if tag == "cpu+omp" else 1, "kind": "reference",
           "sample": f"reference {tag} ({desc}), {iters} iterations of murb -n {n_sample} -s {scheme} (full force pass + integrator), "
                     f"{ms / iters:.2f} ms/iter", "ms_per_iter": ms / iters, "iters": iters, "n_sample": n_sample,
           "cpu_model": _cpu_model()}
    if tag == "cpu+omp":
        # the other reference CPU paths the north star asks for, single thread, same ICs (bounded: a few seconds each)
        others = {}
        for t, it in (("cpu+simd", 3), ("cpu+naive", 1)):
            n_t = n_sample if t == "cpu+simd" else 8000   # cpu+naive at 30000 is ~8 s/iteration; 8000 is the Report's own size
            L.ref_run(t.encode(), n_t, scheme.encode(), SOFT, DT, 1 if t == "cpu+simd" else 0, *none)
            ms_t = L.ref_run(t.encode(), n_t, scheme.encode(), SOFT, DT, it, *none)
            others[t] = {"value": float(n_t) ** 2 * it / (ms_t * 1e-3) / 1e9, "unit": "G-int/s", "cores": 1, "n_sample": n_t,
                         "iters": it, "ms_per_iter": ms_t / it}
        # the same cpu+omp exactly as the reference ships it (no -march: SSE2 MIPP, CMakeLists.txt:128-131)
        shipped = os.path.join(REPO, "oracle", "_ref", "libmurbref.so")
        if os.path.exists(shipped) and "as shipped" not in desc:
            FP = ctypes.POINTER(ctypes.c_float)
            S = ctypes.CDLL(shipped)
            S.ref_run.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_float, ctypes.c_float, ctypes.c_int] + [FP] * 9
            S.ref_run.restype = ctypes.c_double
            S.ref_run(b"cpu+omp", n_sample, scheme.encode(), SOFT, DT, 1, *none)
            ms_s = S.ref_run(b"cpu+omp", n_sample, scheme.encode(), SOFT, DT, 10, *none)
            others["cpu+omp as shipped (-O3 -ffast-math, SSE2 MIPP)"] = {
                "value": float(n_sample) ** 2 * 10 / (ms_s * 1e-3) / 1e9, "unit": "G-int/s", "cores": cores, "n_sample": n_sample,
                "iters": 10, "ms_per_iter": ms_s / 10}
        out["other_reference_paths"] = others
    return out


def oracle_port_rate(scheme, budget_s):
    """Fallback when oracle/_ref was not built: the scalar C restatement (oracle/liboracle.so), one core."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pyoracle", os.path.join(REPO, "oracle", "pyoracle.py"))
    pyoracle = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pyoracle)
    o = pyoracle.load()  # builds oracle/liboracle.so if needed
    n = 8000
    d = o.init_bodies(scheme, n)
    t0 = time.perf_counter()
    it = 0
    while time.perf_counter() - t0 < max(2.0, budget_s / 2):
        o.accel_naive(d)
        it += 1
    dt = time.perf_counter() - t0
    return {"value": float(n) ** 2 * it / dt / 1e9, "unit": "G-int/s", "cores": 1, "kind": "port",
            "sample": f"oracle port of cpu+naive (oracle/nbody_oracle.c), {it} force passes at n={n}", "n_sample": n}


# ------------------------------------------------------------------------------------------------ arms
def run_reference_arm(args, rank):
    if rank != 0:
        return
    n_sample = 30000
    workload = f"murb -n {args.bodies} -i {args.steps} --nv --gf ({args.scheme}); timed on a sample of n={n_sample}"
    cores = omp_env()
    L, desc = load_reference()
    if L is None:
        r = oracle_port_rate(args.scheme, 10)
        ms_per_step, value = None, r["value"]
        base = r
    else:
        none = [None] * 9
        tag = b"cpu+omp"
        for _ in range(max(args.warmup, 1)):
            L.ref_run(tag, n_sample, args.scheme.encode(), SOFT, DT, 1, *none)
        ms = L.ref_run(tag, n_sample, args.scheme.encode(), SOFT, DT, args.steps, *none)
        ms_per_step = ms / args.steps
        value = float(n_sample) ** 2 * args.steps / (ms * 1e-3) / 1e9
        base = {"value": value, "unit": "G-int/s", "cores": cores, "kind": "reference",
                "sample": f"reference cpu+omp ({desc}), each step = one full iteration at n={n_sample} ({args.scheme}), {cores} OpenMP threads"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "G-int/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (reference Bodies generator, srand(0))",
        "config": {"workload": workload, "bodies": args.bodies, "scheme": args.scheme, "soft": SOFT, "dt": DT},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "G-int/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


class StdoutToStderr:
    """NCCL (and anything else native) may print to fd 1; the contract is ONE JSON line on stdout.  Everything written
    to fd 1 while this is active goes to stderr; emit() writes the line to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)


def run_b200_arm(args, rank, world, local_rank):
    out_guard = StdoutToStderr()
    import torch
    import b200nb

    from b200nb import dist as bdist

    dist = bdist.init_process_group()     # None for a single rank
    nccl_id = None
    if dist is not None:
        nccl_id = bdist.broadcast_bytes(dist, b200nb.Context.unique_id() if rank == 0 else None)
    torch.cuda.set_device(local_rank)
    torch.cuda.init()
    n = args.bodies
    bodies = b200nb.init_bodies(args.scheme, n)
    if world > 1:
        ctx = b200nb.Context(n, b200nb.G_F32, SOFT, rank=rank, n_ranks=world, device=local_rank, nccl_id=nccl_id)
    else:
        ctx = b200nb.Context(n, b200nb.G_F32, SOFT, 1)
    names = ("qx", "qy", "qz", "m", "vx", "vy", "vz")
    pinned = b200nb.PinnedArrays(list(names), n)
    for k in names:
        pinned[k][:] = bodies[k]
    up = [pinned[k] for k in names]
    down = {k: pinned[k] for k in ("qx", "qy", "qz", "vx", "vy", "vz")}
    ctx.upload_raw(up)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        ctx.sync()

    def reduce_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: K x computeOneIteration, per-step CUDA events, L2 evicted between steps
    for _ in range(args.warmup):
        ctx.step(DT, 0, 1)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.profile_enable(True)
    launches0 = ctx.launch_count
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    for i in range(args.steps):
        ctx.flush_l2()
        ctx.event_record(0)
        ctx.step(DT, 0, 1)
        ctx.event_record(1)
        dev_ms += ctx.event_elapsed_ms(0, 1)
    barrier()
    bracket_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    force_ms, force_launches = ctx.profile_get()
    ctx.profile_enable(False)
    dev_ms = reduce_max(dev_ms)
    force_ms = reduce_max(force_ms)
    interactions_per_step = float(n) * float(n)
    value = interactions_per_step * args.steps / (dev_ms * 1e-3) / 1e9

    # ---- end-to-end arm: pinned host state in, host state out, every step, through the public C ABI
    # same number of steps, but bounded to ~30 s of wall time for long-step workloads (at least 3 steps)
    e2e_steps = int(max(min(args.steps, 3), min(args.steps, 30e3 / max(dev_ms / args.steps, 1e-3))))
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ctx.upload_raw(up)               # H2D: qx qy qz m vx vy vz
        ctx.step(DT, 0, 1)
        ctx.download_state(down)         # D2H: qx qy qz vx vy vz (joins the device)
    barrier()
    e2e_s = reduce_max(time.perf_counter() - t0)
    e2e_value = interactions_per_step * e2e_steps / e2e_s / 1e9

    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the force pass): FP32 pipe, not HBM, not tensor
    max_mhz, peak_src = sm_max_mhz()
    peak_tflops = N_SMS * FP32_LANES * 2 * max_mhz * 1e6 / 1e12
    per_gpu_int_per_launch = interactions_per_step / world * args.steps / max(force_launches, 1)
    avg_launch_ms = force_ms / max(force_launches, 1)
    kernel_int_per_s = per_gpu_int_per_launch / (avg_launch_ms * 1e-3)
    achieved_tflops = 2 * PIPE_SLOTS_PER_INTERACTION * kernel_int_per_s / 1e12
    meas_mhz = clocks.get("sm_mhz") or max_mhz
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        t = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic.json")))
        if t["workload_bodies"] == n and world == 1:
            traffic = t["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {
        "bound": "fp32_pipe", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": achieved_tflops / peak_tflops,
        "traffic": traffic,
        "note": "compute-bound kernel: neither 'hbm' nor 'tensor' applies (SURVEY §8d). achieved = 12 FP32-pipe slots per "
                "interaction counted as FMA (2 flop) x interactions per launch / CUDA-event launch duration; peak = "
                f"148 SMs x 128 lanes x 2 x {max_mhz:.0f} MHz ({peak_src}; MEASURED_PEAKS.json has no FP32 entry), so "
                "frac == 12*int/s / (148*128*f). The pipe peak is reachable: a pure FFMA2 loop measures 1.987 of 2.0 "
                "warp-instr/clk/SM and FFMA 3.9 of 4.0 on this part (profiles/r01_microbench_pipes.txt).",
        "kernel": ctx.kernel_name, "avg_launch_ms": avg_launch_ms, "launches": force_launches,
        "kernel_share_of_step": force_ms / dev_ms,
        "kernel_gint_per_s_per_gpu": kernel_int_per_s / 1e9,
        "interactions_per_clk_per_sm": kernel_int_per_s / (meas_mhz * 1e6) / N_SMS,
        "frac_at_measured_clock": 2 * PIPE_SLOTS_PER_INTERACTION * kernel_int_per_s / (N_SMS * FP32_LANES * 2 * meas_mhz * 1e6),
        "murb_gflops_20flop_2p30": 20.0 * value * 1e9 / 2 ** 30,
        "hbm_algorithmic_bytes_per_launch": 28.0 * n / world, "hbm_gbs_algorithmic": 28.0 * n / world / (avg_launch_ms * 1e-3) / 1e9,
    }
    cpu = cpu_reference_rate(args.scheme, 30000, 12.0) if world == 1 and not args.no_cpu else None
    # The --gpus N>1 lines run the strong-scaling workload (n = 4,194,304).  So that a 1..8 series has a same-workload
    # 1-GPU point, the single-GPU line also carries the rate on that workload (2 timed steps, ~35 s).
    scaling_base = None
    if world == 1 and not args.no_scaling_base and n != STRONG_SCALING_BODIES:
        ctx.close()
        nb = STRONG_SCALING_BODIES
        big = b200nb.init_bodies(args.scheme, nb)
        with b200nb.Context(nb, b200nb.G_F32, SOFT, 1) as c2:
            c2.upload(*[big[k] for k in names])
            c2.step(DT, 0, 1)
            ms = 0.0
            for _ in range(2):
                c2.event_record(0)
                c2.step(DT, 0, 1)
                c2.event_record(1)
                ms += c2.event_elapsed_ms(0, 1)
        scaling_base = {"bodies": nb, "n_gpus": 1, "value": float(nb) ** 2 * 2 / (ms * 1e-3) / 1e9, "unit": "G-int/s",
                        "ms_per_step": ms / 2, "steps": 2, "warmup": 1,
                        "note": "same workload as the --gpus N>1 lines (BASELINE configs[4]); efficiency(N) = value(N) / (N * this)"}
    out = {
        "metric": METRIC, "value": value, "unit": "G-int/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        # the 1..8 GPU series is the strong-scaling one of BASELINE configs[4]; its same-workload 1-GPU point is
        # `strong_scaling_base` on the --gpus 1 line (whose own `value` is configs[1], n = 200 000)
        "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (reference Bodies generator restated, srand(0))",
        "config": {"workload": f"murb -n {n} -i {args.steps} --nv --im gpu+b200 --gf  ({args.scheme}, soft {SOFT:g}, dt {DT:g})"
                   + ("" if world == 1 else f"; targets sharded over {world} GPUs, ncclAllGather of positions per step; the 1-GPU "
                      "rate on this workload is the `strong_scaling_base` of the --gpus 1 line"),
                   "bodies": n, "scheme": args.scheme, "integrator": "murb-explicit",
                   "l2": "256 MiB memset between timed steps (outside the per-step event pair); inputs are 16 B/body and L2-resident by design",
                   "timing": "sum over steps of CUDA-event pairs on the library's compute stream, max over ranks"},
        "gpu_launches": int(launches), "bracket_wall_ms": bracket_ms,
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "power_w_max": clocks["power_w_max"], "samples": clocks["samples"]},
        "e2e": {"value": e2e_value, "unit": "G-int/s", "h2d_bytes_per_step": 7 * 4 * n * world, "d2h_bytes_per_step": 6 * 4 * n * world,
                "ms_per_step": e2e_s * 1e3 / e2e_steps, "steps": e2e_steps,
                "path": "b200nb_upload (pinned host SoA) + b200nb_step + b200nb_download_state, host wall clock"},
        "roofline": roofline,
    }
    if cpu is not None:
        out["cpu_baseline"] = cpu
    if scaling_base is not None:
        out["strong_scaling_base"] = scaling_base
    out_guard.emit(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bodies", type=int, default=None)
    ap.add_argument("--scheme", default="galaxy", choices=["galaxy", "random"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-scaling-base", action="store_true", help="skip the 1-GPU run of the strong-scaling workload")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(args.gpus, 1) and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                         "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
    if args.bodies is None:
        args.bodies = 200000 if args.gpus <= 1 else STRONG_SCALING_BODIES
    if args.steps is None:
        args.steps = 200 if args.gpus <= 1 else 5
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args, rank)
    else:
        run_b200_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
