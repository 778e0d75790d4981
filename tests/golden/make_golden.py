#!/usr/bin/env python3
"""Generate tests/golden/* from the UNMODIFIED reference, compiled from /root/reference by oracle/build_ref.sh into
oracle/_ref/libmurbref.so (as-shipped flags: -O3 -ffast-math, no -march; see oracle/ref_wrap.cpp for the driver).

Run in the build container (the GPU box has no /root/reference):   python tests/golden/make_golden.py
Outputs (committed):
  ic_checksums.json      sha256 + first values of Bodies<float>(n, scheme, 0) arrays          (Bodies.cpp:158-257)
  murb_ref_golden.npz    cpu+naive trajectories of the four murb-test sections                (test_SimulationNBody.cpp:76-81)
                         cpu+naive accelerations a(x0)                                        (SimulationNBodyNaive.cpp:34-53)
                         Bodies::updatePositionsAndVelocities with synthetic accelerations    (test_CUDABodies.cpp:42-75)
"""
import ctypes
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
FP = ctypes.POINTER(ctypes.c_float)
ref = ctypes.CDLL(os.path.join(REPO, "oracle", "_ref", "libmurbref.so"))
ref.ref_init_bodies.argtypes = [ctypes.c_uint64, ctypes.c_char_p] + [FP] * 8
ref.ref_run.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_float, ctypes.c_float, ctypes.c_int] + [FP] * 9
ref.ref_run.restype = ctypes.c_double
ref.ref_accel.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_float, ctypes.c_int] + [FP] * 3
ref.ref_accel.restype = ctypes.c_double
ref.ref_integrate.argtypes = [ctypes.c_uint64, ctypes.c_char_p] + [FP] * 3 + [ctypes.c_float, ctypes.c_int] + [FP] * 6
KEYS = ("qx", "qy", "qz", "vx", "vy", "vz", "m", "r")


def p(a):
    return a.ctypes.data_as(FP)


def init(n, scheme):
    d = {k: np.empty(n, np.float32) for k in KEYS}
    pad = ref.ref_init_bodies(n, scheme.encode(), *[p(d[k]) for k in KEYS])
    return d, pad


def run(tag, n, scheme, soft, dt, iters):
    o = [np.empty(n, np.float32) for _ in range(9)]
    ms = ref.ref_run(tag.encode(), n, scheme.encode(), soft, dt, iters, *[p(a) for a in o])
    assert ms >= 0
    return o


def main():
    checks = {}
    for scheme in ("galaxy", "random"):
        for n in (1, 127, 2048, 2049, 4000, 30000, 200000):
            d, pad = init(n, scheme)
            checks[f"{scheme}:{n}"] = {
                "padding_mipp_sse": pad,
                "sha256": {k: hashlib.sha256(d[k].tobytes()).hexdigest() for k in KEYS},
                "head": {k: [float(x) for x in d[k][:4]] for k in KEYS},
            }
    json.dump(checks, open(os.path.join(HERE, "ic_checksums.json"), "w"), indent=1, sort_keys=True)

    out = {}
    soft, dt = 2e8, 3600.0
    for n, iters, scheme in ((2048, 1, "random"), (2049, 3, "random"), (2048, 4, "galaxy"), (2049, 3, "galaxy")):
        for it in range(1, iters + 1):
            o = run("cpu+naive", n, scheme, soft, dt, it)
            for nm, a in zip(("qx", "qy", "qz", "vx", "vy", "vz", "ax", "ay", "az"), o):
                if nm[0] == "q" or it == iters:
                    out[f"traj/{scheme}/{n}/it{it}/{nm}"] = a
    for n, scheme in ((2048, "galaxy"), (2049, "random"), (8191, "galaxy")):
        a = [np.empty(n, np.float32) for _ in range(3)]
        ref.ref_accel(b"cpu+naive", n, scheme.encode(), soft, 1, *[p(x) for x in a])
        for nm, x in zip(("ax", "ay", "az"), a):
            out[f"accel0/{scheme}/{n}/{nm}"] = x
    for scheme in ("random", "galaxy"):
        n = 4000
        i = np.arange(n, dtype=np.float32)
        acc = [i + 1, np.full(n, 3.0, np.float32), np.float32(n) - i]
        acc = [np.ascontiguousarray(a, np.float32) for a in acc]
        o = [np.empty(n, np.float32) for _ in range(6)]
        ref.ref_integrate(n, scheme.encode(), *[p(a) for a in acc], 0.01, 4, *[p(a) for a in o])
        for nm, a in zip(("qx", "qy", "qz", "vx", "vy", "vz"), o):
            out[f"integrate/{scheme}/{n}/{nm}"] = a
    np.savez_compressed(os.path.join(HERE, "murb_ref_golden.npz"), **out)
    print("wrote", len(checks), "IC checksums and", len(out), "golden arrays")


if __name__ == "__main__":
    main()
