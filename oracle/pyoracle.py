"""ctypes view of oracle/liboracle.so (oracle/nbody_oracle.h) plus the comparison helpers the tests use.

TEST INFRASTRUCTURE ONLY — imported by tests/conftest.py, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product (nbody-eurohpc_b200/)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FP = ctypes.POINTER(ctypes.c_float)
DP = ctypes.POINTER(ctypes.c_double)
G_F32 = np.float32(6.67384e-11)
SOFT, DT = 2e8, 3600.0
SCHEME_ID = {"galaxy": 0, "random": 1}


def _fp(a):
    return a.ctypes.data_as(FP)


def lib_path(build=True):
    path = os.path.join(HERE, "liboracle.so")
    if build and not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.dirname(HERE), "oracle"], stdout=subprocess.DEVNULL)
    return path


class Oracle:
    """ctypes view of oracle/liboracle.so (oracle/nbody_oracle.h).  TEST INFRASTRUCTURE."""

    def __init__(self, path):
        L = ctypes.CDLL(path)
        u64, f, i = ctypes.c_uint64, ctypes.c_float, ctypes.c_int
        L.oracle_init_bodies.argtypes = [i, u64, ctypes.c_uint] + [FP] * 8
        L.oracle_init_bodies.restype = i
        L.oracle_accel_naive_f32.argtypes = [u64] + [FP] * 4 + [f, f] + [FP] * 3
        L.oracle_accel_naive_f32.restype = None
        L.oracle_accel_f64.argtypes = [u64] + [FP] * 4 + [f, f, ctypes.POINTER(u64), u64] + [DP] * 3
        L.oracle_accel_f64.restype = None
        L.oracle_integrate_murb.argtypes = [u64] + [FP] * 9 + [f]
        L.oracle_integrate_murb.restype = None
        L.oracle_run_naive.argtypes = [u64] + [FP] * 7 + [f, f, f, i] + [FP] * 3
        L.oracle_run_naive.restype = None
        L.oracle_run_f64force.argtypes = [u64] + [FP] * 7 + [f, f, f, i, i]
        L.oracle_run_f64force.restype = None
        L.oracle_energy_f64.argtypes = [u64] + [FP] * 7 + [f, f]
        L.oracle_energy_f64.restype = ctypes.c_double
        L.oracle_metrics_f64.argtypes = [u64] + [FP] * 7 + [f, f, ctypes.POINTER(ctypes.c_double)]
        L.oracle_metrics_f64.restype = None
        self.L = L

    def init_bodies(self, scheme, n, seed=0):
        d = {k: np.empty(n, np.float32) for k in ("qx", "qy", "qz", "vx", "vy", "vz", "m", "r")}
        rc = self.L.oracle_init_bodies(SCHEME_ID[scheme], n, seed, *[_fp(d[k]) for k in ("qx", "qy", "qz", "vx", "vy", "vz", "m", "r")])
        assert rc == 0
        return d

    def accel_naive(self, d, G=G_F32, soft=SOFT):
        n = len(d["qx"])
        a = [np.empty(n, np.float32) for _ in range(3)]
        self.L.oracle_accel_naive_f32(n, _fp(d["qx"]), _fp(d["qy"]), _fp(d["qz"]), _fp(d["m"]), G, soft, *[_fp(x) for x in a])
        return a

    def accel_f64(self, d, idx=None, G=G_F32, soft=SOFT):
        n = len(d["qx"])
        if idx is None:
            n_idx, ip = n, None
        else:
            idx = np.ascontiguousarray(idx, np.uint64)
            n_idx, ip = len(idx), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
        a = [np.empty(n_idx, np.float64) for _ in range(3)]
        self.L.oracle_accel_f64(n, _fp(d["qx"]), _fp(d["qy"]), _fp(d["qz"]), _fp(d["m"]), G, soft, ip, n_idx,
                                *[x.ctypes.data_as(DP) for x in a])
        return a

    def integrate_murb(self, d, ax, ay, az, dt):
        n = len(d["qx"])
        self.L.oracle_integrate_murb(n, *[_fp(d[k]) for k in ("qx", "qy", "qz", "vx", "vy", "vz")], _fp(ax), _fp(ay), _fp(az), dt)

    def run_naive(self, d, n_iter, G=G_F32, soft=SOFT, dt=DT):
        n = len(d["qx"])
        a = [np.empty(n, np.float32) for _ in range(3)]
        self.L.oracle_run_naive(n, *[_fp(d[k]) for k in ("qx", "qy", "qz", "vx", "vy", "vz", "m")], G, soft, dt, n_iter, *[_fp(x) for x in a])
        return a

    def run_f64force(self, d, n_iter, integrator, G=G_F32, soft=SOFT, dt=DT):
        n = len(d["qx"])
        self.L.oracle_run_f64force(n, *[_fp(d[k]) for k in ("qx", "qy", "qz", "vx", "vy", "vz", "m")], G, soft, dt, integrator, n_iter)

    def energy(self, d, G=G_F32, soft=SOFT):
        n = len(d["qx"])
        return self.L.oracle_energy_f64(n, *[_fp(d[k]) for k in ("qx", "qy", "qz", "vx", "vy", "vz", "m")], G, soft)

    METRIC_NAMES = ("energy", "ang_x", "ang_y", "ang_z", "mass", "com_x", "com_y", "com_z", "density_x", "density_y", "density_z")

    def metrics(self, d, G=G_F32, soft=SOFT):
        n = len(d["qx"])
        out = (ctypes.c_double * 11)()
        self.L.oracle_metrics_f64(n, *[_fp(d[k]) for k in ("qx", "qy", "qz", "vx", "vy", "vz", "m")], G, soft, out)
        return dict(zip(self.METRIC_NAMES, list(out)))


def max_rel_err(ref3, got3):
    """max over bodies of |a - a_ref| / |a_ref| (vector norms; per-component ratios are meaningless)."""
    r = np.stack([np.asarray(x, np.float64) for x in ref3])
    g = np.stack([np.asarray(x, np.float64) for x in got3])
    return float(np.max(np.linalg.norm(g - r, axis=0) / np.linalg.norm(r, axis=0)))


def within_rel(a, b, eps):
    """Catch2 WithinRel: |a-b| <= eps * max(|a|,|b|)  (lib/Catch2/include/catch.hpp:11504-11508)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) <= eps * np.maximum(np.abs(a), np.abs(b))



def load():
    return Oracle(lib_path())
