"""b200nb — ctypes binding of libb200nb (include/b200nb.h) and a Python mirror of the MUrB plugin interface.

The product is the CUDA library; this module is plumbing so that tests and bench.py can drive the C ABI on a box that
has no copy of the reference.  The C++ glue that the `murb` CLI links is nbody-eurohpc_b200/glue/.

`SimulationNBodyB200` mirrors the reference's SimulationNBodyInterface<float>
(src/common/core/SimulationNBodyInterface.hpp:15-88): same method names, argument meaning and error behaviour
(`computeOneIteration`, `setDt`, `getDt`, `getBodies().getDataSoA()`, `getFlopsPerIte`, `getAllocatedBytes`).
There is no CPU fallback: if the shared library is missing, or no CUDA device is visible, this module raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_uint, c_uint64, c_void_p

import numpy as np

__all__ = [
    "B200Error", "Context", "SimulationNBodyB200", "B200Bodies", "init_bodies", "lib", "lib_path", "header_functions",
    "G_F32", "INTEGRATOR_MURB", "INTEGRATOR_LEAPFROG", "slice_length", "load_tab",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
G_F32 = np.float32(6.67384e-11)  # SimulationNBodyInterface.hpp:18 (a float literal)
INTEGRATOR_MURB, INTEGRATOR_LEAPFROG = 0, 1
OK, EINVAL, ECUDA, ENCCL, ESTATE = 0, 1, 2, 3, 4
_FP = POINTER(c_float)


class B200Error(RuntimeError):
    def __init__(self, code: int, what: str, msg: str):
        super().__init__(f"{what} failed ({code}): {msg}")
        self.code = code


def lib_path() -> str:
    return os.path.join(_HERE, "libb200nb.so")


def header_functions() -> list[str]:
    """Every function declared in include/b200nb.h (used by the ABI export test)."""
    text = open(os.path.join(_REPO, "include", "b200nb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200nb_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> ctypes.CDLL:
    """Load libb200nb.so (built in-tree by `make lib` / __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise B200Error(-1, "load", f"{path} not found: build it with `make lib` (there is no CPU fallback)")
    L = ctypes.CDLL(path)
    ctx = c_void_p
    sig = {
        "b200nb_create": (c_int, [POINTER(ctx), c_uint64, c_int, c_float, c_float]),
        "b200nb_create_sharded": (c_int, [POINTER(ctx), c_uint64, c_int, POINTER(c_int), c_float, c_float]),
        "b200nb_create_rank": (c_int, [POINTER(ctx), c_uint64, c_float, c_float, c_int, c_int, c_int, c_void_p]),
        "b200nb_comm_unique_id": (c_int, [c_void_p]),
        "b200nb_destroy": (None, [ctx]),
        "b200nb_last_error": (c_char_p, [ctx]),
        "b200nb_upload": (c_int, [ctx] + [_FP] * 7),
        "b200nb_download_state": (c_int, [ctx] + [_FP] * 6),
        "b200nb_download_slice": (c_int, [ctx] + [_FP] * 6),
        "b200nb_slice_bounds": (c_int, [ctx, c_int, POINTER(c_uint64), POINTER(c_uint64)]),
        "b200nb_download_accel": (c_int, [ctx] + [_FP] * 3),
        "b200nb_step": (c_int, [ctx, c_float, c_int, c_int]),
        "b200nb_accel": (c_int, [ctx]),
        "b200nb_integrate_host_accel": (c_int, [ctx, _FP, _FP, _FP, c_float]),
        "b200nb_energy": (c_int, [ctx, POINTER(c_double)]),
        "b200nb_metrics": (c_int, [ctx, POINTER(c_double)]),
        "b200nb_sync": (c_int, [ctx]),
        "b200nb_n_bodies": (c_uint64, [ctx]),
        "b200nb_n_local_gpus": (c_int, [ctx]),
        "b200nb_allocated_bytes": (c_uint64, [ctx]),
        "b200nb_launch_count": (c_uint64, [ctx]),
        "b200nb_kernel_name": (c_char_p, [ctx]),
        "b200nb_exchange_name": (c_char_p, [ctx]),
        "b200nb_event_record": (c_int, [ctx, c_int]),
        "b200nb_event_elapsed_ms": (c_int, [ctx, c_int, c_int, POINTER(c_float)]),
        "b200nb_profile_enable": (c_int, [ctx, c_int]),
        "b200nb_profile_get": (c_int, [ctx, POINTER(c_double), POINTER(c_uint64)]),
        "b200nb_flush_l2": (c_int, [ctx]),
        "b200nb_host_alloc": (c_int, [POINTER(c_void_p), c_uint64]),
        "b200nb_host_free": (c_int, [c_void_p]),
        "b200nb_init_bodies": (c_int, [c_int, c_uint64, c_uint] + [_FP] * 8),
        "b200nb_slice_length": (c_uint64, [c_uint64, c_int]),
        "b200nb_tab_count": (c_int, [c_char_p, POINTER(c_uint64)]),
        "b200nb_tab_load": (c_int, [c_char_p, c_uint64] + [_FP] * 8),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray | None):
    return a.ctypes.data_as(_FP) if a is not None else None


_SCHEMES = {"galaxy": 0, "random": 1}


def init_bodies(scheme: str, n: int, seed: int = 0) -> dict[str, np.ndarray]:
    """Bodies<float>(n, scheme, seed) — host only (Bodies.cpp:158-257 restated in csrc/host_ic.cpp)."""
    if scheme not in _SCHEMES:
        raise ValueError(f"scheme must be one of {list(_SCHEMES)}")
    out = {k: np.empty(n, dtype=np.float32) for k in ("qx", "qy", "qz", "vx", "vy", "vz", "m", "r")}
    rc = lib().b200nb_init_bodies(_SCHEMES[scheme], n, seed, *[_p(out[k]) for k in ("qx", "qy", "qz", "vx", "vy", "vz", "m", "r")])
    if rc != OK:
        raise B200Error(rc, "b200nb_init_bodies", "bad arguments")
    return out


def load_tab(path: str) -> dict[str, np.ndarray]:
    """Bodies<float>::initMilkyWayAndromeda (Bodies.cpp:82-153): the reference's `.tab` initial-condition file."""
    n = c_uint64()
    if lib().b200nb_tab_count(path.encode(), byref(n)) != OK:
        raise B200Error(EINVAL, "b200nb_tab_count", f"cannot read {path}")
    out = {k: np.empty(n.value, dtype=np.float32) for k in ("qx", "qy", "qz", "vx", "vy", "vz", "m", "r")}
    rc = lib().b200nb_tab_load(path.encode(), n.value, *[_p(out[k]) for k in ("qx", "qy", "qz", "vx", "vy", "vz", "m", "r")])
    if rc != OK:
        raise B200Error(rc, "b200nb_tab_load", f"parse error in {path}")
    return out


def slice_length(n: int, n_ranks: int) -> int:
    """Targets per rank (padded); rank r owns global bodies [r*L, min((r+1)*L, n))."""
    return int(lib().b200nb_slice_length(n, n_ranks))


class PinnedArrays:
    """A block of page-locked float32 arrays (cudaHostAlloc) for the end-to-end path."""

    def __init__(self, names: list[str], n: int):
        self._ptr = c_void_p()
        nbytes = 4 * n * len(names)
        rc = lib().b200nb_host_alloc(byref(self._ptr), nbytes)
        if rc != OK:
            raise B200Error(rc, "b200nb_host_alloc", lib().b200nb_last_error(None).decode())
        buf = (ctypes.c_float * (n * len(names))).from_address(self._ptr.value)
        flat = np.frombuffer(buf, dtype=np.float32)
        self.arrays = {nm: flat[i * n:(i + 1) * n] for i, nm in enumerate(names)}

    def __getitem__(self, k):
        return self.arrays[k]

    def close(self):
        if self._ptr:
            self.arrays = {}
            lib().b200nb_host_free(self._ptr)
            self._ptr = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """RAII wrapper of b200nb_ctx.  One process driving `n_gpus` devices, or one rank of a torchrun job."""

    def __init__(self, n: int, G: float = G_F32, soft: float = 2e8, n_gpus: int = 1, *, rank: int | None = None,
                 n_ranks: int = 1, device: int = 0, nccl_id: bytes | None = None, devices: list[int] | None = None):
        self._L = lib()
        self._ctx = c_void_p()
        self.n = int(n)
        if devices is not None:  # explicit placement, one shard per entry; a device may repeat (virtual shards)
            arr = (c_int * len(devices))(*devices)
            rc = self._L.b200nb_create_sharded(byref(self._ctx), n, len(devices), arr, float(G), float(soft))
            what = "b200nb_create_sharded"
        elif rank is None:
            rc = self._L.b200nb_create(byref(self._ctx), n, n_gpus, float(G), float(soft))
            what = "b200nb_create"
        else:
            idbuf = ctypes.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
            rc = self._L.b200nb_create_rank(byref(self._ctx), n, float(G), float(soft), rank, n_ranks, device,
                                            ctypes.cast(idbuf, c_void_p) if idbuf is not None else None)
            what = "b200nb_create_rank"
        if rc != OK:
            raise B200Error(rc, what, self._L.b200nb_last_error(None).decode())

    @staticmethod
    def unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        rc = lib().b200nb_comm_unique_id(ctypes.cast(buf, c_void_p))
        if rc != OK:
            raise B200Error(rc, "b200nb_comm_unique_id", lib().b200nb_last_error(None).decode())
        return buf.raw

    def _check(self, rc: int, what: str):
        if rc != OK:
            raise B200Error(rc, what, self._L.b200nb_last_error(self._ctx).decode())

    def close(self):
        if self._ctx:
            self._L.b200nb_destroy(self._ctx)
            self._ctx = c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state
    def upload(self, qx, qy, qz, m, vx, vy, vz):
        arrs = [_f32(a) for a in (qx, qy, qz, m, vx, vy, vz)]
        for a in arrs:
            if a.shape != (self.n,):
                raise ValueError(f"expected arrays of {self.n} floats, got {a.shape}")
        self._check(self._L.b200nb_upload(self._ctx, *[_p(a) for a in arrs]), "b200nb_upload")

    def upload_raw(self, arrs):
        """float32 C-contiguous arrays (e.g. pinned) passed without conversion: the C side reads n floats from each."""
        if len(arrs) != 7:
            raise ValueError("expected 7 arrays: qx qy qz m vx vy vz")
        for a in arrs:
            if a.dtype != np.float32 or a.shape != (self.n,) or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"expected C-contiguous float32 arrays of {self.n} elements, got {a.dtype} {a.shape}")
        self._check(self._L.b200nb_upload(self._ctx, *[_p(a) for a in arrs]), "b200nb_upload")

    _STATE = ("qx", "qy", "qz", "vx", "vy", "vz")

    def _state_out(self, out):
        if out is None:
            return {k: np.empty(self.n, dtype=np.float32) for k in self._STATE}
        for k in self._STATE:
            a = out.get(k)
            if a is not None and (a.dtype != np.float32 or a.shape != (self.n,) or not a.flags["C_CONTIGUOUS"]):
                raise ValueError(f"{k}: expected a C-contiguous float32 array of {self.n} elements, got {a.dtype} {a.shape}")
        return out

    def download_state(self, out: dict[str, np.ndarray] | None = None) -> dict[str, np.ndarray]:
        out = self._state_out(out)
        self._check(self._L.b200nb_download_state(self._ctx, *[_p(out.get(k)) for k in self._STATE]), "b200nb_download_state")
        return out

    def download_slice(self, out: dict[str, np.ndarray] | None = None) -> dict[str, np.ndarray]:
        """Only the local shards' own bodies are written (global indexing); see slice_bounds()."""
        out = self._state_out(out)
        self._check(self._L.b200nb_download_slice(self._ctx, *[_p(out.get(k)) for k in self._STATE]), "b200nb_download_slice")
        return out

    def slice_bounds(self, local_shard: int = 0) -> tuple[int, int]:
        first, count = c_uint64(), c_uint64()
        self._check(self._L.b200nb_slice_bounds(self._ctx, local_shard, byref(first), byref(count)), "b200nb_slice_bounds")
        return first.value, count.value

    def download_accel(self) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        a = [np.empty(self.n, dtype=np.float32) for _ in range(3)]
        self._check(self._L.b200nb_download_accel(self._ctx, *[_p(x) for x in a]), "b200nb_download_accel")
        return a[0], a[1], a[2]

    # ---- hot path
    def step(self, dt: float, integrator: int = INTEGRATOR_MURB, n_steps: int = 1):
        self._check(self._L.b200nb_step(self._ctx, float(dt), integrator, n_steps), "b200nb_step")

    def accel(self):
        self._check(self._L.b200nb_accel(self._ctx), "b200nb_accel")

    def integrate_host_accel(self, ax, ay, az, dt: float):
        a = [_f32(x) for x in (ax, ay, az)]
        for x in a:
            if x.shape != (self.n,):
                raise ValueError(f"expected arrays of {self.n} floats, got {x.shape}")
        self._check(self._L.b200nb_integrate_host_accel(self._ctx, *[_p(x) for x in a], float(dt)), "b200nb_integrate_host_accel")

    def energy(self) -> float:
        e = c_double()
        self._check(self._L.b200nb_energy(self._ctx, byref(e)), "b200nb_energy")
        return e.value

    METRIC_NAMES = ("energy", "ang_x", "ang_y", "ang_z", "mass", "com_x", "com_y", "com_z", "density_x", "density_y", "density_z")

    def metrics(self) -> dict:
        """energy, angular momentum vector, total mass, centre of mass, density centre (include/b200nb.h: b200nb_metrics)."""
        out = (c_double * len(self.METRIC_NAMES))()
        self._check(self._L.b200nb_metrics(self._ctx, out), "b200nb_metrics")
        return dict(zip(self.METRIC_NAMES, list(out)))

    def sync(self):
        self._check(self._L.b200nb_sync(self._ctx), "b200nb_sync")

    # ---- introspection / measurement
    @property
    def n_local_gpus(self) -> int:
        return self._L.b200nb_n_local_gpus(self._ctx)

    @property
    def allocated_bytes(self) -> int:
        return self._L.b200nb_allocated_bytes(self._ctx)

    @property
    def launch_count(self) -> int:
        return self._L.b200nb_launch_count(self._ctx)

    @property
    def kernel_name(self) -> str:
        return self._L.b200nb_kernel_name(self._ctx).decode()

    @property
    def exchange_name(self) -> str:
        """'none' | 'p2p-push' | 'nccl-allgather' (include/b200nb.h: b200nb_exchange_name)."""
        return self._L.b200nb_exchange_name(self._ctx).decode()

    def event_record(self, slot: int):
        self._check(self._L.b200nb_event_record(self._ctx, slot), "b200nb_event_record")

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = c_float()
        self._check(self._L.b200nb_event_elapsed_ms(self._ctx, a, b, byref(ms)), "b200nb_event_elapsed_ms")
        return ms.value

    def flush_l2(self):
        self._check(self._L.b200nb_flush_l2(self._ctx), "b200nb_flush_l2")

    def profile_enable(self, on: bool = True):
        self._check(self._L.b200nb_profile_enable(self._ctx, int(on)), "b200nb_profile_enable")

    def profile_get(self) -> tuple[float, int]:
        ms, n = c_double(), c_uint64()
        self._check(self._L.b200nb_profile_get(self._ctx, byref(ms), byref(n)), "b200nb_profile_get")
        return ms.value, n.value


class B200Bodies:
    """Mirror of the glue's B200Bodies : Bodies<float> — host SoA mirror with lazy device -> host copy
    (pattern: CUDABodies::getDataSoA, src/common/core/CUDABodies.cu:63-93)."""

    def __init__(self, n: int, scheme: str = "galaxy", randInit: int = 0):
        self.n = int(n)
        self.scheme = scheme
        self.dataSoA = init_bodies(scheme, n, randInit)
        self._ctx: Context | None = None
        self._host_current = True

    def getN(self) -> int:
        return self.n

    def bind(self, G, soft, n_gpus=1, **rank_kw):
        if self._ctx is not None:
            self.getDataSoA()
            self._ctx.close()
        self._ctx = Context(self.n, G, soft, n_gpus, **rank_kw)
        d = self.dataSoA
        self._ctx.upload(d["qx"], d["qy"], d["qz"], d["m"], d["vx"], d["vy"], d["vz"])
        self._host_current = True

    def context(self) -> Context:
        if self._ctx is None:
            self.bind(G_F32, 1.0)
        return self._ctx

    def invalidateDataSoA(self):
        self._host_current = False

    def getDataSoA(self) -> dict[str, np.ndarray]:
        if not self._host_current and self._ctx is not None:
            self._ctx.download_state(self.dataSoA)
            self._host_current = True
        return self.dataSoA

    def updatePositionsAndVelocities(self, ax, ay, az, dt: float):
        self.context().integrate_host_accel(ax, ay, az, dt)
        self.invalidateDataSoA()


class SimulationNBodyB200:
    """Mirror of SimulationNBodyB200 : SimulationNBodyInterface<float> (glue/SimulationNBodyB200.hpp)."""

    def __init__(self, n: int, scheme: str = "galaxy", soft: float = 2e8, leapfrog: bool = False, n_gpus: int = 1,
                 randInit: int = 0, **rank_kw):
        self.G = G_F32
        self.soft = np.float32(soft)
        self.dt = np.float32(np.inf)  # SimulationNBodyInterface.cpp:12
        self.integrator = INTEGRATOR_LEAPFROG if leapfrog else INTEGRATOR_MURB
        self.bodies = B200Bodies(n, scheme, randInit)
        self.bodies.bind(self.G, self.soft, n_gpus, **rank_kw)
        self.flopsPerIte = np.float32(20.0) * np.float32(n) * np.float32(n)  # SimulationNBodyNaive.cpp:15

    def getBodies(self) -> B200Bodies:
        return self.bodies

    def setDt(self, dt: float):
        self.dt = np.float32(dt)

    def getDt(self) -> float:
        return float(self.dt)

    def getFlopsPerIte(self) -> float:
        return float(self.flopsPerIte)

    def getAllocatedBytes(self) -> float:
        return float(self.bodies.context().allocated_bytes + 8 * 2 * 4 * self.bodies.n)

    def computeOneIteration(self):
        ctx = self.bodies.context()
        ctx.step(self.dt, self.integrator, 1)
        if ctx.n_local_gpus > 1:
            ctx.sync()
        self.bodies.invalidateDataSoA()

    def computeAccelerationsOnly(self):
        self.bodies.context().accel()

    def getAccSoA(self):
        return self.bodies.context().download_accel()

    def computeEnergy(self) -> float:
        return self.bodies.context().energy()
