"""Initial conditions: the product's host generator (csrc/host_ic.cpp) and the oracle's (oracle/nbody_oracle.c) are both
bit-identical to the reference's Bodies<float> (Bodies.cpp:158-257), pinned by checksums generated from the reference
(tests/golden/make_golden.py) and, when oracle/_ref is present, by the live reference."""
import ctypes
import hashlib
import os

import numpy as np
import pytest

from conftest import FP, REPO

KEYS = ("qx", "qy", "qz", "vx", "vy", "vz", "m", "r")
CASES = [(s, n) for s in ("galaxy", "random") for n in (1, 127, 2048, 2049, 4000, 30000, 200000)]


@pytest.mark.parametrize("scheme,n", CASES)
def test_product_ic_matches_reference_checksums(b200, ic_checksums, scheme, n):
    d = b200.init_bodies(scheme, n)
    want = ic_checksums[f"{scheme}:{n}"]["sha256"]
    for k in KEYS:
        assert hashlib.sha256(d[k].tobytes()).hexdigest() == want[k], (scheme, n, k)


@pytest.mark.parametrize("scheme,n", CASES)
def test_oracle_ic_matches_reference_checksums(oracle, ic_checksums, scheme, n):
    d = oracle.init_bodies(scheme, n)
    want = ic_checksums[f"{scheme}:{n}"]["sha256"]
    for k in KEYS:
        assert hashlib.sha256(d[k].tobytes()).hexdigest() == want[k], (scheme, n, k)


def test_galaxy_shape(b200):
    d = b200.init_bodies("galaxy", 1000)
    assert d["m"][0] == np.float32(2.0e24) and d["qx"][0] == 0 and d["r"][0] == 0  # Bodies.cpp:171-180
    rad = np.sqrt(d["qx"][1:].astype(np.float64) ** 2 + d["qy"][1:] ** 2 + d["qz"][1:] ** 2)
    assert rad.min() >= 1e8 * 0.999 and rad.max() <= 2e8 * 1.001
    assert np.all(d["vz"] == 0)


@pytest.mark.skipif(not os.path.exists(os.path.join(REPO, "oracle", "_ref", "libmurbref.so")), reason="reference not built here")
@pytest.mark.parametrize("scheme,n", [("galaxy", 5000), ("random", 5001)])
def test_live_reference(b200, oracle, scheme, n):
    ref = ctypes.CDLL(os.path.join(REPO, "oracle", "_ref", "libmurbref.so"))
    ref.ref_init_bodies.argtypes = [ctypes.c_uint64, ctypes.c_char_p] + [FP] * 8
    r = {k: np.empty(n, np.float32) for k in KEYS}
    ref.ref_init_bodies(n, scheme.encode(), *[r[k].ctypes.data_as(FP) for k in KEYS])
    p, o = b200.init_bodies(scheme, n), oracle.init_bodies(scheme, n)
    for k in KEYS:
        assert np.array_equal(r[k].view(np.uint32), p[k].view(np.uint32)), k
        assert np.array_equal(r[k].view(np.uint32), o[k].view(np.uint32)), k
