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


def _write_tab(path, n, seed=3):
    rng = np.random.default_rng(seed)
    rows = rng.normal(size=(n, 7)) * np.array([1e-5, 3.0, 3.0, 0.5, 0.7, 0.7, 0.2]) + np.array([2e-5, 0, 0, 0, 0, 0, 0])
    with open(path, "w") as f:
        for k, r in enumerate(rows):
            f.write(" ".join(f"{v:.7e}" for v in r) + "\n")
            if k % 1000 == 7:
                f.write("\n")  # empty lines are skipped by the reference


def test_tab_loader_components(b200, tmp_path):
    """initMilkyWayAndromeda (Bodies.cpp:82-153): component-dependent rescaling and the 1e5 radius."""
    n = 83000  # crosses every component boundary: 16384 / 32768 / 40960 / 49152 / 65536 / 81920
    path = str(tmp_path / "milkyway_andromeda.tab")
    _write_tab(path, n)
    d = b200.load_tab(path)
    raw = np.array([[float(x) for x in l.split()] for l in open(path) if l.strip()], dtype=np.float32)
    assert len(d["m"]) == n and np.all(d["r"] == np.float32(1e5))
    for i, mw in ((0, True), (16383, True), (16384, False), (32768, True), (40959, True), (40960, False), (49152, True),
                  (65535, True), (65536, False), (82999, False)):
        sm, sq, sv = (4.5e10, 4.0, 220) if mw else (9.4e10, 6.0, 260)
        assert d["m"][i] == np.float32(float(raw[i, 0]) * sm)
        assert d["qx"][i] == np.float32(float(raw[i, 1]) * sq)
        assert d["vz"][i] == np.float32(raw[i, 6] * np.float32(sv))


@pytest.mark.skipif(not os.path.exists(os.path.join(REPO, "oracle", "_ref", "libmurbref.so")), reason="reference not built here")
def test_tab_loader_matches_live_reference(b200, tmp_path, monkeypatch):
    n = 50000
    _write_tab(str(tmp_path / "milkyway_andromeda.tab"), n, seed=11)
    monkeypatch.chdir(tmp_path)  # the reference opens the file relative to the CWD (Bodies.cpp:85)
    ref = ctypes.CDLL(os.path.join(REPO, "oracle", "_ref", "libmurbref.so"))
    ref.ref_init_bodies.argtypes = [ctypes.c_uint64, ctypes.c_char_p] + [FP] * 8
    r = {k: np.empty(n, np.float32) for k in KEYS}
    ref.ref_init_bodies(n, b"milkyway", *[r[k].ctypes.data_as(FP) for k in KEYS])
    d = b200.load_tab(str(tmp_path / "milkyway_andromeda.tab"))
    for k in KEYS:
        assert np.array_equal(r[k].view(np.uint32), d[k].view(np.uint32)), k
