"""The C-ABI library loads and exports every symbol include/b200nb.h declares; without a GPU the compute entry
points fail loudly (no CPU fallback)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import REPO, has_gpu


def test_header_symbols_exported(b200):
    names = b200.header_functions()
    assert len(names) >= 25
    L = ctypes.CDLL(b200.lib_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/b200nb.h but not exported: {missing}"


def test_header_cites_reference_lines():
    text = open(os.path.join(REPO, "include", "b200nb.h")).read()
    for needle in ("SimulationNBodyInterface.hpp", "CUDABodies.cu", "SimulationNBodyNaive.cpp", "main.cpp"):
        assert needle in text


def test_no_oracle_in_product():
    """The product path must not import, link or call the oracle."""
    pkg = os.path.join(REPO, "nbody-eurohpc_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h", ".py")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert not any(w in src for w in ("nbody_oracle", "liboracle", "libmurbref", "pyoracle", "oracle/")), f
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "b200nb", "libb200nb.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "murbref" not in out


@pytest.mark.skipif(has_gpu(), reason="CPU-only behaviour")
def test_fails_loudly_without_gpu(b200):
    with pytest.raises(b200.B200Error) as e:
        b200.Context(1024)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(b200.B200Error):
        b200.SimulationNBodyB200(1024)


def test_bad_arguments(b200):
    L = b200.lib()
    assert L.b200nb_step(None, 1.0, 0, 1) == 1
    assert L.b200nb_init_bodies(7, 10, 0, *[None] * 8) == 1
    assert L.b200nb_slice_length(200000, 1) == 200192  # 256-body granularity
    assert L.b200nb_slice_length(200000, 8) == 25088
    assert L.b200nb_slice_length(4194304, 8) == 524288
    assert L.b200nb_slice_length(5, 0) == 0
