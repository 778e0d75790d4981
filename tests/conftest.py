"""Shared fixtures.  `-m "not gpu"` runs here on CPU (oracle vs golden vectors, host logic, ABI exports);
`-m gpu` are the parity tests proper, through the C ABI, on a B200."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "nbody-eurohpc_b200"))
sys.path.insert(0, REPO)

import importlib.util

_spec = importlib.util.spec_from_file_location("pyoracle", os.path.join(REPO, "oracle", "pyoracle.py"))
pyoracle = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(pyoracle)
Oracle, max_rel_err, within_rel = pyoracle.Oracle, pyoracle.max_rel_err, pyoracle.within_rel
FP, DP, G_F32, SOFT, DT, SCHEME_ID = pyoracle.FP, pyoracle.DP, pyoracle.G_F32, pyoracle.SOFT, pyoracle.DT, pyoracle.SCHEME_ID


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    return pyoracle.load()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(REPO, "tests", "golden", "murb_ref_golden.npz"))


@pytest.fixture(scope="session")
def ic_checksums():
    return json.load(open(os.path.join(REPO, "tests", "golden", "ic_checksums.json")))


@pytest.fixture(scope="session")
def b200():
    import b200nb
    if not os.path.exists(b200nb.lib_path()):  # fresh checkout: build the product library (nvcc cross-compiles without a GPU)
        subprocess.check_call(["make", "-C", REPO, "lib"])
    return b200nb


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
