"""The reference-side patchers (oracle/patch_main.py, oracle/patch_test.py) change nothing but the registration branch /
the class under test.  Needs the reference sources, so it only runs in the build container."""
import difflib
import os
import subprocess
import sys

import pytest

from conftest import REPO

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF + "/src"), reason="reference sources not present")


def _changed(a_path, b_path):
    a, b = open(a_path).read().split("\n"), open(b_path).read().split("\n")
    removed, added = [], []
    for l in difflib.unified_diff(a, b, lineterm="", n=0):
        if l.startswith("-") and not l.startswith("---"):
            removed.append(l[1:])
        elif l.startswith("+") and not l.startswith("+++"):
            added.append(l[1:])
    return removed, added


def test_patch_test_retargets_only_the_class_under_test(tmp_path):
    subprocess.check_call([sys.executable, os.path.join(REPO, "oracle", "patch_test.py"), REF, str(tmp_path)])
    removed, added = _changed(REF + "/src/test/implem/test_SimulationNBody.cpp", str(tmp_path / "test_SimulationNBody_b200.cpp"))
    # removed: exactly the compile-time target switch (test_SimulationNBody.cpp:36-42)
    assert [l.strip() for l in removed] == [
        "#ifdef USE_CUDA", "CUDABodiesAllocator<float> targetAllocator(n, scheme);",
        "SimulationNBodyCUDATileFullDevice<float> simuTest(targetAllocator, soft);", "#else",
        "BodiesAllocator<float> targetAllocator(n, scheme);", "SimulationNBodyOpenMP<float> simuTest(targetAllocator, soft);", "#endif"]
    assert len(added) == 3 and sum("SimulationNBodyB200" in l for l in added) == 2 and any("B200BodiesAllocator" in l for l in added)
    # the loop, the golden model, the sections and their tolerances are the reference's
    text = (tmp_path / "test_SimulationNBody_b200.cpp").read_text()
    for needle in ('SimulationNBodyNaive<float> simuRef(naiveAllocator, soft);', 'WithinRel(xTest[b], e)',
                   'test_nbody_correctness(2048, 2e+08, 3600, 1, "random", 1e-3)', 'test_nbody_correctness(2049, 2e+08, 3600, 3, "galaxy", 1e-1)'):
        assert needle in text
    removed, added = _changed(REF + "/src/test/implem/test_CUDABodies.cpp", str(tmp_path / "test_CUDABodies_b200.cpp"))
    assert [l.strip() for l in removed] == ["CUDABodies<float> cudaBodies(n, scheme);"] * 2
    assert sorted(l.strip() for l in added) == sorted(["B200Bodies cudaBodies(n, scheme);"] * 2 + ['#include "SimulationNBodyB200.hpp" // gpu+b200'])


def test_patch_main_adds_one_branch(tmp_path):
    out = tmp_path / "main_b200.cpp"
    subprocess.check_call([sys.executable, os.path.join(REPO, "oracle", "patch_main.py"), REF + "/src/murb/main.cpp", str(out), "1"])
    removed, added = _changed(REF + "/src/murb/main.cpp", str(out))
    assert removed == []
    assert any('ImplTag == "gpu+b200"' in l for l in added) and any("SimulationNBodyB200.hpp" in l for l in added)
    assert len(added) <= 8
