"""tools/sass_resched.py at build time (no GPU): the committed order applies to the loop this compiler produces, the emitted
header carries a patched cubin, and an order derived from a different loop yields an EMPTY header (the library then keeps
launching the kernel ptxas scheduled)."""
import json
import os
import re
import shutil
import subprocess
import sys

import pytest

from conftest import REPO

ORDER = os.path.join(REPO, "nbody-eurohpc_b200", "csrc", "resched", "pk_t32_r8_tj2_st2_cta_u1_mb8.order.json")
VARIANT = "32, 8, 2, 2, 1, false, 1, 8, 1"
pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None, reason="needs nvcc and cuobjdump")


def _run(tmp_path, order):
    inc, cubin = tmp_path / "x.inc", tmp_path / "x.cubin"
    r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "sass_resched.py"), VARIANT, str(cubin), "--order-in", str(order),
                        "--emit-header", str(inc)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout, inc.read_text(), cubin


def test_committed_order_applies_and_is_a_permutation(tmp_path):
    rec = json.load(open(ORDER))
    assert sorted(rec["order"]) == list(range(rec["instructions"])) and rec["variant"] == VARIANT
    out, inc, cubin = _run(tmp_path, ORDER)
    m = re.search(r"(\d+) of (\d+) instructions moved", out)
    assert m and int(m.group(2)) == rec["instructions"] and int(m.group(1)) > 50
    size = int(re.search(r"force_resched_cubin_size = (\d+)ull", inc).group(1))
    assert size == os.path.getsize(cubin) > 10000
    assert "force_kernelILi32ELi8ELi2ELi2E" in inc
    # the patched loop keeps every instruction exactly once (the tool writes it next to the cubin)
    loop = [l.split("| ", 1)[1] for l in open(str(cubin) + ".txt").read().strip().split("\n")]
    assert len(loop) == rec["instructions"] and loop[-1].startswith("BRA")
    assert sum(l.startswith("FFMA2") for l in loop) == 96 and sum(l.startswith("MUFU") for l in loop) == 32


def test_foreign_order_gives_an_empty_header(tmp_path):
    rec = json.load(open(ORDER))
    rec["fingerprint"] = "0" * 64
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps(rec))
    out, inc, _ = _run(tmp_path, bad)
    assert "different loop" in out
    assert "force_resched_cubin_size = 0ull" in inc and "EMPTY" in inc
