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


def test_dependence_model_guards_the_hazards_found_on_hardware(tmp_path):
    """The dependence graph the search runs under, rebuilt from this compiler's SASS: values still pending at loop entry
    (soft^2 loaded by LDCU in the block-loop header on a barrier index the loop reuses; the target coordinates) are
    guarded by the first in-loop wait, MUFUs and LDSs keep their relative order (only the last MUFU of a pair carries a
    barrier), and the committed order honours every edge."""
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import sass_resched as sr

    cu, cubin = tmp_path / "k.cu", tmp_path / "k.cubin"
    cu.write_text(f'#include "{sr.HDR}"\nnamespace b200nb {{ template __global__ void force_kernel<{VARIANT}>(const ForceArgs); }}\n')
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-cubin", "-o", str(cubin), str(cu)])
    sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
    ins = sr.parse_sass(sass)
    s, e = sr.find_loop(ins)
    body = ins[s:e + 1]
    for x in body:
        x["c"] = sr.ctrl(x["hi"])
        x["op"], x["d"], x["s"] = sr.operands(x["text"])
    rec = json.load(open(ORDER))
    assert sr.loop_fingerprint(body) == rec["fingerprint"], "this compiler emits a different loop: re-run the search (DESIGN.md 3.1)"
    body[0]["entry"] = sr.pending_at_entry(ins, s)
    pend = {b: r for b, r in body[0]["entry"] if len(r) == 1 and next(iter(r)).startswith("UR")}
    assert pend, "the LDCU of soft^2 ahead of the loop was not found"
    (b_soft, res_soft), = pend.items()
    edges = sr.build_edges(body)
    order = rec["order"]
    sr.check_order(order, edges)                       # every same-iteration edge holds in the committed order
    pos = {i: k for k, i in enumerate(order)}
    # the first waiter on the LDCU's barrier precedes every reader of soft^2
    waiter = next(k for k, x in enumerate(body) if x["c"]["wait"] >> b_soft & 1)
    readers = [k for k, x in enumerate(body) if any(rs & res_soft for _, rs in x["s"]) and k != waiter]
    assert len(readers) >= 15 and all(pos[waiter] < pos[k] for k in readers)
    assert all((waiter, k, 2, 0) in set(edges) for k in readers)
    # an instruction that waits on a barrier nothing in the loop sets (the target loads) stays where it was
    set_inside = {b for x in body for b in (x["c"]["wr"], x["c"]["rd"]) if b != 7}
    pinned = [k for k, x in enumerate(body) if any((x["c"]["wait"] >> b & 1) and b not in set_inside for b in range(6))]
    assert pinned and all(order[k] == k for k in range(pinned[-1] + 1))
    # variable-latency instructions of one pipe keep their relative order
    for kind in ("MUFU", "LDS"):
        idx = [k for k, x in enumerate(body) if x["op"].startswith(kind)]
        assert [i for i in order if i in set(idx)] == idx
    # the loop ends with its branch, and the emitted stalls fit the 4-bit field
    words, total = sr.emit(body, order, edges)
    assert order[-1] == len(body) - 1 and all(sr.ctrl(hi)["stall"] <= 15 for _, hi in words)
