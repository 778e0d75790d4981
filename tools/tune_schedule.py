#!/usr/bin/env python3
"""Offline schedule tuner for force_kernel: compiles every combination of the B200NB_KNOB_* statement-order knobs for a
kernel variant, extracts the inner loop from the SASS and scores it with the register-read model measured on B200
(profiles/r01_microbench_pipes.txt):

    cycles = sum over FP32 instructions of max(pipe cycles, #even source regs, #odd source regs)   [operands served by
             the `.reuse` cache are free; a MUFU between two instructions clobbers the cache]
           + per MUFU: 0.375 for each neighbouring FMA-pipe instruction that reads >= 4 vector registers, 0.09 for one
             that reads 3, 0 for 2 (profiles/r01_microbench_mufu_coissue.txt: 0.75 / 0.18 / 0.0 cycles per MUFU)

No GPU needed.  Prints the combinations sorted by estimated cycles per source pair.
    python tools/tune_schedule.py "256, 2, 2, 3, 1, false, 2, 3" [--jobs 8]
"""
import argparse
import itertools
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(REPO, "nbody-eurohpc_b200", "csrc", "force_sm100.cuh")


def inner_loop(sass):
    lines = [re.sub(r"/\*.*?\*/", "", l).strip() for l in sass.split("\n")]
    lines = [l for l in lines if l]
    # the hot loop is the backward-branch region with the most FFMA2
    best = []
    for i, l in enumerate(lines):
        if l.startswith("LDS.128"):
            body = []
            for m in lines[i:]:
                body.append(m)
                if m.startswith("BRA"):
                    break
            if sum("FFMA2" in x or "FFMA " in x for x in body) > sum("FFMA2" in x or "FFMA " in x for x in best):
                best = body
            
    return best


def _vec_reads(l):
    m = re.match(r"(FFMA2|FMUL2|FADD2|FFMA|FMUL|FADD)\s+(R\d+), (.*?) ;", l)
    if not m:
        return None
    regs = set()
    for s in m.group(3).split(", "):
        r = re.search(r"(?<![U])R(\d+)", s)
        if r:
            b = int(r.group(1))
            regs |= {b, b + 1} if "F32x2" in s else {b}
    return len(regs)


def mufu_cost(loop):
    cost = 0.0
    for i, l in enumerate(loop):
        if not l.startswith("MUFU"):
            continue
        prev = next((_vec_reads(loop[j]) for j in range(i - 1, -1, -1) if _vec_reads(loop[j]) is not None), 4)
        nxt = next((_vec_reads(loop[j]) for j in range(i + 1, len(loop)) if _vec_reads(loop[j]) is not None), 4)
        for n in (prev, nxt):
            cost += 0.375 if n >= 4 else (0.09 if n == 3 else 0.0)
    return cost


def score(loop):
    cache, total, mufu, acc, acc_reused, pairs = {}, 0, 0, 0, 0, 0
    for l in loop:
        m = re.match(r"(FFMA2|FMUL2|FADD2|FFMA|FMUL|FADD|MUFU\.RSQ)\s+(R\d+), (.*?) ;", l)
        if not m:
            continue
        op, srcs = m.group(1), m.group(3).split(", ")
        if op.startswith("MUFU"):
            mufu += 1
            cache = {}
            continue
        regs, newcache, hit = set(), {}, False
        for slot, s in enumerate(srcs):
            r = re.search(r"(?<![U])R(\d+)", s)
            if not r:
                continue
            base, pair = int(r.group(1)), "F32x2" in s
            if cache.get(slot) == (base, pair):
                hit = True
            else:
                regs |= {base, base + 1} if pair else {base}
            if "reuse" in s:
                newcache[slot] = (base, pair)
        cache = newcache
        ev = len([x for x in regs if x % 2 == 0])
        total += max(2 if op.endswith("2") else 1, ev, len(regs) - ev)
        if op == "FFMA2" and len(set(re.findall(r"(?<![U])R(\d+)\.(?:reuse\.)?F32x2", m.group(3)))) == 3:
            acc += 1
            acc_reused += hit
    n_pairs = mufu / 2 if mufu else 1
    mc = mufu_cost(loop)
    return {"cycles_per_pair": (total + mc) / n_pairs, "acc": acc, "acc_reused": acc_reused, "mufu": mufu,
            "mufu_cycles_per_pair": mc / n_pairs}


def build(args_str, defs, workdir, tag):
    cu = os.path.join(workdir, f"{tag}.cu")
    cubin = os.path.join(workdir, f"{tag}.cubin")
    with open(cu, "w") as f:
        f.write(f'#include "{HDR}"\nnamespace b200nb {{ template __global__ void force_kernel<{args_str}>(const ForceArgs); }}\n')
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", *defs, "-cubin", "-o", cubin, cu,
                        "-Xptxas", "-v"], capture_output=True, text=True)
    if r.returncode:
        return None
    regs = re.search(r"Used (\d+) registers", r.stderr)
    spill = re.search(r"(\d+) bytes spill stores", r.stderr)
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    s = score(inner_loop(sass))
    s["regs"] = int(regs.group(1)) if regs else -1
    s["spill"] = int(spill.group(1)) if spill else 0
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variant", help='template arguments, e.g. "256, 2, 2, 3, 1, false, 2, 3"')
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--top", type=int, default=12)
    a = ap.parse_args()
    combos = list(itertools.product(range(2), range(2), range(6), range(6)))
    with tempfile.TemporaryDirectory() as wd:
        def one(c):
            fm, lp, do, ao = c
            defs = [f"-DB200NB_KNOB_FMUL={fm}", f"-DB200NB_KNOB_LOOP={lp}", f"-DB200NB_KNOB_DORD={do}", f"-DB200NB_KNOB_AORD={ao}"]
            return c, build(a.variant, defs, wd, "k%d%d%d%d" % c)
        with ThreadPoolExecutor(a.jobs) as ex:
            res = [r for r in ex.map(one, combos) if r[1]]
    res.sort(key=lambda r: (r[1]["spill"] > 0, r[1]["cycles_per_pair"]))
    print(f"variant <{a.variant}>: {len(res)} combinations; (FMUL, LOOP, DORD, AORD)")
    for c, s in res[: a.top] + res[-3:]:
        print(f"  {c}  est {s['cycles_per_pair']:.2f} clk/pair (MUFU {s['mufu_cycles_per_pair']:.2f})  acc reuse {s['acc_reused']}/{s['acc']}"
              f"  regs {s['regs']} spill {s['spill']}")


if __name__ == "__main__":
    main()
