#!/usr/bin/env python3
"""Instruction histogram of the hot loop of strip-kernel instances, from their SASS.

  python scripts/sass_hotloop.py [--lib PATH | --cubin PATH] [--match REGEX] [--json OUT] [--txt OUT]

For every matching sw_strip_kernel instance in the library (cuobjdump -sass), finds the innermost
loop that holds most of the packed DPX instructions (the 4-column step loop), and counts its
instructions by mnemonic and by pipe.  A trip of that loop covers 4 columns x R rows = 4 R cell
pairs (two subjects share every register), so
    alu_pipe_per_cell_pair   = ALU-pipe instructions / (4 R)
    issue_slots_per_cell_pair = all instructions / (4 R)
bench.py reads the JSON for its roofline denominator; the text form is the committed evidence
(profiles/r02_sass_hotloop.txt).  Pipe map as measured by microbench/pipe_pairs.cu
(profiles/r01_pipe_pairs_1024thr.json): packed min/max/add-max, LOP3, SHF, PRMT, ISETP, SEL, LEA on
the ALU pipe; VIADD.16x2, IMAD(.MOV), HFMA2-class on the FMA-side pipe; IADD3 dual-issues.
"""
import argparse
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_LIB = os.path.join(ROOT, "smith-waterman-fpga-module_b200", "libsw_b200.so")

ALU = ("VIADDMNMX", "VIMNMX", "VIMNMX3", "LOP3", "SHF", "PRMT", "ISETP", "SEL", "LEA", "PLOP3", "BMSK", "SGXT", "FLO", "POPC",
       "VABSDIFF", "IABS", "I2F", "F2I", "P2R", "R2P")
FMA = ("VIADD", "IMAD", "HFMA2", "HADD2", "FFMA", "FADD", "FMUL", "HMUL2")
EITHER = ("IADD3", "MOV", "IADD", "CS2R", "S2R")
MEM = ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDC", "LDCU", "ATOM", "ATOMG", "RED", "CCTL", "LDL", "STL", "ULDC")
CTRL = ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "NOP", "BAR", "CALL", "RET", "BREAK", "YIELD", "DEPBAR", "ERRBAR", "MEMBAR")
SHFL = ("SHFL", "REDUX", "VOTE", "VOTEU", "MATCH")


def pipe_of(op):
    base = op.split(".")[0]
    if base == "VIADD":
        return "fma"
    if base in ("VIADDMNMX", "VIMNMX", "VIMNMX3"):
        return "alu"
    if base in FMA:
        return "fma"
    if base in ALU:
        return "alu"
    if base in EITHER:
        return "either"
    if base in MEM:
        return "mem"
    if base in SHFL:
        return "shfl"
    if base in CTRL:
        return "ctrl"
    if base.startswith("U") or base.startswith("R2U"):
        return "uniform"
    return "other"


def sass_functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    fn, rows = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if fn:
                yield fn, rows
            fn, rows = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and fn:
            text = m.group(2).strip()
            pred = ""
            mm = re.match(r"(@!?U?P\d+)\s+(.*)", text)
            if mm:
                pred, text = mm.group(1), mm.group(2)
            rows.append((int(m.group(1), 16), text.split()[0], text, pred))
    if fn:
        yield fn, rows


def demangle_params(fn):
    m = re.search(r"sw_strip_kernelILi(\d+)ELi(\d+)ELi(\d+)ENS\w*8ArithS16ELb([01])ELi(\d+)ELi(\d+)ELi(n?\d+)ELi(n?\d+)ELb([01])ELi(\d+)(?:ELi(\d+))?", fn)
    U = int(m.group(10)) if m else 4
    FL = int(m.group(11)) if m and m.group(11) else 0
    if not m:
        m2 = re.search(r"sw_strip_kernelILi(\d+)ELi(\d+)ELi(\d+)ENS\w*8ArithS16ELb([01])ELi(\d+)ELi(\d+)ELi(n?\d+)ELi(n?\d+)ELi", fn)
        if not m2:
            return None
        g = m2.groups() + ("0",)
    else:
        g = m.groups()
    num = lambda s: -int(s[1:]) if s.startswith("n") else int(s)
    return {"RS": int(g[0]), "S": int(g[1]), "G": int(g[2]), "w12": g[3] == "1", "BT": int(g[4]), "MINB": int(g[5]),
            "goe": num(g[6]), "ge": num(g[7]), "direct": g[8] == "1", "U": U, "FL": FL}


def all_loops(rows):
    """(first index, last index, VIADDMNMX count) of every backward branch with packed DPX work."""
    addr_index = {a: i for i, (a, _, _, _) in enumerate(rows)}
    out = []
    for i, (a, op, text, _p) in enumerate(rows):
        if not op.startswith("BRA"):
            continue
        m = re.search(r"0x([0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addr_index:
            continue
        j = addr_index[tgt]
        n = sum(1 for r in rows[j:i + 1] if r[1].startswith("VIADDMNMX"))
        if n >= 16:
            out.append((j, i, n))
    return out


def general_loop(rows, hot):
    """Instances with interior trips (FL bit 0): the loop of the general trips = the smallest other
    DPX loop that does not overlap the hot (interior) loop."""
    h0 = rows.index(hot[0])
    h1 = h0 + len(hot) - 1
    cand = [(i - j, j, i) for j, i, n in all_loops(rows) if i < h0 or j > h1]
    if not cand:
        return None
    _, j, i = min(cand)
    return rows[j:i + 1]


def hot_loop(rows):
    """Innermost backward branch whose body holds the most VIADDMNMX (ties: the shortest body)."""
    addr_index = {a: i for i, (a, _, _, _) in enumerate(rows)}
    best = None
    for i, (a, op, text, _p) in enumerate(rows):
        if not op.startswith("BRA"):
            continue
        m = re.search(r"0x([0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addr_index:
            continue
        j = addr_index[tgt]
        body = rows[j:i + 1]
        n = sum(1 for r in body if r[1].startswith("VIADDMNMX"))
        if n == 0:
            continue
        key = (n / max(1, len(body)), n)
        # prefer the loop with the densest DPX body that still has a substantial count
        if best is None or (n >= 0.5 * best[1] and len(body) < best[2] and n >= 16) or n > 2 * best[1]:
            best = (key, n, len(body), j, i)
    if best is None:
        return None
    return rows[best[3]:best[4] + 1]


def analyse(path, pattern):
    res = {}
    for fn, rows in sass_functions(path):
        if "sw_strip_kernel" not in fn:
            continue
        p = demangle_params(fn)
        if not p:
            continue
        name = ("strip_s16x2_R%dx%d_G%d" % (p["RS"], p["S"], p["G"]) + ("_U%d" % p["U"] if p["U"] != 4 or p["FL"] else "")
                + ("_F%d" % p["FL"] if p["FL"] else ""))
        kind = "direct" if p["direct"] else "w12" if p["w12"] else ("fixed(%d,%d)" % (p["goe"], p["ge"])) if p["goe"] else "runtime"
        label = f"{name} [{kind}]"
        if pattern and not re.search(pattern, label):
            continue
        body = hot_loop(rows)
        if not body:
            continue
        hist = collections.Counter(op for _a, op, _t, _p in body)
        pipes = collections.Counter()
        for op, n in hist.items():
            pipes[pipe_of(op)] += n
        pairs = p["U"] * p["RS"] * p["S"]
        res[label] = {"variant": name, "instance": kind, "function": fn, "loop_instructions": len(body),
                      "cell_pairs_per_trip": pairs, "columns_per_trip": p["U"], "by_pipe": dict(pipes),
                      "alu_pipe_per_cell_pair": pipes["alu"] / pairs,
                      "fma_pipe_per_cell_pair": pipes["fma"] / pairs,
                      "issue_slots_per_cell_pair": len(body) / pairs,
                      "histogram": dict(sorted(hist.items(), key=lambda kv: -kv[1]))}
        if p["FL"] & 1:
            # interior + general trips: weight the two loop bodies by the trips of a 150-column pass
            # (sw_strip.cuh: interior trips cover [t_int0, t_int), the general loop the rest)
            gen = general_loop(rows, body)
            if gen:
                gpairs = (1 if p["FL"] & 16 else p["U"]) * p["RS"] * p["S"]
                galu = sum(1 for _a, op, _t, _p in gen if pipe_of(op) == "alu") / gpairs
                ahead = (5 if p["FL"] & 2 else 9) if p["FL"] & 8 else 9
                U, cols = p["U"], 150
                nsteps = (cols + p["S"] - 1 + U - 1) // U * U
                t_int = (cols - ahead) // U * U
                n_int = max(0, t_int - (0 if p["FL"] & 8 else U)) // U
                n_all = nsteps // U
                r = res[label]
                r["interior_alu_pipe_per_cell_pair"] = r["alu_pipe_per_cell_pair"]
                r["general_alu_pipe_per_cell_pair"] = galu
                r["general_issue_slots_per_cell_pair"] = len(gen) / gpairs
                r["interior_trips_of_150_columns"] = [n_int, n_all]
                r["alu_pipe_per_cell_pair"] = (n_int * r["interior_alu_pipe_per_cell_pair"] + (n_all - n_int) * galu) / n_all
                r["issue_slots_per_cell_pair"] = (n_int * r["issue_slots_per_cell_pair"] + (n_all - n_int) * len(gen) / gpairs) / n_all
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=DEFAULT_LIB)
    ap.add_argument("--cubin", default=None)
    ap.add_argument("--match", default=r"R25x2_G1\w* \[fixed\(-16,-4\)\]|R25x3_G1\w* \[fixed\(-16,-4\)\]|R38x2_G1\w* \[fixed\(-16,-4\)\]|R25x2_G1 \[runtime\]|R16x1_G32 \[fixed")
    ap.add_argument("--json", default=None)
    ap.add_argument("--txt", default=None)
    a = ap.parse_args()
    res = analyse(a.cubin or a.lib, a.match)
    lines = []
    for label, r in sorted(res.items()):
        lines.append(f"== {label}")
        lines.append(f"   hot loop: {r['loop_instructions']} instructions per trip = {r['columns_per_trip']} columns x "
                     f"{r['cell_pairs_per_trip'] // r['columns_per_trip']} rows = {r['cell_pairs_per_trip']} cell pairs")
        lines.append(f"   per cell pair: ALU pipe {r['alu_pipe_per_cell_pair']:.3f}, FMA-side pipe {r['fma_pipe_per_cell_pair']:.3f}, "
                     f"issue slots {r['issue_slots_per_cell_pair']:.3f}")
        if "general_alu_pipe_per_cell_pair" in r:
            lines.append(f"   (interior loop shown; ALU pipe {r['interior_alu_pipe_per_cell_pair']:.3f} in the interior trips, "
                         f"{r['general_alu_pipe_per_cell_pair']:.3f} in the general trips; the per-cell-pair figures above are "
                         f"weighted {r['interior_trips_of_150_columns'][0]} : {r['interior_trips_of_150_columns'][1] - r['interior_trips_of_150_columns'][0]} "
                         f"as in a pass over 150 columns)")
        lines.append("   by pipe: " + ", ".join(f"{k} {v}" for k, v in sorted(r["by_pipe"].items(), key=lambda kv: -kv[1])))
        lines.append("   " + ", ".join(f"{k} {v}" for k, v in r["histogram"].items()))
    text = "\n".join(lines)
    print(text)
    if a.txt:
        with open(a.txt, "w") as f:
            f.write("# scripts/sass_hotloop.py -- hot-loop instruction histogram (cuobjdump -sass of the shipped library)\n" + text + "\n")
    if a.json:
        kernels = {}
        for label, r in res.items():
            # bench.py looks the kernel up by variant name; the fixed(-16,-4) instance is what the bench runs
            if r["instance"].startswith("fixed(-16") or r["variant"] not in kernels:
                kernels[r["variant"]] = {k: r[k] for k in ("instance", "alu_pipe_per_cell_pair", "fma_pipe_per_cell_pair",
                                                            "issue_slots_per_cell_pair", "loop_instructions", "cell_pairs_per_trip")}
        with open(a.json, "w") as f:
            json.dump({"source": "scripts/sass_hotloop.py over libsw_b200.so (cuobjdump -sass)", "kernels": kernels,
                       "detail": res}, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
