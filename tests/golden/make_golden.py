#!/usr/bin/env python3
"""Regenerates tests/golden/golden.json from the reference checkout.

Run in the build container only (needs /root/reference); the GPU box and the
test-suite read the committed JSON, never the reference tree.

Sources (all under /root/reference):
  data/*.fa                                  FASTA inputs (ScoreBank_v1_tb.sv:184-216 parsing rules)
  data/*_out.txt                             RTL simulation outputs (ScoreBank_v1_tb.sv:280-281)
  data/score.txt, data/score500.txt          ssearch36 -3 -n -R scores (6th field)
  data/sw_testing.txt:209-224                python swalign scores (first gap = -12 => gap_open=-8)
  capi_sample_aligner/software-C,C++/build/{query,library,main_test_output.txt}   CAPI end-to-end
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
DATA = os.path.join(REF, "data")
HERE = os.path.dirname(os.path.abspath(__file__))


def read_fasta_tokens(path):
    """Whitespace-token parser equivalent to the testbench's $fscanf("%s") loop."""
    toks = open(path).read().split()
    recs = []
    i = 0
    while i + 1 < len(toks):
        if toks[i].startswith(">"):
            recs.append([toks[i][1:], toks[i + 1]])
            i += 2
        else:
            break
    return recs


def parse_out(path):
    rows = []
    for line in open(path):
        m = re.match(r"@\s*(\d+)ns:\s+>(\S+) score:\s+(-?\d+)", line)
        if m:
            rows.append([m.group(2), int(m.group(3)), int(m.group(1))])
    return rows


def parse_ssearch(path):
    rows = []
    for line in open(path):
        if line.startswith("#") or line.startswith(">"):
            continue
        f = line.split()
        if len(f) >= 6 and f[0].startswith("db"):
            rows.append([f[0], int(f[5])])
    return rows


g = {"fasta": {}, "rtl": [], "ssearch": [], "format_samples": {}}
for fn in sorted(os.listdir(DATA)):
    if fn.endswith(".fa"):
        g["fasta"][fn] = read_fasta_tokens(os.path.join(DATA, fn))

for fn in sorted(os.listdir(DATA)):
    m = re.match(r"(data\d+\.fa)_(query\d+\.fa)_out\.txt", fn)
    if m:
        g["rtl"].append({"file": fn, "db": m.group(1), "query": m.group(2),
                         "rows": parse_out(os.path.join(DATA, fn))})

g["ssearch"].append({"file": "score.txt", "db": "data100.fa", "query": "query100.fa",
                     "rows": parse_ssearch(os.path.join(DATA, "score.txt"))})
g["ssearch"].append({"file": "score500.txt", "db": "data500.fa", "query": "query100.fa",
                     "rows": parse_ssearch(os.path.join(DATA, "score500.txt"))})

sw = []
for line in open(os.path.join(DATA, "sw_testing.txt")).read().splitlines()[208:224]:
    m = re.match(r"(db\d+):\s+(-?\d+)", line)
    if m:
        sw.append([m.group(1), int(m.group(2))])
g["swalign"] = {"db": "data1.fa", "query": "query1.fa",
                "params": {"match": 5, "mismatch": -4, "gap_open": -8, "gap_extend": -4},
                "rows": sw}

B = os.path.join(REF, "capi_sample_aligner", "software-C,C++", "build")
out = open(os.path.join(B, "main_test_output.txt")).read()
m = re.search(r"result: (-?\d+), biased: (\d+)", out)
g["capi"] = {"query": open(os.path.join(B, "query")).read().split()[0],
             "library": open(os.path.join(B, "library")).read().split()[0],
             "result": int(m.group(1)), "biased": int(m.group(2))}

# layout samples for the text writers
g["format_samples"]["out_txt_first_lines"] = \
    open(os.path.join(DATA, "data1.fa_query1.fa_out.txt")).read().splitlines()[:3]
g["format_samples"]["ssearch_R_first_lines"] = \
    open(os.path.join(DATA, "score500.txt")).read().splitlines()[:5]

n_rtl = sum(len(s["rows"]) for s in g["rtl"])
n_ss = sum(len(s["rows"]) for s in g["ssearch"])
print("rtl pairs", n_rtl, "ssearch pairs", n_ss, "swalign", len(sw), "capi", g["capi"]["result"])
with open(os.path.join(HERE, "golden.json"), "w") as f:
    json.dump(g, f, separators=(",", ":"))
