#!/usr/bin/env python3
"""A/B of strip-kernel variants by name on one resident database (config-3 shape by default).

  python scripts/variant_ab.py [--subjects N] [--len L] [--queries Q] [--qlen M] [--reps K] name [name ...]

The database is generated and uploaded once; every variant scores it K+1 times (first pass untimed)
and its score matrix is compared with the first variant's.  One JSON line per variant (kernel time
from the library's CUDA events on its compute stream)."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("smith-waterman-fpga-module_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--subjects", type=int, default=4_000_000)
    ap.add_argument("--len", type=int, default=150)
    ap.add_argument("--queries", type=int, default=100)
    ap.add_argument("--qlen", type=int, default=150)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("names", nargs="+")
    a = ap.parse_args()
    q = pkg.random_packed_db(a.queries, a.qlen, seed=20160911)
    db = pkg.random_packed_db(a.subjects, a.len, seed=20160912)
    pkg.plant_homologs(db, q, 0.01, seed=20160919)
    ref = None
    with pkg.Engine(gpu_ids=[0]) as e:
        e.set_queries(q)
        e.load_db(db)
        for name in a.names:
            try:
                e.set_kernel_name(name)
                ms = []
                for _ in range(a.reps + 1):
                    e.score_db(); e.wait(); ms.append(e.last_kernel_ms)
                out = e.fetch_db()
                chk = int(out[:, :: max(1, a.subjects // 65536)].astype(np.int64).sum())
                if ref is None:
                    ref = out.copy()
                equal = bool(np.array_equal(out, ref))
                t = min(ms[1:])
                print(json.dumps({"kernel": e.last_kernel_name, "asked": name, "gcups": round(e.last_cells / t / 1e6, 1),
                                  "ms": round(t, 2), "all_ms": [round(x, 2) for x in ms], "equal_to_first": equal,
                                  "checksum": chk, "err_bits": e.device_error_bits}), flush=True)
            except Exception as ex:
                print(json.dumps({"asked": name, "error": repr(ex)}), flush=True)


if __name__ == "__main__":
    main()
