#!/usr/bin/env python3
"""A/B of the band-pipelined kernel's instances (SW_B200_WAVE_INSTANCE) over few-long-pair shapes.

  python scripts/wave_ab.py [instances ...]      one JSON line per (shape, instance)
"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("smith-waterman-fpga-module_b200")

SHAPES = [  # (name, queries, query length, subjects, subject length)
    ("1 x 100k / 100k", 1, 100000, 1, 100000),
    ("1 x 30k / 30k", 1, 30000, 1, 30000),
    ("1 x 100k / 8 x 50k", 1, 100000, 8, 50000),
    ("1 x 20k / 64 x 20k", 1, 20000, 64, 20000),
    ("1 x 20k / 256 x 20k", 1, 20000, 256, 20000),
    ("1 x 10k / 200 x 5k", 1, 10000, 200, 5000),
    ("1 x 10k / 500 x 2k", 1, 10000, 500, 2000),
    ("1 x 10k / 2000 x 1k", 1, 10000, 2000, 1000),
    ("1 x 10k / 2000 x 2k", 1, 10000, 2000, 2000),
    ("1 x 10k / 1000 x 1k", 1, 10000, 1000, 1000),
    ("1 x 10k / 500 x 1k", 1, 10000, 500, 1000),
    ("1 x 4k / 3000 x 1k", 1, 4000, 3000, 1000),
    ("1 x 50k / 100 x 10k", 1, 50000, 100, 10000),
    ("1 x 2k / 100 x 500", 1, 2000, 100, 500),
]


def main():
    insts = [int(a) for a in sys.argv[1:]] or [-1]
    for name, nq, ql, ns, sl in SHAPES:
        q = pkg.random_packed_db(nq, ql, 8)
        d = pkg.random_packed_db(ns, sl, 9)
        ref = None
        for i in insts:
            if i >= 0:
                os.environ["SW_B200_WAVE_INSTANCE"] = str(i)
            else:
                os.environ.pop("SW_B200_WAVE_INSTANCE", None)
            with pkg.Engine() as e:
                e.set_queries(q)
                e.load_db(d)
                ms = []
                for _ in range(6):
                    e.score_db()
                    e.wait()
                    ms.append(e.last_kernel_ms)
                got = e.fetch_db()
                if ref is None:
                    ref = got
                same = bool((got == ref).all())
                best = min(ms[1:])
                print(json.dumps({"shape": name, "instance": i, "kernel": e.last_kernel_name, "kernel_ms": round(best, 4),
                                  "gcups": round(e.last_cells / best / 1e6, 1), "same_as_first": same}), flush=True)


if __name__ == "__main__":
    main()
