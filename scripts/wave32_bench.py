#!/usr/bin/env python3
"""One long, nearly identical pair: its score leaves 16 bits, so the pair is recomputed from the
overflow list -- by 256-row bands on many warps in 32-bit arithmetic (default) or, with
SW_B200_WAVE32=0, by ONE thread of the 32-bit scorer.

  python scripts/wave32_bench.py [length ...]
"""
import importlib
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("smith-waterman-fpga-module_b200")


def main():
    lens = [int(a) for a in sys.argv[1:]] or [20000, 100000]
    rng = random.Random(1)
    for n in lens:
        a = "".join(rng.choice("ACGT") for _ in range(n))
        b = list(a)
        for _ in range(n // 200):                      # 0.5 % substitutions
            b[rng.randrange(n)] = rng.choice("ACGT")
        b = "".join(b)
        for enable in (True, False):
            if not enable and n > 30000:
                continue                                # minutes on one thread
            with pkg.Engine() as e:
                e.set_overflow_wave(enable)
                e.score([a], [b])                      # warm-up (allocations)
                t0 = time.perf_counter()
                got = e.score([a], [b])
                dt = time.perf_counter() - t0
                print(json.dumps({"pair_nt": n, "overflow_wave": enable, "score": int(got[0, 0]), "call_ms": round(dt * 1e3, 2),
                                  "kernel_ms": round(e.last_kernel_ms, 3), "gcups_call": round(n * n / dt / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
