#!/usr/bin/env python3
"""Kernel-time GCUPS of BASELINE configs 2, 4 and 5 (config 3 is bench.py).  Prints JSON lines."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("smith-waterman-fpga-module_b200")


def run(name, queries, db, reps=3, kernel=None):
    with pkg.Engine() as e:
        if kernel:
            e.set_kernel_name(kernel)
        e.set_queries(queries)
        e.load_db(db)
        ms = []
        for _ in range(reps + 1):
            e.score_db()
            e.wait()
            ms.append(e.last_kernel_ms)
        best = min(ms[1:])
        print(json.dumps({"config": name, "kernel": e.last_kernel_name, "cells": e.last_cells,
                          "kernel_ms": best, "gcups": e.last_cells / best / 1e6}), flush=True)


def mixed_db(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    lens = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.uint32)
    nbytes = (lens.astype(np.uint64) + 3) // 4
    off = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.uint64)
    packed = rng.integers(0, 256, size=int(nbytes.sum()) + 16, dtype=np.uint8)
    return packed, lens, off


def latency(name, queries, db, reps=200):
    """Wall-clock of sw_score_batch + sw_fetch through the ABI (host buffers), after warm-up."""
    import time
    with pkg.Engine() as e:
        e.set_queries(queries)
        out = np.empty((len(queries[1]), len(db[1])), dtype=np.int32)
        for _ in range(20):
            e.score_batch(db); e.fetch(out=out)
        t0 = time.perf_counter()
        for _ in range(reps):
            e.score_batch(db); e.fetch(out=out)
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"config": name, "kernel": e.last_kernel_name, "e2e_us_per_call": dt * 1e6,
                          "kernel_us": e.last_kernel_ms * 1e3, "gcups_e2e": e.last_cells / dt / 1e9}), flush=True)


which = sys.argv[1:] or ["2", "4", "4w", "5"]
if "2" in which:    # query100 x data500 shape: 128-nt query, 499 x 128-nt subjects (latency bound)
    run("2: 1 x 128 nt query vs 499 x 128 nt", pkg.random_packed_db(1, 128, 1), pkg.random_packed_db(499, 128, 2), reps=5)
KERNELS = os.environ.get("SW_KERNELS", "").split()
if "2" in which or "lat" in which:
    latency("2 (latency): sw_score_batch + sw_fetch, 1 x 128 nt query vs 499 x 128 nt",
            pkg.random_packed_db(1, 128, 1), pkg.random_packed_db(499, 128, 2))
    latency("1 pair (latency): 32 nt query vs one 128 nt subject (the CAPI sample's job)",
            pkg.random_packed_db(1, 32, 1), pkg.random_packed_db(1, 128, 2))
if "4" in which:    # 10 kb query vs 1 kb subjects; 200k subjects = 2e12 cells per query
    for k in KERNELS or [None]:
        run("4: 1 x 10 kb query vs 200k x 1 kb (%s)" % (k or "automatic variant"), pkg.random_packed_db(1, 10000, 3),
            pkg.random_packed_db(200000, 1000, 4), kernel=k)
if "4full" in which:   # BASELINE config 4 at its stated size: 1 M x 1 kb subjects, one 10 kb query = 1e13 cells
    run("4 (full size): 1 x 10 kb query vs 1M x 1 kb (automatic variant)", pkg.random_packed_db(1, 10000, 3),
        pkg.random_packed_db(1000000, 1000, 4), reps=2)
if "4w" in which:   # same shape, few pairs: the warp-wide systolic (intra-task) variant
    run("4w: 1 x 10 kb query vs 2000 x 1 kb (automatic variant)", pkg.random_packed_db(1, 10000, 3),
        pkg.random_packed_db(2000, 1000, 4))
    run("4w: same, forced G=32 wavefront", pkg.random_packed_db(1, 10000, 3), pkg.random_packed_db(2000, 1000, 4),
        kernel="strip_s16x2_R16x1_G32")
if "5" in which:    # lengths log-uniform 32..4096
    for k in KERNELS or [None]:
        run("5: 8 x (32..4096) queries vs 300k subjects log-uniform 32..4096 (%s)" % (k or "automatic variant"),
            mixed_db(8, 32, 4096, 5), mixed_db(300000, 32, 4096, 6), kernel=k)
