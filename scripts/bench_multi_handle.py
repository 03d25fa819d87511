#!/usr/bin/env python3
"""One handle driving several GPUs (sw_init with a gpu_ids list): end-to-end GCUPS through
sw_score_batch + sw_fetch with pinned host buffers.  usage: bench_multi_handle.py [ngpus] [subjects]"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("smith-waterman-fpga-module_b200")
ng = int(sys.argv[1]) if len(sys.argv) > 1 else pkg.device_count()
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 8_000_000
q = pkg.random_packed_db(100, 150, seed=1)
db = pkg.random_packed_db(ns, 150, seed=2)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
db = tuple(pin(x) for x in db)
out = torch.empty((100, ns), dtype=torch.int32, pin_memory=True).numpy()
for g in sorted({1, ng}):
    with pkg.Engine(gpu_ids=list(range(g))) as e:
        e.set_queries(q)
        e.score_batch(db); e.fetch(out=out)
        t0 = time.perf_counter()
        reps = 3
        e.score_batch(db)
        for k in range(reps):
            if k + 1 < reps:
                e.score_batch(db)
            e.fetch(out=out)
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"gpus_in_handle": g, "subjects": ns, "e2e_gcups": e.last_cells / dt / 1e9,
                          "ms_per_batch": dt * 1e3, "kernel_ms_max": e.last_kernel_ms, "kernel": e.last_kernel_name,
                          "checksum": int(out[:, ::4099].astype(np.int64).sum())}), flush=True)
