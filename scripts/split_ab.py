"""A/B of the pass split (sw_set_pass_split 0 vs -1 = automatic) on shapes between one and a few rounds of
long strip-kernel work items; with the split off the old rule sends partly filled rounds to the
band-pipelined kernel.  Kernel-only GCUPS, best of 2 after warm-up; both matrices must be identical."""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("smith-waterman-fpga-module_b200")
SHAPES = (("10k query vs 100k x 1k (1.32 rounds)", 1, 10000, 100000, 1000), ("10k query vs 120k x 1k (1.58)", 1, 10000, 120000, 1000),
          ("10k query vs 151552 x 1k (2.00)", 1, 10000, 151552, 1000), ("10k query vs 260k x 1k (3.43)", 1, 10000, 260000, 1000),
          ("2k query vs 400k x 500 (5.3)", 1, 2000, 400000, 500), ("3 x 4k queries vs 60k x 2k (2.38)", 3, 4000, 60000, 2000))
for name, nq, ql, ns, sl in SHAPES:
    q = pkg.random_packed_db(nq, ql, 3)
    db = pkg.random_packed_db(ns, sl, 4)
    mats = []
    for mode in (0, -1):
        with pkg.Engine() as e:
            e.set_pass_split(mode)
            e.set_queries(q); e.load_db(db)
            ms = []
            for _ in range(3):
                e.score_db(); e.wait(); ms.append(e.last_kernel_ms)
            m = e.fetch_db()
            mats.append(int(m.astype(np.int64).sum()))
            print(json.dumps({"shape": name, "pass_split": mode, "kernel": e.last_kernel_name, "parts": e.last_pass_parts,
                              "gcups": round(e.last_cells / min(ms[1:]) / 1e6, 1), "ms": round(min(ms[1:]), 2), "err_bits": e.device_error_bits,
                              "checksum": mats[-1]}), flush=True)
    assert mats[0] == mats[1], name
