#!/usr/bin/env python3
"""BASELINE configs 2, 4 and 5 and the latency regime (config 3 is bench.py's headline).

  python scripts/bench_configs.py [2] [lat] [4] [4full] [4w] [5] [pair]      JSON lines, one per case
  bench.py imports bench_blocks() and adds its result to the bench line as `configs` / `latency`.

Every case is checked against the CPU oracle on a bounded sample inside the run (the oracle is the
checker here, never the thing measured), and reports its fraction of the ALU-pipe bound.
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SM_COUNT, SM_MHZ, R_INT, ALU_PER_PAIR = 148, 1965.0, 64.0, 3.5
PIPE_BOUND_GCUPS = SM_COUNT * SM_MHZ * 1e6 * R_INT * 2.0 / ALU_PER_PAIR / 1e9
# the reference's own (simulated) numbers for its data sets, BASELINE.md section 2
REF_BANK_US_DATA500 = 66.094 - 0.028      # data/data500.fa_query100.fa_out.txt:499, query loaded @28 ns
REF_PAIR_US = 0.538                       # data/data1.fa_query1.fa_out.txt:1 (simulation time)


def _oracle():
    from oracle import oracle as om
    om.build_oracle()
    return om.Oracle()


def _subset(db, idx):
    """Sub-database (packed, len, off) of the records idx."""
    packed, ln, off = db
    bufs, offs, o = [], [], 0
    for i in idx:
        nb = (int(ln[i]) + 3) // 4
        bufs.append(packed[int(off[i]): int(off[i]) + nb])
        offs.append(o)
        o += nb
    flat = np.concatenate(bufs + [np.zeros(16, np.uint8)])
    return flat, np.asarray(ln)[idx].astype(np.uint32), np.array(offs, dtype=np.uint64)


def mixed_db(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    lens = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.uint32)
    nbytes = (lens.astype(np.uint64) + 3) // 4
    off = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.uint64)
    packed = rng.integers(0, 256, size=int(nbytes.sum()) + 16, dtype=np.uint8)
    # unused tail bits of every record are zero, as sw_pack_2bit leaves them
    tail = (lens % 4).astype(np.int64)
    last = (off + nbytes - 1).astype(np.int64)
    m = tail > 0
    packed[last[m]] &= ((1 << (2 * tail[m])) - 1).astype(np.uint8)
    return packed, lens, off


def run(pkg, name, queries, db, reps=3, kernel=None, check=0, seed=0):
    """Kernel-time GCUPS on the resident database; `check` subjects are compared with the oracle."""
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
        res = {"config": name, "kernel": e.last_kernel_name, "cells": e.last_cells, "kernel_ms": best,
               "gcups": e.last_cells / best / 1e6, "pipe_frac": e.last_cells / best / 1e6 / PIPE_BOUND_GCUPS,
               "pass_parts": e.last_pass_parts}
        if check:
            got = e.fetch_db()
            ns = len(db[1])
            idx = np.unique(np.random.default_rng(seed).integers(0, ns, size=min(check, ns)))
            sub = _subset(db, idx)
            want, _ = _oracle().score_batch_packed(queries[0], queries[1], queries[2], sub[0], sub[1], sub[2])
            assert np.array_equal(got[:, idx], want), f"{name}: GPU scores differ from the oracle"
            res["oracle_checked_pairs"] = int(len(idx) * len(queries[1]))
            res["max_score"] = int(got.max())
    return res


def latency(pkg, name, queries, db, reps=300, check=True):
    """Wall clock of sw_score_batch + sw_fetch through the ABI (host buffers), after warm-up."""
    with pkg.Engine() as e:
        e.set_queries(queries)
        out = np.empty((len(queries[1]), len(db[1])), dtype=np.int32)
        for _ in range(30):
            e.score_batch(db); e.fetch(out=out)
        ts, kms = [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            e.score_batch(db); e.fetch(out=out)
            ts.append(time.perf_counter() - t0)
            kms.append(e.last_kernel_ms)
        ts.sort()
        med = ts[len(ts) // 2]
        res = {"config": name, "kernel": e.last_kernel_name, "e2e_us_median": med * 1e6, "e2e_us_min": ts[0] * 1e6,
               "e2e_us_p90": ts[int(len(ts) * 0.9)] * 1e6, "device_us_median": sorted(kms)[len(kms) // 2] * 1e3,
               "gcups_e2e": e.last_cells / med / 1e9, "cells": e.last_cells, "calls": reps}
        if check:
            want, _ = _oracle().score_batch_packed(queries[0], queries[1], queries[2], db[0], db[1], db[2])
            assert np.array_equal(out, want), f"{name}: GPU scores differ from the oracle"
            res["oracle_checked_pairs"] = int(want.size)
    return res


def c_latency(pkg, name, nq, qlen, ns, slen, seed, iters=3000):
    """The same job timed from C (bin/sw_b200_latency: sw_score_batch + sw_fetch in a loop, no Python in
    the timed region), with and without the CUDA events of the latency path; the scores of the last
    iteration are checked against the oracle."""
    import subprocess
    import tempfile
    seqio = importlib.import_module("smith-waterman-fpga-module_b200.seqio")
    exe = os.path.join(ROOT, "bin", "sw_b200_latency")
    if not os.path.exists(exe):
        return {"error": "bin/sw_b200_latency missing"}
    q = pkg.random_packed_db(nq, qlen, seed)
    db = pkg.random_packed_db(ns, slen, seed + 1)
    qs = [seqio.unpack_to_str(q[0], qlen, int(o)) for o in q[2]]
    ds = [seqio.unpack_to_str(db[0], slen, int(o)) for o in db[2]]
    o = _oracle()
    res = {"config": name, "timed_in": "C (bin/sw_b200_latency), clock_gettime around sw_score_batch + sw_fetch"}
    with tempfile.TemporaryDirectory() as td:
        qf, lf, sf = os.path.join(td, "q.fa"), os.path.join(td, "l.fa"), os.path.join(td, "s.txt")
        open(qf, "w").write("".join(f">q{i}\n{s}\n" for i, s in enumerate(qs)))
        open(lf, "w").write("".join(f">s{i}\n{s}\n" for i, s in enumerate(ds)))
        for ev in (1, 0):
            r = subprocess.run([exe, "-q", qf, "-l", lf, "-n", str(iters), "-e", str(ev), "-o", sf],
                               capture_output=True, text=True, timeout=120)
            if r.returncode != 0:
                return {"config": name, "error": (r.stdout + r.stderr)[-300:]}
            d = json.loads(r.stdout.strip().splitlines()[-1])
            got = np.array([int(l.split()[2]) for l in open(sf)], dtype=np.int32).reshape(nq, ns)
            want, _ = o.score_batch_packed(q[0], q[1], q[2], db[0], db[1], db[2])
            assert np.array_equal(got, want), f"{name}: C-driver scores differ from the oracle"
            key = "with_events" if ev else "without_events"
            res[key] = {k: d[k] for k in ("e2e_us_median", "e2e_us_min", "e2e_us_p90", "e2e_us_p99", "device_us_mean")}
            res["kernel"] = d["kernel"]
            res["cells"] = d["cells"]
    res["oracle_checked_pairs"] = nq * ns
    res["e2e_us_median"] = res["without_events"]["e2e_us_median"]
    res["device_us"] = res["with_events"]["device_us_mean"]
    res["gcups_e2e"] = res["cells"] / res["e2e_us_median"] / 1e3
    return res


def case_latency(pkg):
    a = latency(pkg, "2 (latency): 1 x 128 nt query vs 499 x 128 nt, sw_score_batch + sw_fetch",
                pkg.random_packed_db(1, 128, 1), pkg.random_packed_db(499, 128, 2))
    a["reference_simulated_us"] = REF_BANK_US_DATA500
    b = latency(pkg, "1 pair (latency): 32 nt query vs one 128 nt subject (the CAPI sample's job)",
                pkg.random_packed_db(1, 32, 1), pkg.random_packed_db(1, 128, 2))
    b["reference_simulated_us"] = REF_PAIR_US
    c = c_latency(pkg, "2 (latency, C driver): 1 x 128 nt query vs 499 x 128 nt", 1, 128, 499, 128, 1)
    c["reference_simulated_us"] = REF_BANK_US_DATA500
    d = c_latency(pkg, "1 pair (latency, C driver): 32 nt query vs one 128 nt subject", 1, 32, 1, 128, 3)
    d["reference_simulated_us"] = REF_PAIR_US
    return {"config2_499x128": c, "single_pair_32x128": d, "config2_499x128_python_ctypes": a,
            "single_pair_32x128_python_ctypes": b}


def case_4full(pkg):
    return run(pkg, "4: 1 x 10 kb query vs 1M x 1 kb subjects (full size, 1e13 cells)", pkg.random_packed_db(1, 10000, 3),
               pkg.random_packed_db(1000000, 1000, 4), reps=2, check=1000, seed=4)


def case_4(pkg, kernel=None):
    return run(pkg, "4 (200k): 1 x 10 kb query vs 200k x 1 kb (%s)" % (kernel or "automatic variant"),
               pkg.random_packed_db(1, 10000, 3), pkg.random_packed_db(200000, 1000, 4), kernel=kernel, check=300, seed=5)


def case_4w(pkg, kernel=None):
    return run(pkg, "4w: 1 x 10 kb query vs 2000 x 1 kb (%s)" % (kernel or "automatic variant"),
               pkg.random_packed_db(1, 10000, 3), pkg.random_packed_db(2000, 1000, 4), kernel=kernel, check=200, seed=6)


def case_5(pkg, kernel=None):
    return run(pkg, "5: 8 x (32..4096) queries vs 300k subjects log-uniform 32..4096 (%s)" % (kernel or "automatic variant"),
               mixed_db(8, 32, 4096, 5), mixed_db(300000, 32, 4096, 6), kernel=kernel, check=300, seed=7)


def case_pair(pkg, n=100000):
    """One long pair: the multi-warp (band-pipelined) path."""
    return run(pkg, f"single pair: {n} x {n} nt", pkg.random_packed_db(1, n, 8), pkg.random_packed_db(1, n, 9),
               reps=2, check=0)


def case_pair_overflow(pkg, n=100000):
    """One long, nearly identical pair (0.5 % substitutions): the score leaves 16 bits, so the pair is
    recomputed from the overflow list by the 32-bit band-pipelined kernel.  Whole call with host buffers
    (sw_score_batch + sw_fetch); the expected score is checked against a lower bound only (the oracle
    would need minutes): >= 5 x n - 9 x substitutions."""
    import random
    rng = random.Random(1)
    a = "".join(rng.choice("ACGT") for _ in range(n))
    b = list(a)
    nsub = n // 200
    for _ in range(nsub):
        b[rng.randrange(n)] = rng.choice("ACGT")
    b = "".join(b)
    with pkg.Engine() as e:
        e.score([a], [b])
        t0 = time.perf_counter()
        got = e.score([a], [b])
        dt = time.perf_counter() - t0
        score = int(got[0, 0])
        assert 5 * n - 9 * nsub <= score <= 5 * n, score
        return {"config": f"one nearly identical pair {n} x {n} nt, score beyond 16 bits (overflow list -> 32-bit band-pipelined pass)",
                "kernel": e.last_kernel_name + " + sw_wave32_kernel", "score": score, "call_ms": dt * 1e3, "kernel_ms": e.last_kernel_ms,
                "cells": n * n, "gcups_call": n * n / dt / 1e9}


def case_wave_mid(pkg):
    """A few hundred long pairs: 256 x 20 kb subjects against one 20 kb query (band-pipelined, 4 columns per step)."""
    return run(pkg, "1 x 20 kb query vs 256 x 20 kb subjects", pkg.random_packed_db(1, 20000, 8), pkg.random_packed_db(256, 20000, 9),
               reps=2, check=3, seed=5)


def case_penalties(pkg, n=2_000_000):
    """Run-time loadable penalties (ld_penalties, ScoreBank_v2.v:34,161): an arbitrary set with run-time
    operands, the same set specialised at run time (NVRTC), and the compiled-in default set, all on the
    same variant and the same 150-nt workload."""
    q = pkg.random_packed_db(100, 150, 11)
    db = pkg.random_packed_db(n, 150, 12)
    custom = dict(match=5, mismatch=-4, gap_open=-10, gap_extend=-3)
    name = "strip_s16x2_R25x2_G1_U4_F31"
    out = {"workload": f"{n} x 150 nt subjects vs 100 x 150 nt queries, {name}"}
    ok, msg = pkg.jit_compile_check(name, custom["gap_open"], custom["gap_extend"])
    out["nvrtc_specialisation"] = "ok" if ok == 1 else msg[:200]
    for label, params, jit in (("default_set_compiled_in", {}, 0), ("custom_set_runtime_operands", custom, 0),
                               ("custom_set_specialised_at_run_time", custom, 2)):
        with pkg.Engine(**params) as e:
            e.set_jit(jit)
            e.set_kernel_name(name)
            e.set_queries(q)
            e.load_db(db)
            ms = []
            for _ in range(4):
                e.score_db()
                e.wait()
                ms.append(e.last_kernel_ms)
            best = min(ms[1:])
            got = e.fetch_db()
            idx = np.arange(0, n, max(1, n // 400))
            sub = _subset(db, idx)
            from oracle import oracle as om
            want, _ = om.Oracle(**params).score_batch_packed(q[0], q[1], q[2], sub[0], sub[1], sub[2])
            assert np.array_equal(got[:, idx], want), f"penalties {label}: GPU scores differ from the oracle"
            out[label] = {"kernel": e.last_kernel_name, "gcups": e.last_cells / best / 1e6, "kernel_ms": best,
                          "oracle_checked_pairs": int(want.size)}
    d = out["default_set_compiled_in"]["gcups"]
    out["custom_runtime_vs_default"] = out["custom_set_runtime_operands"]["gcups"] / d
    out["custom_jit_vs_default"] = out["custom_set_specialised_at_run_time"]["gcups"] / d
    return out


def bench_blocks(pkg, log=lambda *a: None):
    """What bench.py adds to its JSON line (rank 0, N = 1): ~60 s in total."""
    out = {"configs": {}, "latency": None}
    t0 = time.perf_counter()
    for key, fn in (("latency", case_latency), ("config4_full", case_4full), ("config4_200k", case_4),
                    ("config4_2000_subjects", case_4w), ("config5_mixed", case_5), ("single_pair_100kb", case_pair),
                    ("single_pair_100kb_score_beyond_16_bits", case_pair_overflow), ("long_256_x_20kb", case_wave_mid),
                    ("run_time_penalties", case_penalties)):
        try:
            r = fn(pkg)
        except Exception as ex:
            r = {"error": repr(ex)}
        if key == "latency":
            out["latency"] = r
        else:
            out["configs"][key] = r
        log(key, json.dumps(r)[:400])
    out["configs"]["pipe_bound_gcups"] = PIPE_BOUND_GCUPS
    out["configs"]["seconds"] = time.perf_counter() - t0
    return out


if __name__ == "__main__":
    pkg = importlib.import_module("smith-waterman-fpga-module_b200")
    which = [a for a in sys.argv[1:]] or ["lat", "4", "4w", "5"]
    kernels = os.environ.get("SW_KERNELS", "").split() or [None]
    for w in which:
        if w in ("2", "lat"):
            for v in case_latency(pkg).values():
                print(json.dumps(v), flush=True)
        elif w == "4full":
            print(json.dumps(case_4full(pkg)), flush=True)
        elif w == "pair":
            print(json.dumps(case_pair(pkg)), flush=True)
        elif w == "ovf":
            print(json.dumps(case_pair_overflow(pkg)), flush=True)
        elif w == "mid":
            print(json.dumps(case_wave_mid(pkg)), flush=True)
        elif w == "pen":
            print(json.dumps(case_penalties(pkg)), flush=True)
        elif w in ("4", "4w", "5"):
            for k in kernels:
                print(json.dumps({"4": case_4, "4w": case_4w, "5": case_5}[w](pkg, k)), flush=True)
