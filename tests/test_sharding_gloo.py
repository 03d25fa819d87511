"""N > 1 host logic on CPU: the shard plan of sw_load_db (sw_plan_shards, pure host code) and a
world_size-2 gloo run in which every rank scores its own shard (with the CPU oracle standing in
for the GPU) and rank 0 gathers -- the same no-collective-on-the-data-path structure bench.py
and a multi-GPU handle use.  Also checks bench.py --impl reference under a 2-rank launch."""
import json
import os
import random
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_shards_properties(pkg):
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 8):
        for ns in (0, 1, 7, 1000):
            ln = rng.integers(0, 400, size=ns).astype(np.uint32)
            st = pkg.plan_shards(ln, n)
            assert st[0] == 0 and st[-1] == ns and np.all(np.diff(st.astype(np.int64)) >= 0)
            if ns >= 100:
                tot = ln.sum()
                per = [ln[int(st[g]):int(st[g + 1])].sum() for g in range(n)]
                assert max(per) - min(per) <= 2 * 400 + tot * 0.01
    # uniform lengths split evenly
    st = pkg.plan_shards(np.full(1000, 150, np.uint32), 8)
    assert np.diff(st.astype(np.int64)).tolist() == [125] * 8


WORKER = r'''
import importlib, os, sys, random
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["SW_ROOT"])
pkg = importlib.import_module("smith-waterman-fpga-module_b200")
from oracle import oracle as om
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = random.Random(99)
queries = ["".join(rng.choice("ACGT") for _ in range(n)) for n in (40, 75)]
subjects = ["".join(rng.choice("ACGT") for _ in range(rng.randint(1, 120))) for _ in range(301)]
lens = np.array([len(s) for s in subjects], dtype=np.uint32)
starts = pkg.plan_shards(lens, world)
s0, s1 = int(starts[rank]), int(starts[rank + 1])
o = om.Oracle()
mine = np.array([[o.score(q, t) for t in subjects[s0:s1]] for q in queries], dtype=np.int32)
parts = [None] * world
dist.all_gather_object(parts, (s0, s1, mine))          # control plane only: results, not data
if rank == 0:
    full = np.zeros((len(queries), len(subjects)), dtype=np.int32)
    for a, b, m in parts:
        full[:, a:b] = m
    want = np.array([[o.score(q, t) for t in subjects] for q in queries], dtype=np.int32)
    assert np.array_equal(full, want)
    assert all(b > a for a, b, _ in parts)
    print("GATHER_OK", [(a, b) for a, b, _ in parts])
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_shard_and_gather(tmp_path, oracle_mod):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SW_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER_OK" in out.stdout


def test_bench_reference_arm_two_ranks(oracle_mod):
    """Under torchrun only rank 0 runs and prints the CPU arm; the other rank exits 0."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29732", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                          "--ref-subjects", "200"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "port"
    assert d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
