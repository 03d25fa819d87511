import importlib, os, sys, json
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("smith-waterman-fpga-module_b200")
for name, ql, ns, sl in (("4k query vs 6k x 8k", 4000, 6000, 8000), ("4k query vs 50k x 4k", 4000, 50000, 4000), ("10k query vs 30k x 1k", 10000, 30000, 1000),
                         ("10k query vs 100k x 1k", 10000, 100000, 1000), ("10k query vs 151552 x 1k (2 rounds)", 10000, 151552, 1000), ("10k query vs 200k x 1k", 10000, 200000, 1000),
                         ("3 x 5k queries vs 20k x 3k", 5000, 20000, 3000), ("1.2k query vs 40k x 600", 1200, 40000, 600)):
    for mode in (1,):
        with pkg.Engine() as e:
            e.set_wave_mode(mode)
            e.set_queries(pkg.random_packed_db(3 if name.startswith('3 x') else 1, ql, 3)); e.load_db(pkg.random_packed_db(ns, sl, 4))
            ms = []
            for _ in range(3):
                e.score_db(); e.wait(); ms.append(e.last_kernel_ms)
            print(json.dumps({"shape": name, "wave_mode": mode, "kernel": e.last_kernel_name, "gcups": round(e.last_cells / min(ms[1:]) / 1e6, 1), "ms": round(min(ms[1:]), 2)}), flush=True)
