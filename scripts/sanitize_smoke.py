#!/usr/bin/env python3
"""Small mixed workload for compute-sanitizer (memcheck): every kernel family once, tiny sizes."""
import importlib
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("smith-waterman-fpga-module_b200")
rng = random.Random(1)
R = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
queries = [R(1), R(150), R(333)]
subjects = [R(rng.randint(1, 260)) for _ in range(150)] + ["", "A"]
for name in ["", "strip_s16x2_R25x2_G1", "strip_s16x2_R38x1_G4", "strip_s16x2_R16x1_G32", "generic32"]:
    with pkg.Engine() as e:
        if name == "generic32":
            e.set_kernel_choice(0, 0, True, -1)
        elif name:
            e.set_kernel_name(name)
        s = e.score(queries, subjects)
        e.load_db(subjects); e.score_db(); e.wait(); b = e.fetch_best()
        print(name or "auto", e.last_kernel_name, int(s.sum()), b[0].tolist())
with pkg.Engine(score_width=12) as e:
    print("w12", int(e.score([R(500)], [R(500), R(30)]).sum()))
