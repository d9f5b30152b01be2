#!/usr/bin/env python3
"""Dev tool: run `reps` stage1+stage2 steps of one variant, then one steps(reps) call (ncu target).
usage: run_sweep.py N dtype variant reps"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
n = int(sys.argv[1]); dtype = sys.argv[2]; variant = int(sys.argv[3]); reps = int(sys.argv[4])
rng = np.random.default_rng(5)
r = ((rng.random((n, 3)) * 2 - 1) * 1e10).astype(dtype)
v = np.zeros((n, 3), dtype); m = ((rng.random(n) + 0.5) * 2).astype(dtype)
sh = _shim.Shard(n, dtype)
sh.upload(r, v, m, 6.6740831e-11, 2e12)
sh.set_variant(variant)
best = 1e30
for _ in range(reps):
    sh.stage1(); sh.stage2()
    t = sh.timings(); best = min(best, t['sweep_ms'])
sh.steps(reps)
steps_ms = sh.timings()['steps_ms'] / reps
print(json.dumps(dict(n=n, dtype=dtype, variant=variant, best_ms=best, steps_ms_per_step=steps_ms,
    tera_inter_s=n * (n - 1) / best / 1e9, sm_mhz=t['sm_mhz'], info=sh.info())))
