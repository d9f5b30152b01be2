#!/usr/bin/env python3
"""Dev tool (GPU box): time per step at small N — automatic variant vs forced symmetric variants — through
steps(k) (device time per step, no host in between) and through stage1()+stage2() (wall clock, one sync per step)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from oracle import oracle
for dtype in ('float32', 'float64'):
    names = _shim.sym_variant_names(dtype)
    forced = {'float32': [106, 102], 'float64': [102, 106]}[dtype]
    for lg in range(4, 15):
        n = 1 << lg
        r, v, m, G, T = oracle.uniform_universe(n, 7, dtype)
        sh = _shim.Shard(n, dtype)
        sh.upload(r, v, m, G, T)
        for vid in [-1] + [f for f in forced if n >= 256]:
            sh.set_variant(vid)
            sh.steps(8)
            best_dev = 1e30
            for _ in range(3):
                sh.steps(64); best_dev = min(best_dev, sh.timings()['steps_ms'] / 64)
            best_wall = 1e30
            for _ in range(40):
                t0 = time.perf_counter(); sh.stage1(); sh.stage2(); best_wall = min(best_wall, (time.perf_counter() - t0) * 1e3)
            info = sh.info()
            print(json.dumps(dict(dtype=dtype, n=n, forced=vid, variant=info['variant'], grid=info['grid'], threads=info['threads'], r=info['bodies_per_thread'],
                dev_us_per_step=round(best_dev * 1e3, 2), wall_us_per_step=round(best_wall * 1e3, 2), g_inter_s=round(n * (n - 1) / best_dev / 1e6, 1))), flush=True)
        sh.close()
