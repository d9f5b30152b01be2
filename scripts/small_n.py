#!/usr/bin/env python3
"""Dev tool (GPU box): time per step at small N — the automatic choice, every persistent small-N variant
(ids SMALL_BASE + k) and the round-1 kernels (ordered / symmetric, forced) — through steps(k) (device time per
step, no host in between) and through stage1()+stage2() (wall clock, one sync per step).
  python scripts/small_n.py [--quick] > gpurun_out/small_n.jsonl"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from oracle import oracle

quick = '--quick' in sys.argv
K = 256
for dtype in ('float32', 'float64'):
    small = list(range(_shim.SMALL_BASE, _shim.SMALL_BASE + len(_shim.small_variant_names())))
    old = {'float32': [3, 2, 106, 101], 'float64': [2, 1, 102, 101]}[dtype]
    sizes = [16, 64, 256, 1024, 2048, 3000, 4096, 6000, 8192, 9472, 12288, 13312, 14000, 16384]
    if quick:
        sizes = [256, 4096, 8192]
    for n in sizes:
        r, v, m, G, T = oracle.uniform_universe(n, 7, dtype)
        sh = _shim.Shard(n, dtype)
        sh.upload(r, v, m, G, T)
        for vid in [-1] + small + old:
            try:
                sh.set_variant(vid)
            except _shim.GravB200Error as e:
                continue
            sh.steps(8)
            best_dev = 1e30
            for _ in range(3):
                sh.steps(K); best_dev = min(best_dev, sh.timings()['steps_ms'] / K)
            best_wall = 1e30
            for _ in range(40):
                t0 = time.perf_counter(); sh.stage1(); sh.stage2(); best_wall = min(best_wall, (time.perf_counter() - t0) * 1e3)
            info = sh.info()
            print(json.dumps(dict(dtype=dtype, n=n, forced=vid, variant=info['variant'], grid=info['grid'], threads=info['threads'],
                tile=info['tile'], dev_us_per_step=round(best_dev * 1e3, 2), wall_us_per_step=round(best_wall * 1e3, 2),
                g_inter_s=round(n * (n - 1) / best_dev / 1e6, 1))), flush=True)
        sh.close()
