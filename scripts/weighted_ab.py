#!/usr/bin/env python3
"""Dev tool (GPU box): chunk-granular CTA ranges cut by chunk COUNT (GRAVB200_SPLIT_WEIGHTED=0) against the
cost-weighted cut for several weights "w_sym,w_diag" of a chunk of a symmetric / diagonal tile (GRAVB200_SPLIT_W) —
device time per step of steps(k) with the twin forced, sampled parity.   python scripts/weighted_ab.py > gpurun_out/weighted_ab.jsonl"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from oracle import oracle

plan = {'float32': [9600, 12288, 16384, 24576, 32768, 49152], 'float64': [5000, 7000, 8192, 12288, 16384, 24576, 32768]}
weights = {'float32': [None, '4,3', '3,2', '8,5', '2,1'], 'float64': [None, '5,4', '4,3', '3,2', '2,1']}   # None: cut by chunk count (GRAVB200_SPLIT_WEIGHTED=0)
for dtype, sizes in plan.items():
    tol = 1e-4 if dtype == 'float32' else 1e-11
    for n in sizes:
        r, v, m, G, T = oracle.uniform_universe(n, 7, dtype)
        rows = np.unique(np.linspace(0, n - 1, 257).astype(np.int64))
        ref = oracle.stage1_f64(r, m, G, rows = rows)
        for w in weights[dtype]:
            os.environ['GRAVB200_SPLIT_WEIGHTED'] = '0' if w is None else '1'
            os.environ['GRAVB200_SPLIT_W'] = w or ''
            sh = _shim.Shard(n, dtype)
            sh.upload(r, v, m, G, T)
            sh.set_split(1)
            sh.stage1(); sh.stage2()
            a = sh.download(r = False, v = False, a = True)[2]
            err = oracle.max_rel_err(a[rows], ref)
            info = sh.info()
            est_ms = max(n * (n - 1) / 2.5e12 * 1e3, 0.02)
            k = int(min(256, max(8, 60.0 / est_ms))) // 8 * 8
            sh.steps(16)
            best = 1e30
            for _ in range(3):
                sh.steps(k); best = min(best, sh.timings()['steps_ms'] / k)
            print(json.dumps(dict(dtype = dtype, n = n, weights = w, variant = info['variant'], split = info['split'], grid = info['grid'],
                us_per_step = round(best * 1e3, 2), g_inter_s = round(n * (n - 1) / best / 1e6, 1), max_rel = err,
                ok = bool(err <= tol and np.isfinite(a).all()))), flush = True)
            sh.close()
