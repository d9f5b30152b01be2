#!/usr/bin/env python3
"""Dev tool (GPU box): chunk-granular CTA ranges cut by chunk COUNT (GRAVB200_SPLIT_WEIGHTED=0) against the
cost-weighted cut (default: a chunk of a diagonal tile counts 3/4 (fp32) or 4/5 (fp64) of a symmetric one) —
device time per step of steps(k), sampled parity.  Rows with split_mode 1 force the twin where the automatic
rule keeps whole tiles.   python scripts/weighted_ab.py > gpurun_out/weighted_ab.jsonl"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from oracle import oracle

plan = {'float32': [9600, 12288, 13312, 16384, 20000, 24576, 32768, 49152, 65536, 98304], 'float64': [5000, 7000, 8192, 12288, 16384, 24576, 32768, 65536]}
for dtype, sizes in plan.items():
    tol = 1e-4 if dtype == 'float32' else 1e-11
    for n in sizes:
        r, v, m, G, T = oracle.uniform_universe(n, 7, dtype)
        rows = np.unique(np.linspace(0, n - 1, 257).astype(np.int64))
        ref = oracle.stage1_f64(r, m, G, rows = rows)
        for weighted, split in (('0', -1), ('1', -1), ('1', 1), ('1', 0)):
            os.environ['GRAVB200_SPLIT_WEIGHTED'] = weighted
            sh = _shim.Shard(n, dtype)
            sh.upload(r, v, m, G, T)
            sh.set_split(split)
            sh.stage1(); sh.stage2()
            a = sh.download(r = False, v = False, a = True)[2]
            err = oracle.max_rel_err(a[rows], ref)
            info = sh.info()
            est_ms = max(n * (n - 1) / 2.5e12 * 1e3, 0.02)
            k = int(min(256, max(8, 60.0 / est_ms))) // 8 * 8
            sh.steps(16)
            best = 1e30
            for _ in range(4):
                sh.steps(k); best = min(best, sh.timings()['steps_ms'] / k)
            print(json.dumps(dict(dtype = dtype, n = n, weighted = int(weighted), split_mode = split, variant = info['variant'], split = info['split'], grid = info['grid'],
                us_per_step = round(best * 1e3, 2), g_inter_s = round(n * (n - 1) / best / 1e6, 1), max_rel = err,
                ok = bool(err <= tol and np.isfinite(a).all()))), flush = True)
            sh.close()
