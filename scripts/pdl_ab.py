#!/usr/bin/env python3
"""Dev tool (GPU box): symmetric step with and without programmatic dependent launch (GRAVB200_PDL, read when a
context is created) — device time per step of steps(k), parity against the float64 oracle on sampled rows, and
steps(k) against k x (stage1, stage2).   python scripts/pdl_ab.py > gpurun_out/pdl_ab.jsonl"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from oracle import oracle

plan = {'float32': [9600, 12288, 16384, 24576, 32768, 49152, 65536, 262144], 'float64': [5000, 8192, 16384, 32768, 65536]}
for dtype, sizes in plan.items():
    tol = 1e-4 if dtype == 'float32' else 1e-11
    for n in sizes:
        r, v, m, G, T = oracle.uniform_universe(n, 7, dtype)
        rows = np.unique(np.linspace(0, n - 1, 257).astype(np.int64))
        ref = oracle.stage1_f64(r, m, G, rows = rows)
        for pdl in ('0', '1'):
            os.environ['GRAVB200_PDL'] = pdl
            sh = _shim.Shard(n, dtype)
            sh.upload(r, v, m, G, T)
            sh.stage1(); sh.stage2()
            a = sh.download(r = False, v = False, a = True)[2]
            err = oracle.max_rel_err(a[rows], ref)
            sh.stage1(); sh.stage2(); sh.stage1(); sh.stage2()
            r3 = sh.download()[0].astype(np.float64)
            sh.upload(r, v, m, G, T)
            sh.steps(3)
            r3b = sh.download()[0].astype(np.float64)
            rep = float(np.abs(r3 - r3b).max() / np.abs(r3).max())
            info = sh.info()
            est_ms = max(n * (n - 1) / 2.5e12 * 1e3, 0.02)
            k = int(min(256, max(8, 60.0 / est_ms))) // 8 * 8
            sh.steps(16)
            best = 1e30
            for _ in range(4):
                sh.steps(k); best = min(best, sh.timings()['steps_ms'] / k)
            print(json.dumps(dict(dtype = dtype, n = n, pdl = int(pdl), variant = info['variant'], split = info['split'], grid = info['grid'],
                us_per_step = round(best * 1e3, 2), g_inter_s = round(n * (n - 1) / best / 1e6, 1), max_rel = err, steps_vs_stages = rep,
                ok = bool(err <= tol and np.isfinite(a).all() and rep <= (1e-6 if dtype == 'float32' else 1e-13)))), flush = True)
            sh.close()
