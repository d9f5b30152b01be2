#!/usr/bin/env python3
"""Dev tool (GPU box): the symmetric sweep with CTA ranges cut at tile granularity (split 0) against chunk
granularity (split 1, gravb200_set_split) — device time per step of steps(k) for the variants that have the twin,
over the mid-sized universes where a CTA holds only a few tiles, plus sampled parity against the float64 oracle.
  python scripts/split_ab.py [--quick] > gpurun_out/split_ab.jsonl"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from oracle import oracle

quick = '--quick' in sys.argv
plan = {
    'float32': ([100, 101, 102, 106, 111], [13312, 16384, 20000, 24576, 32768, 49152, 65536, 98304, 131072, 262144]),
    'float64': ([101, 102, 109], [8192, 12288, 16384, 24576, 32768, 65536, 131072]),
}
if quick:
    plan = {'float32': ([100, 101], [16384, 65536]), 'float64': ([101], [16384])}
for dtype, (variants, sizes) in plan.items():
    tol = 1e-4 if dtype == 'float32' else 1e-11
    for n in sizes:
        r, v, m, G, T = oracle.uniform_universe(n, 7, dtype)
        rows = np.unique(np.linspace(0, n - 1, 257).astype(np.int64))
        ref = oracle.stage1_f64(r, m, G, rows = rows)
        sh = _shim.Shard(n, dtype)
        sh.upload(r, v, m, G, T)
        # the automatic choice first (forced = -1, split = -1), then every twin in both modes
        for vid, split in [(-1, -1)] + [(vid, s) for vid in variants for s in (0, 1)]:
            sh.upload(r, v, m, G, T)
            sh.set_variant(vid)
            sh.set_split(split)
            sh.stage1(); sh.stage2()
            a = sh.download(r = False, v = False, a = True)[2]
            err = oracle.max_rel_err(a[rows], ref)
            info = sh.info()
            est_ms = max(n * (n - 1) / 2.5e12 * 1e3, 0.02)
            k = int(min(256, max(8, 60.0 / est_ms))) // 8 * 8
            sh.steps(8)
            best = 1e30
            for _ in range(3):
                sh.steps(k); best = min(best, sh.timings()['steps_ms'] / k)
            print(json.dumps(dict(dtype = dtype, n = n, forced = vid, split_mode = split, variant = info['variant'], split = info['split'],
                grid = info['grid'], tile = info['tile'], us_per_step = round(best * 1e3, 2), g_inter_s = round(n * (n - 1) / best / 1e6, 1),
                max_rel = err, ok = bool(err <= tol and np.isfinite(a).all()))), flush = True)
        sh.set_split(-1)
        sh.close()
