#!/usr/bin/env python3
"""Dev tool (GPU box): where the automatic choice changes hands between the persistent small-N kernel, the ordered
sweep and the symmetric sweep with chunk-granular CTA ranges — device time per step of steps(k), sampled parity.
  python scripts/mid_n.py > gpurun_out/mid_n.jsonl"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from oracle import oracle

plan = {
    'float32': ([(-1, -1), (200, -1), (101, 1), (106, 1), (102, 1)], [8192, 9472, 9600, 10240, 11264, 12288, 12800, 13312, 14336, 15360]),
    'float64': ([(-1, -1), (200, -1), (2, -1), (1, -1), (101, 1), (102, 1), (109, 1)], [4096, 4736, 5000, 5500, 6000, 7000, 8192, 9216, 10240, 11264, 12288, 14336]),
}
for dtype, (configs, sizes) in plan.items():
    tol = 1e-4 if dtype == 'float32' else 1e-11
    for n in sizes:
        r, v, m, G, T = oracle.uniform_universe(n, 7, dtype)
        rows = np.unique(np.linspace(0, n - 1, 257).astype(np.int64))
        ref = oracle.stage1_f64(r, m, G, rows = rows)
        sh = _shim.Shard(n, dtype)
        for vid, split in configs:
            try:
                sh.upload(r, v, m, G, T)
                sh.set_split(split)
                sh.set_variant(vid)
            except _shim.GravB200Error as e:
                continue
            sh.stage1(); sh.stage2()
            a = sh.download(r = False, v = False, a = True)[2]
            err = oracle.max_rel_err(a[rows], ref)
            info = sh.info()
            sh.steps(16)
            best = 1e30
            for _ in range(3):
                sh.steps(256); best = min(best, sh.timings()['steps_ms'] / 256)
            print(json.dumps(dict(dtype = dtype, n = n, forced = vid, split_mode = split, variant = info['variant'], split = info['split'],
                grid = info['grid'], us_per_step = round(best * 1e3, 2), g_inter_s = round(n * (n - 1) / best / 1e6, 1),
                max_rel = err, ok = bool(err <= tol and np.isfinite(a).all()))), flush = True)
        sh.close()
