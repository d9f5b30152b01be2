#!/usr/bin/env python3
"""Dev tool (GPU box): parity and timing of the symmetric fp64 sweeps."""
import json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from oracle import oracle
G, T = 6.6740831e-11, 2.0e12
names = _shim.sym_variant_names('float64')
for n in (2, 33, 3001, 20011):
    r, v, m, _, _ = oracle.uniform_universe(n, 100 + n, 'float64')
    ref = oracle.stage1_f64(r, m, G)
    sh = _shim.Shard(n, 'float64'); sh.upload(r, v, m, G, T)
    worst = 0; exact = True
    for k, name in enumerate(names):
        sh.set_variant(_shim.SYM_BASE + k)
        sh.stage1(); sh.stage2()
        rr, vv, a = sh.download(a=True)
        r_ref, v_ref = r.copy(), v.copy(); oracle.stage2(r_ref, v_ref, a, T)
        worst = max(worst, oracle.max_rel_err(a, ref)); exact = exact and bool(np.array_equal(rr, r_ref) and np.array_equal(vv, v_ref)) and bool(np.isfinite(a).all())
        sh.upload(r, v, m, G, T)
    print('check', n, 'max_rel', worst, 'stage2 exact & finite', exact, flush=True)
    sh.close()
for n in (16384, 65536, 262144):
    r, v, m, _, _ = oracle.uniform_universe(n, 11, 'float64')
    sh = _shim.Shard(n, 'float64'); sh.upload(r, v, m, G, T)
    for vid, name in [(0, 'ordered0')] + [(_shim.SYM_BASE + k, nm) for k, nm in enumerate(names)]:
        if n == 262144 and vid not in (0, 100, 101, 104, 105, 107): continue
        sh.set_variant(vid)
        best = 1e30
        for _ in range(3):
            sh.stage1(); sh.stage2(); best = min(best, sh.timings()['sweep_ms'])
        print(n, name, round(best, 4), 'ms', round(n * (n - 1) / best / 1e9, 3), 'T/s grid', sh.info()['grid'], flush=True)
    sh.close()
