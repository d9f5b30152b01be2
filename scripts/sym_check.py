#!/usr/bin/env python3
"""Dev tool (GPU box): parity and timing of the symmetric fp32 sweeps (variant ids 100+k)."""
import json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from oracle import oracle

G, T = 6.6740831e-11, 2.0e12
names = _shim.sym_variant_names()
def emit(**kw): print(json.dumps(kw), flush=True)

for n in (2, 33, 3001, 20011):
    r, v, m, _, _ = oracle.uniform_universe(n, 100 + n, 'float32')
    ref = oracle.stage1_f64(r, m, G)
    sh = _shim.Shard(n, 'float32')
    sh.upload(r, v, m, G, T)
    for k, name in enumerate(names):
        sh.set_variant(_shim.SYM_BASE + k)
        sh.stage1(); sh.stage2()
        rr, vv, a = sh.download(a=True)
        r_ref, v_ref = r.copy(), v.copy(); oracle.stage2(r_ref, v_ref, a, T)
        emit(kind='check', n=n, variant=name, max_rel=oracle.max_rel_err(a, ref), finite=bool(np.isfinite(a).all()),
             stage2_exact=bool(np.array_equal(rr, r_ref) and np.array_equal(vv, v_ref)), grid=sh.info()['grid'])
        sh.upload(r, v, m, G, T)
    sh.close()

sizes = [4096, 8192, 16384, 32768, 65536, 262144] + ([1048576] if '--big' in sys.argv else [])
for n in sizes:
    r, v, m, _, _ = oracle.uniform_universe(n, 11, 'float32')
    sh = _shim.Shard(n, 'float32')
    sh.upload(r, v, m, G, T)
    for vid, name in [(0, 'ordered0')] + [(_shim.SYM_BASE + k, nm) for k, nm in enumerate(names)]:
        sh.set_variant(vid)
        best = 1e30
        for _ in range(3):
            sh.stage1(); sh.stage2(); best = min(best, sh.timings()['sweep_ms'])
        emit(kind='time', n=n, variant=name, ms=best, tera=n * (n - 1) / best / 1e9, mhz=sh.timings()['sm_mhz'], grid=sh.info()['grid'])
    sh.close()
