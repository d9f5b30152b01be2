#!/usr/bin/env python3
"""Dev tool (GPU box): random sizes x symmetric variants against the float64 oracle (all rows) — different N
move the CTA ranges over row ends, diagonal blocks and ragged tiles, which is where the deferred j-combine
and the ring/tile barriers could go wrong.  Three steps per case: accelerations, bit-exact stage 2, and the
same state again from steps(3) (graph replay on small N).  usage: sym_fuzz.py [cases] [seed]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from oracle import oracle
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
bad = 0
for case in range(cases):
    dtype = 'float32' if case % 3 else 'float64'
    names = _shim.sym_variant_names(dtype)
    n = int(rng.integers(1100, 60000))
    vid = _shim.SYM_BASE + int(rng.integers(0, len(names)))
    r, v, m, G, T = oracle.uniform_universe(n, 500 + case, dtype)
    ref = oracle.stage1_f64(r, m, G)
    sh = _shim.Shard(n, dtype)
    sh.upload(r, v, m, G, T)
    sh.set_variant(vid)
    sh.stage1(); sh.stage2()
    r1, v1, a1 = sh.download(a=True)
    err = oracle.max_rel_err(a1, ref)
    r_ref, v_ref = r.copy(), v.copy(); oracle.stage2(r_ref, v_ref, a1, T)
    exact = bool(np.array_equal(r1, r_ref) and np.array_equal(v1, v_ref))
    sh.stage1(); sh.stage2(); sh.stage1(); sh.stage2()
    r3, v3, _ = sh.download()
    sh.upload(r, v, m, G, T)
    sh.steps(3)
    r3b, v3b, _ = sh.download()
    sh.close()
    den = np.abs(r3.astype(np.float64)).max()
    rep = float(np.abs(r3.astype(np.float64) - r3b.astype(np.float64)).max() / den)
    ok = err <= (1e-4 if dtype == 'float32' else 1e-11) and exact and rep <= (1e-6 if dtype == 'float32' else 1e-13) and bool(np.isfinite(a1).all())
    bad += 0 if ok else 1
    print(json.dumps(dict(case=case, dtype=dtype, n=n, variant=names[vid - _shim.SYM_BASE], max_rel=err, stage2_exact=exact, steps_vs_stages=rep, ok=ok)), flush=True)
print('FUZZ', 'FAILED %d' % bad if bad else 'OK', cases)
sys.exit(1 if bad else 0)
