#!/usr/bin/env python3
"""Dev tool (GPU box, >= 2 GPUs): random N x dtype on all visible GPUs in one process (kernel module with
threads = P): partition (aligned or plain, ragged last shard), variant and exchange mode as the library
picks them; sampled accelerations against the float64 oracle, then 3 steps against one GPU.
usage: multi_fuzz.py [cases] [seed]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
from gravitation_b200.kernel import b200
from oracle import oracle
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
gpus = _shim.device_count()
bad = 0
for case in range(cases):
    dtype = 'float32' if case % 3 else 'float64'
    n = int(rng.integers(20000, 300000 if dtype == 'float32' else 150000))
    if case % 4 == 0:
        n = (n // 3072) * 3072 + int(rng.integers(0, 3))   # near block multiples
    r, v, m, G, T = oracle.uniform_universe(n, 700 + case, dtype)
    u = b200.universe(T=T, G=G, scale_off=True, dtype=dtype, threads=gpus)
    u.add_objects(r, v, m, scale_off=True)
    u.start()
    infos = [sh.info() for sh in u._shards]
    parts = [(sh.row0, sh.n_local) for sh in u._shards]
    u.step_stage1()
    a = np.array(u.accelerations())
    rows = np.unique(np.concatenate([np.linspace(0, n - 1, 256).astype(np.int64)] + [np.clip(np.arange(p0 - 2, p0 + 2), 0, n - 1) for p0, _ in parts]))
    err = oracle.max_rel_err(a[rows], oracle.stage1_f64(r, m, G, rows=rows))
    u.step_stage2(); u.step_stage3()
    u.steps(2)
    u.accelerations()
    r3 = np.array(u.mass_r_array)
    u.stop()
    sh = _shim.Shard(n, dtype)
    sh.upload(r, v, m, G, T)
    sh.steps(3)
    r1, _, _ = sh.download()
    sh.close()
    drift = float(np.abs(r3.astype(np.float64) - r1.astype(np.float64)).max() / np.abs(r1.astype(np.float64)).max())
    ok = err <= (1e-4 if dtype == 'float32' else 1e-11) and drift <= (1e-6 if dtype == 'float32' else 1e-13) and bool(np.isfinite(a).all())
    bad += 0 if ok else 1
    print(json.dumps(dict(case=case, gpus=gpus, dtype=dtype, n=n, parts=parts[:2] + parts[-1:], variant=infos[0]['variant'], mode=infos[0]['exchange_mode'],
                          max_rel=err, drift_vs_one_gpu=drift, ok=ok)), flush=True)
print('MULTI FUZZ', 'FAILED %d' % bad if bad else 'OK', cases)
sys.exit(1 if bad else 0)
