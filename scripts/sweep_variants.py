#!/usr/bin/env python3
"""Dev tool (GPU box): time every kernel variant of libgravb200 and check it against a chunked
numpy float64 evaluation.  Not part of the product path.  Output: JSON lines on stdout."""
import json, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim

def universe(n, seed, dtype):
    rng = np.random.default_rng(seed)
    r = (rng.random((n, 3)) * 2.0 - 1.0) * 1e10
    m = (rng.random(n) + 0.5) * 2.0
    v = np.zeros((n, 3))
    return r.astype(dtype), v.astype(dtype), m.astype(dtype)

def ref_acc(r, m, G, rows=None):
    r = r.astype(np.float64); m = m.astype(np.float64)
    n = r.shape[0]
    rows = np.arange(n) if rows is None else rows
    out = np.zeros((len(rows), 3))
    for c0 in range(0, len(rows), 256):
        idx = rows[c0:c0 + 256]
        d = r[None, :, :] - r[idx, None, :]
        d2 = (d * d).sum(-1)
        d2[np.arange(len(idx)), idx] = np.inf
        s = m[None, :] / (d2 * np.sqrt(d2))
        out[c0:c0 + 256] = G * (d * s[:, :, None]).sum(1)
    return out

def relerr(a, b):
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)))

def emit(**kw):
    print(json.dumps(kw), flush=True)

G, T = 6.6740831e-11, 2.0e12
emit(kind='peak', **_shim.peak_probe(0))
emit(kind='peak2', **_shim.peak_probe(0))

quick = '--quick' in sys.argv
# ---- correctness on a ragged N
for dtype in ('float32', 'float64'):
    n = 3001
    r, v, m = universe(n, 7, dtype)
    ref = ref_acc(r, m, G)
    sh = _shim.Shard(n, dtype)
    sh.upload(r, v, m, G, T)
    names = _shim.variant_names(dtype)
    for vi, name in enumerate(names):
        sh.set_variant(vi)
        sh.stage1(); sh.stage2()
        _, _, a = sh.download(r=False, v=False, a=True)
        sh.upload(r, v, m, G, T)
        emit(kind='check', dtype=dtype, n=n, variant=vi, name=name, max_rel=relerr(a.astype(np.float64), ref), info=sh.info())
    sh.close()

# ---- timing
def time_variant(sh, reps):
    best = 1e30
    for _ in range(reps):
        sh.stage1(); sh.stage2()
        t = sh.timings()
        best = min(best, t['sweep_ms'])
    time_variant.mhz = t['sm_mhz']
    return best

for dtype, sizes in (('float32', [65536]), ('float64', [65536, 262144])):
    names = _shim.variant_names(dtype)
    results = {}
    for n in sizes:
        r, v, m = universe(n, 11, dtype)
        sh = _shim.Shard(n, dtype)
        sh.upload(r, v, m, G, T)
        for vi, name in enumerate(names):
            sh.set_variant(vi)
            time_variant(sh, 1)
            ms = time_variant(sh, 3)
            rate = n * (n - 1) / (ms * 1e-3) / 1e12
            results[(n, vi)] = rate
            emit(kind='time', dtype=dtype, n=n, variant=vi, name=name, sweep_ms=ms, tera_inter_s=rate, sm_mhz=time_variant.mhz, info=sh.info())
        sh.close()
    if dtype == 'float32' and not quick:
        n = 1 << 20
        top = [0]
        r, v, m = universe(n, 13, dtype)
        sh = _shim.Shard(n, dtype)
        sh.upload(r, v, m, G, T)
        rows = np.arange(0, n, n // 512)
        ref = ref_acc(r, m, G, rows)
        for vi in top:
            sh.set_variant(vi)
            time_variant(sh, 1)
            ms = time_variant(sh, 2)
            _, _, a = sh.download(r=False, v=False, a=True)
            emit(kind='time', dtype=dtype, n=n, variant=vi, name=names[vi], sweep_ms=ms,
                 tera_inter_s=n * (n - 1) / (ms * 1e-3) / 1e12, sm_mhz=time_variant.mhz, max_rel_sampled=relerr(a[rows].astype(np.float64), ref), info=sh.info())
            sh.upload(r, v, m, G, T)
        sh.close()
