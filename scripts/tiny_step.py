#!/usr/bin/env python3
"""Dev tool: a few tiny steps (for compute-sanitizer runs)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gravitation_b200 import _shim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 700
dtype = sys.argv[2] if len(sys.argv) > 2 else 'float32'
rng = np.random.default_rng(1)
r = ((rng.random((n, 3)) * 2 - 1) * 1e10).astype(dtype)
v = np.zeros((n, 3), dtype)
m = ((rng.random(n) + 0.5) * 2).astype(dtype)
sh = _shim.Shard(n, dtype)
sh.upload(r, v, m, 6.6740831e-11, 2e12)
ids = list(range(len(_shim.variant_names(dtype)))) + [_shim.SYM_BASE + k for k in range(len(_shim.sym_variant_names(dtype)))]
if len(sys.argv) > 3:
    ids = [int(x) for x in sys.argv[3].split(',')]
for vi in ids:
    sh.set_variant(vi)
    sh.stage1(); sh.stage2()
sh.set_variant(-1)
sh.steps(3)
rr, vv, aa = sh.download(a=True)
print('ok', n, dtype, float(np.abs(aa).max()), bool(np.isfinite(rr).all()))
