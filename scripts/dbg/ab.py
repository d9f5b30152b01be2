#!/usr/bin/env python3
"""Dev tool (GPU box): best-of-3 sweep time of given variant ids with a given build of the library.
usage: ab.py LIB.so dtype log2N id [id...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gravitation_b200 import _shim
_shim.LIB_PATH = os.path.abspath(sys.argv[1])
from oracle import oracle
dtype = sys.argv[2]; n = 1 << int(sys.argv[3])
r, v, m, G, T = oracle.uniform_universe(n, 11, dtype)
sh = _shim.Shard(n, dtype)
sh.upload(r, v, m, G, T)
for vid in [int(x) for x in sys.argv[4:]]:
    sh.set_variant(vid)
    best = 1e30
    for _ in range(3):
        sh.stage1(); sh.stage2(); best = min(best, sh.timings()['sweep_ms'])
    print(json.dumps(dict(lib=os.path.basename(sys.argv[1]), dtype=dtype, n=n, vid=vid, ms=round(best, 4), tera=round(n * (n - 1) / best / 1e9, 4))), flush=True)
sh.close()
