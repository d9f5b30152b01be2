#!/usr/bin/env python3
"""Dev tool (GPU box): time a few symmetric variants with a debug build of the library.
usage: sym_dbg.py LIB.so log2N [variant names...]   prints ms, divergent chunk entries, cycles in jbar waits"""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gravitation_b200 import _shim
_shim.LIB_PATH = os.path.abspath(sys.argv[1])
from oracle import oracle
n = 1 << int(sys.argv[2])
want = sys.argv[3:]
names = _shim.sym_variant_names()
r, v, m, G, T = oracle.uniform_universe(n, 11, 'float32')
sh = _shim.Shard(n, 'float32')
sh.upload(r, v, m, G, T)
for k, name in enumerate(names):
    if want and name not in want:
        continue
    sh.set_variant(_shim.SYM_BASE + k)
    out = []
    for _ in range(3):
        sh.stage1(); sh.stage2()
        ms = (ctypes.c_float * 22)()
        sh._lib.gravb200_timings(sh._ctx, ms, 22)
        out.append((round(ms[0], 3), ms[10], round(ms[11], 3)))
    print(json.dumps(dict(lib=os.path.basename(sys.argv[1]), n=n, variant=name, runs=out)), flush=True)
sh.close()
