#!/usr/bin/env python3
"""Dev tool (GPU box): per-CTA life times + debug counters of symmetric variants (needs a -DSYM_DEBUG build).
usage: sym_dbg2.py LIB.so N variant_id [variant_id...]"""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gravitation_b200 import _shim
_shim.LIB_PATH = os.path.abspath(sys.argv[1])
from oracle import oracle
import cuda.bindings.runtime as rt
n = int(sys.argv[2])
r, v, m, G, T = oracle.uniform_universe(n, 11, 'float32')
sh = _shim.Shard(n, 'float32')
sh.upload(r, v, m, G, T)
for vid in [int(x) for x in sys.argv[3:]]:
    sh.set_variant(vid)
    for rep in range(4):
        sh.stage1(); sh.stage2()
        ms = (ctypes.c_float * 22)()
        host = np.zeros(2048, dtype=np.uint64)
        rt.cudaMemcpy(host.ctypes.data, sh.device_ptr(4), 2048 * 8, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        sh._lib.gravb200_timings(sh._ctx, ms, 22)
        g = sh.info()['grid']
        st = host[2:2 + 2 * g:2].astype(np.float64); en = host[3:3 + 2 * g:2].astype(np.float64)
        dur = (en - st) / 1e3; start = (st - st.min()) / 1e3
        order = np.argsort(-dur)[:5]
        nlog = int(host[2035])
        if nlog:
            ev = [(int(x >> 52), int((x >> 48) & 15), int((x >> 40) & 255), int((x >> 32) & 255), hex(int(x & 0xffffffff))) for x in host[1900:1900 + min(nlog, 60)]]
            print('diverged entries (cta, warp, k, chunk, mask):', ev[:40], flush=True)
        rt.cudaMemset(sh.device_ptr(4) + 2035 * 8, 0, 8)
        print(json.dumps(dict(n=n, vid=vid, rep=rep, sweep_ms=round(ms[0], 4), divergent=ms[10], div_points=[ms[12 + i] for i in range(8)], wait_Mcyc=round(ms[11], 2), grid=g,
            span_us=round(float((en.max() - st.min()) / 1e3), 1), start_spread_us=round(float(start.max()), 1),
            dur_us=[round(float(dur.min()), 1), round(float(np.median(dur)), 1), round(float(dur.max()), 1)],
            slowest=[(int(i), round(float(dur[i]), 1)) for i in order])), flush=True)
sh.close()
