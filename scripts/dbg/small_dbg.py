#!/usr/bin/env python3
"""Dev tool (GPU box): where one step of the persistent small-N kernel spends its time — time stamps of the middle
CTA's last step from a -DSMALL_DEBUG build (gravitation_b200/libgravb200_dbg.so, built in the dev container with
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DSMALL_DEBUG -shared \
       -o gravitation_b200/libgravb200_dbg.so gravitation_b200/csrc/gravb200.cu -ldl).
usage: small_dbg.py N [dtype] [variant]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gravitation_b200 import _shim
_shim.LIB_PATH = os.path.join(ROOT, 'gravitation_b200', 'libgravb200_dbg.so')
from oracle import oracle
import cuda.bindings.runtime as rt
n = int(sys.argv[1]); dtype = sys.argv[2] if len(sys.argv) > 2 else 'float32'; vid = int(sys.argv[3]) if len(sys.argv) > 3 else -1
r, v, m, G, T = oracle.uniform_universe(n, 11, dtype)
sh = _shim.Shard(n, dtype)
sh.upload(r, v, m, G, T)
sh.set_variant(vid)
sh.steps(16)
nw = sh.info()['threads'] // 32
acc = []
for rep in range(5):
    sh.steps(64)
    host = np.zeros(128, dtype = np.uint64)
    rt.cudaMemcpy(host.ctypes.data, sh.device_ptr(4), 128 * 8, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    s = host[64:].astype(np.int64)
    t0 = s[0]
    landed = (s[3:3 + nw] - t0) / 1e3; done = (s[19:19 + nw] - t0) / 1e3
    acc.append(dict(barrier_us = (s[1] - t0) / 1e3, tma_issued_us = (s[2] - t0) / 1e3, slice_landed_us = [round(float(x), 2) for x in landed],
        compute_done_us = [round(float(x), 2) for x in done], reduced_us = (s[35] - t0) / 1e3, step_end_us = (s[36] - t0) / 1e3))
print(json.dumps(dict(n = n, dtype = dtype, info = sh.info(), per_step_us = sh.timings()['steps_ms'] / 64 * 1e3, last = acc[-1], median_step_end_us = float(np.median([a['step_end_us'] for a in acc])))))
sh.close()
