import sys, os, json, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gravitation_b200 import _shim
from oracle import oracle
G, T = 6.6740831e-11, 2.0e12
n = 32768
r, v, m, _, _ = oracle.uniform_universe(n, 11, 'float32')
sh = _shim.Shard(n, 'float32'); sh.upload(r, v, m, G, T)
cudart = ctypes.CDLL('libcudart.so') if False else None
for vid in (103, 101):
    sh.set_variant(vid)
    for rep in range(5):
        sh.stage1(); sh.stage2(); t = sh.timings()
        ptr = sh.device_ptr(4)
        buf = torch.empty(2048, dtype=torch.int64)
        # copy device clk buffer through torch (cudaMemcpy via ctypes on libgravb200's cudart is not exported)
        src = torch.from_blob if False else None
        import cuda.bindings.runtime as rt
        host = np.zeros(2048, dtype=np.uint64)
        rt.cudaMemcpy(host.ctypes.data, ptr, 2048 * 8, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        g = sh.info()['grid']
        st = host[2:2 + 2 * g:2].astype(np.float64); en = host[3:3 + 2 * g:2].astype(np.float64)
        t0 = st.min()
        dur = (en - st) / 1e3; start = (st - t0) / 1e3; end = (en - t0) / 1e3
        order = np.argsort(-dur)[:6]
        print(vid, rep, 'sweep_ms', round(t['sweep_ms'], 3), 'kernel span us', round(end.max(), 1), 'start spread us', round(start.max(), 1),
              'dur us min/med/max', round(dur.min(), 1), round(np.median(dur), 1), round(dur.max(), 1), 'slowest ctas', [(int(i), round(dur[i], 1), round(start[i], 1)) for i in order], flush=True)
sh.close()
