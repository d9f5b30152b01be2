# -*- coding: utf-8 -*-
"""gravitation_b200 — a B200-native kernel for pleiszenburg/gravitation's n-body hot path.

Only what the path needs lives here:
  csrc/            CUDA kernels (sm_100a) + the C-ABI shim  -> libgravb200.so
  _shim.py         ctypes binding of include/gravb200.h
  kernel/          host-side mirror of the reference's kernel API (`_base_.py`) and the drop-in
                   kernel module `b200.py`
  lib/, cli/       the callers either side of the path: inventory, scenario builders, timers, worker
"""

__version__ = '0.1.0'
