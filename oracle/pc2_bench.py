# -*- coding: utf-8 -*-
"""pc2_bench.py — TEST/BENCH INFRASTRUCTURE (bench.py's baseline leg only).

Runs the reference's own GPU kernel `update_all` (pc2, /root/reference/src/gravitation/kernel/pc2.py:59-91),
compiled for sm_100a by oracle/build_pc2.py into oracle/_ref/pc2_<dtype>_<N>.cubin, with the reference's
launch shape (block 256, grid ceil(N/256), `pc2.py:142-145`) on cuda:0, and reports
  * kernel only (device-resident inputs), and
  * the reference's full step: 3 H2D copies + kernel + 3 D2H copies (`pc2.py:147-162`) + numpy stage 2
    (`pc2.py:164-168`), blocking like pycuda's memcpy_htod/dtoh.
PyCUDA is not installable here, so the module is loaded with cuda-python; torch supplies memory/events."""

import ctypes
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def available(n, dtype = 'float32'):
	return os.path.isfile(os.path.join(HERE, '_ref', 'pc2_%s_%d.cubin' % (dtype, n)))


def run(log2n, steps = 5, warmup = 2, dtype = 'float32'):
	import torch
	from cuda.bindings import driver as cu
	from . import oracle
	n = 1 << log2n
	path = os.path.join(HERE, '_ref', 'pc2_%s_%d.cubin' % (dtype, n))
	torch.cuda.init()
	torch.zeros(1, device = 'cuda') # creates the primary context
	def ck(res):
		err = res[0]
		if int(err) != 0:
			raise RuntimeError('CUDA driver error %s' % str(err))
		return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)
	mod = ck(cu.cuModuleLoadData(open(path, 'rb').read()))
	fn = ck(cu.cuModuleGetFunction(mod, b'update_all'))
	r, v, m, G, T = oracle.uniform_universe(n, 1000 + log2n, dtype)
	tdt = torch.float32 if dtype == 'float32' else torch.float64
	# host arrays as pc2 keeps them: Fortran-ordered (N,3) so columns are contiguous (pc2.py:108-112)
	r_h = np.asfortranarray(r); v_h = np.asfortranarray(v); a_h = np.zeros((n, 3), dtype = dtype, order = 'F')
	vt_h = np.zeros((n, 3), dtype = dtype, order = 'F')
	d = [torch.empty(n, dtype = tdt, device = 'cuda') for _ in range(7)] # rx ry rz ax ay az m
	d[6].copy_(torch.from_numpy(m))
	for k in range(3):
		d[k].copy_(torch.from_numpy(np.ascontiguousarray(r_h[:, k])))
	ptrs = [ctypes.c_void_p(t.data_ptr()) for t in d]
	args = (ctypes.c_void_p * 7)(*[ctypes.addressof(p) for p in ptrs])
	block, grid = 256, (n + 255) // 256
	stream = torch.cuda.current_stream().cuda_stream
	def launch():
		ck(cu.cuLaunchKernel(fn, grid, 1, 1, block, 1, 1, 0, stream, ctypes.addressof(args), 0))
	# kernel only
	for _ in range(warmup):
		launch()
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing = True), torch.cuda.Event(enable_timing = True)
	kt = []
	for _ in range(steps):
		e0.record(); launch(); e1.record(); torch.cuda.synchronize()
		kt.append(e0.elapsed_time(e1))
	# full reference step
	def step():
		for k in range(3):
			d[k].copy_(torch.from_numpy(r_h[:, k])) # memcpy_htod, blocking
		launch()
		for k in range(3):
			a_h[:, k] = d[3 + k].cpu().numpy() # memcpy_dtoh, blocking
		np.multiply(a_h, T, out = a_h)
		np.add(v_h, a_h, out = v_h)
		np.multiply(v_h, T, out = vt_h)
		np.add(r_h, vt_h, out = r_h)
	step()
	torch.cuda.synchronize()
	ft = []
	for _ in range(steps):
		t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ft.append(time.perf_counter() - t0)
	# sanity: accelerations of the reference kernel vs the float64 oracle on a few rows
	for k in range(3):
		d[k].copy_(torch.from_numpy(np.ascontiguousarray(r[:, k])))
	launch(); torch.cuda.synchronize()
	acc = np.stack([d[3 + k].cpu().numpy() for k in range(3)], axis = 1)
	rows = np.linspace(0, n - 1, 64).astype(np.int64)
	err = oracle.max_rel_err(acc[rows], oracle.stage1_f64(r, m, G, rows = rows))
	inter = float(n) * float(n - 1)
	return {
		'n': n, 'dtype': dtype,
		'kernel_ms': float(np.min(kt)), 'kernel_g_inter_s': inter / (float(np.min(kt)) * 1e-3) / 1e9,
		'step_ms': float(np.min(ft)) * 1e3, 'step_g_inter_s': inter / float(np.min(ft)) / 1e9,
		'max_rel_err_vs_f64': err,
		'what': 'reference pc2 `update_all` (pc2.py:59-91) compiled for sm_100a, block 256 x grid N/256; '
			'step = 3 H2D + kernel + 3 D2H + numpy stage 2 (pc2.py:147-168)',
		}
