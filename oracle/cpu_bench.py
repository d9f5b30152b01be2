# -*- coding: utf-8 -*-
"""cpu_bench.py — TEST/BENCH INFRASTRUCTURE (bench.py's cpu_baseline and `--impl reference` legs only).

Times the reference's CPU implementation of the hot path on the host cores: stage 1 = the reference's
fastest CPU kernel c4b (oracle/_ref/lib4.so, i.e. /root/reference/src/gravitation/kernel/_lib4_/lib.c
compiled by oracle/Makefile; zero-copy numpy binding as c4b.py:143-179) when it was built, otherwise
the oracle's OpenMP port (oracle.c); stage 2 = the four numpy passes of np2.py:110-115 (c4b itself falls
back to the per-object Python stage 2 of _base_.py:152-156, which would only make the baseline slower).
Run in a fresh process with OMP_NUM_THREADS set (c4a.py:63)."""

import time

import numpy as np

from . import oracle


def run(log2n, steps, warmup):
	n = 1 << log2n
	r, v, m, G, T = oracle.uniform_universe(n, 1000 + log2n, 'float32')
	threads = oracle.threads()
	if oracle.have_ref():
		kind = 'reference'
		c4 = oracle.RefC4(r, m, G)
		def stage1():
			for k in range(3): # c4b reads positions in place from SoA columns
				c4.cols[k][:] = r[:, k]
			return c4.stage1()
		what = 'c4b (_lib4_/lib.c, SSE + OpenMP, unique pairs) stage 1 + numpy stage 2 (np2.py:110-115)'
	else:
		kind = 'port'
		def stage1():
			return oracle.stage1_f32(r, m, G)
		what = 'oracle.c N x N float32 port (OpenMP) stage 1 + numpy stage 2'
	def step():
		nonlocal r, v
		a = stage1()
		r, v = oracle.np2_stage2(r, v, a, T)
	for _ in range(warmup):
		step()
	times = []
	for _ in range(steps):
		t0 = time.perf_counter()
		step()
		times.append(time.perf_counter() - t0)
	total = sum(times)
	inter = float(n) * float(n - 1)
	return {
		'kind': kind, 'threads': threads, 'n': n,
		'g_inter_s': inter * steps / total / 1e9, 'g_inter_s_best': inter / min(times) / 1e9,
		'ms_per_step': total / steps * 1e3,
		'sample': '%s; N=2^%d bodies of the same uniform universe, %d steps, %d threads; interactions '
			'credited as N(N-1) although the kernel exploits pair symmetry' % (what, log2n, steps, threads),
		}
