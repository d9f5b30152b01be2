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


def _best(fn, steps, warmup):
	for _ in range(warmup):
		fn()
	best = 1e30
	for _ in range(steps):
		t0 = time.perf_counter()
		fn()
		best = min(best, time.perf_counter() - t0)
	return best


def matrix(budget_s = 40.0):
	"""SURVEY.md section 8d "CPU baselines (reported, not gating)" as one table, wall clock, best of >= 3 steps
	(the worker's min-of-steps rule, cli/worker.py:135-136): c4b raw stage 1 and stage 1 + numpy stage 2 at
	N = 2^12, 2^14, 2^16; c1a (the reference's scalar C kernel) at 2^14; the numpy restatement of np2 at 2^12
	(BASELINE.json configs[0]; bit-identical to the reference's np2 at float32, tests/test_oracle.py); the
	float64 OpenMP oracle at 2^16.  Rows whose turn comes after `budget_s` seconds are skipped and say so."""
	t_start = time.perf_counter()
	threads = oracle.threads()
	rows = []
	def add(kernel, log2n, what, seconds, kind, cores):
		n = 1 << log2n
		rows.append({'kernel': kernel, 'n': n, 'timed': what, 'best_s': seconds, 'g_inter_s': float(n) * float(n - 1) / seconds / 1e9,
			'cores': cores, 'kind': kind})
	def skipped(kernel, log2n):
		rows.append({'kernel': kernel, 'n': 1 << log2n, 'skipped': 'time budget of %.0f s used up' % budget_s})
	have_ref = oracle.have_ref()
	for log2n in (12, 14, 16):
		if time.perf_counter() - t_start > budget_s:
			skipped('c4b', log2n); continue
		n = 1 << log2n
		r, v, m, G, T = oracle.uniform_universe(n, 1000 + log2n, 'float32')
		if have_ref:
			c4 = oracle.RefC4(r, m, G)
			state = {'r': r, 'v': v}
			def full():
				for k in range(3):
					c4.cols[k][:] = state['r'][:, k]
				a = c4.stage1()
				state['r'], state['v'] = oracle.np2_stage2(state['r'], state['v'], a, T)
			add('c4b', log2n, 'raw C stage 1 (_lib4_/lib.c step_stage1)', _best(c4.stage1, 5, 2), 'reference', threads)
			add('c4b', log2n, 'stage 1 + numpy stage 2 (np2.py:110-115)', _best(full, 5, 1), 'reference', threads)
		else:
			add('oracle_f32', log2n, 'oracle.c N x N float32 (OpenMP) stage 1', _best(lambda: oracle.stage1_f32(r, m, G), 5, 1), 'port', threads)
	if time.perf_counter() - t_start <= budget_s:
		r, v, m, G, T = oracle.uniform_universe(1 << 14, 1014, 'float32')
		if have_ref:
			add('c1a', 14, 'raw C stage 1 (_lib1_/lib.c step_stage1, scalar, 1 thread)', _best(lambda: oracle.ref_c1a_stage1(r, m, G), 3, 1), 'reference', 1)
		else:
			add('oracle_pairs_f32', 14, 'oracle.c unique-pair float32 loop (1 thread)', _best(lambda: oracle.stage1_pairs_f32(r, m, G), 3, 1), 'port', 1)
	else:
		skipped('c1a', 14)
	if time.perf_counter() - t_start <= budget_s:
		r, v, m, G, T = oracle.uniform_universe(1 << 12, 1012, 'float32')
		state = {'r': r, 'v': v}
		def np2_step():
			a = oracle.np2_stage1(state['r'], m, G)
			state['r'], state['v'] = oracle.np2_stage2(state['r'], state['v'], a, T)
		add('np2', 12, 'full step of the numpy restatement of np2 (np2.py:89-115), 10 steps: BASELINE.json configs[0]', _best(np2_step, 10, 1), 'port', 1)
	else:
		skipped('np2', 12)
	if time.perf_counter() - t_start <= budget_s:
		r, v, m, G, T = oracle.uniform_universe(1 << 16, 1016, 'float32')
		add('oracle_f64', 16, 'oracle.c N x N float64 (OpenMP) stage 1 — the correctness oracle', _best(lambda: oracle.stage1_f64(r, m, G), 3, 1), 'port', threads)
	else:
		skipped('oracle_f64', 16)
	return {'threads': threads, 'rows': rows, 'seconds': time.perf_counter() - t_start}


def pick_log2n(per_step_budget_s, target_log2n = 20, probe_log2n = 14):
	"""largest N = 2^k <= 2^target whose c4b step fits `per_step_budget_s`, extrapolated from one probe step
	(O(N^2): x4 per doubling of N) — the reference arm runs the biggest sample of the workload it can afford"""
	n = 1 << probe_log2n
	r, v, m, G, T = oracle.uniform_universe(n, 1000 + probe_log2n, 'float32')
	if oracle.have_ref():
		c4 = oracle.RefC4(r, m, G)
		t = _best(c4.stage1, 2, 1)
	else:
		t = _best(lambda: oracle.stage1_f32(r, m, G), 2, 1)
	k = probe_log2n
	while k < target_log2n and t * 4.0 <= per_step_budget_s:
		t *= 4.0
		k += 1
	return k
