#!/usr/bin/env python3
# -*- coding: utf-8 -*-
"""bench.py — the driver's measurement contract for the hot path (all-pairs step, stage 1 + fused stage 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--bodies LOG2N] [--dtype f32|f64]
  torchrun ... bench.py --gpus N ...        (one rank per GPU)

Workload (BASELINE.json `metric`): synthetic uniform universe, N = 2^20 bodies, float32, the same N at
every GPU count (strong scaling: rows are sharded, positions all-gathered once per step).  One "step" =
one sweep over all N*(N-1) ordered interactions + the fused velocity/position update (+ the exchange).

Printed (rank 0, one JSON line): `value` = G interactions/s, device-timed (CUDA events on the launching
stream, summed over K steps, max over ranks) with state resident in HBM; `e2e` = the same metric through
the reference-facing kernel module with HOST buffers (pinned), H2D of the step's inputs and D2H of its
results inside the timed region; `roofline` against the FP32 non-tensor FMA peak (measured FFMA2
microbenchmark on this GPU, nominal beside it; 20 FLOP per interaction);
`cpu_baseline` = the reference's fastest CPU kernel (c4b, oracle/_ref/lib4.so) on this box's host cores
on a bounded sample.  `--impl reference` times that CPU kernel alone and prints the same line shape.
Nothing here reads /root/reference at run time."""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'G body-interactions/s at N=2^20 fp32'
UNIT = 'G interactions/s'
FLOP_PER_INTERACTION = 20.0 # SURVEY.md section 8d convention
G_SI, T_STEP = 6.6740831e-11, 2.0e12


def uniform_universe(n, seed, dtype):
	"""SURVEY.md section 8d synthetic universe (i): post-scaling magnitudes of the galaxy scenario"""
	rng = np.random.default_rng(seed)
	r = (rng.random((n, 3)) * 2.0 - 1.0) * 1.0e10
	m = (rng.random(n) + 0.5) * 2.0
	v = np.zeros((n, 3))
	return r.astype(dtype), v.astype(dtype), m.astype(dtype)


# -------------------------------------------------------------------------------------------------
# clocks (NVML) sampled during the timed region
# -------------------------------------------------------------------------------------------------

class ClockSampler:
	REASONS = {
		0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
		0x80: 'hw_power_brake_slowdown', 0x2: 'applications_clocks_setting', 0x10: 'sync_boost',
		}

	def __init__(self, index):
		self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
		self._stop = threading.Event()
		try:
			import pynvml
			pynvml.nvmlInit()
			self._nv = pynvml
			self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
			self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
			self.ok = True
		except Exception:
			self.ok = False
		self._t = threading.Thread(target = self._run, daemon = True)

	def _run(self):
		while not self._stop.is_set():
			try:
				self.samples.append(int(self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM)))
				mask = int(self._nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
				for bit, name in self.REASONS.items():
					if mask & bit:
						self.reasons.add(name)
			except Exception:
				pass
			self._stop.wait(0.1)

	def __enter__(self):
		if self.ok:
			self._t.start()
		return self

	def __exit__(self, *exc):
		self._stop.set()
		if self.ok:
			self._t.join()

	def summary(self):
		if not self.samples:
			return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}
		return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


# -------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own C kernel (c4b) on the host cores
# -------------------------------------------------------------------------------------------------

def cpu_reference_run(log2n, steps, warmup, threads):
	"""fresh process (OMP_NUM_THREADS must be set before libgomp loads, c4a.py:63); returns dict"""
	env = dict(os.environ, OMP_NUM_THREADS = str(threads))
	code = (
		'import sys, json; sys.path.insert(0, %r)\n'
		'from oracle import cpu_bench\n'
		'print(json.dumps(cpu_bench.run(%d, %d, %d)))\n'
		) % (ROOT, log2n, steps, warmup)
	out = subprocess.run([sys.executable, '-c', code], env = env, capture_output = True, text = True, timeout = 1500)
	if out.returncode != 0:
		raise RuntimeError('cpu reference run failed: %s' % out.stderr[-2000:])
	return json.loads(out.stdout.strip().splitlines()[-1])


def reference_arm(args):
	rank = int(os.environ.get('RANK', 0))
	if rank != 0:
		return 0
	threads = os.cpu_count() or 1
	log2n = 16 # bounded sample of the 2^20 workload: c4b needs ~N^2/2 pair updates per step
	res = cpu_reference_run(log2n, args.steps, args.warmup, threads)
	line = {
		'impl': 'reference',
		'metric': METRIC, 'value': res['g_inter_s'], 'unit': UNIT,
		'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
		'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong',
		'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
		'config': {'workload': 'all-pairs step (stage 1 + stage 2), uniform universe, N=2^20 fp32',
			'n_bodies': 1 << 20, 'parallelism': 'host threads'},
		'cpu_baseline': {'value': res['g_inter_s'], 'unit': UNIT, 'cores': res['threads'], 'kind': res['kind'],
			'sample': res['sample']},
		'e2e': {'value': res['g_inter_s'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
		'gpu_launches': 0,
		}
	_emit(line)
	return 0


# -------------------------------------------------------------------------------------------------
# own arm
# -------------------------------------------------------------------------------------------------

def own_arm(args):
	import torch
	from gravitation_b200 import _shim, dist
	from gravitation_b200.kernel import b200

	rank, world, local_rank = dist.init_process_group()
	if world != args.gpus:
		if world == 1 and args.gpus > 1:
			raise SystemExit('--gpus %d needs torchrun with %d ranks (one process per GPU)' % (args.gpus, args.gpus))
	torch.cuda.set_device(local_rank)
	dtype = {'f32': 'float32', 'f64': 'float64'}[args.dtype]
	n = 1 << args.bodies
	interactions = float(n) * float(n - 1)
	r, v, m = uniform_universe(n, 1000 + args.bodies, dtype)

	# L2 flush between timed steps: write a buffer larger than the 126 MB L2
	flush = torch.empty(256 << 20, dtype = torch.uint8, device = 'cuda')

	def flush_l2():
		flush.zero_()
		torch.cuda.synchronize()

	# ---- device-resident arm: C-ABI shard, state in HBM ------------------------------------------
	shard = dist.make_shard(n, dtype)
	shard.upload(r, v, m, G_SI, T_STEP)

	def one_step():
		flush_l2()
		dist.barrier() if world > 1 else None
		shard.stage1()
		shard.stage2()
		t = shard.timings()
		return t['sweep_ms'] + (t['exchange_ms'] if world > 1 else 0.0), t

	# warm-up step 1 doubles as the parity sample: its accelerations belong to the initial positions
	flush_l2()
	shard.stage1(); shard.sync()
	_, _, a0 = shard.download(r = False, v = False, a = True)   # this rank's rows
	shard.stage2()
	for _ in range(args.warmup - 1):
		one_step()
	launches0 = shard.info()['launches']
	torch.cuda.synchronize(); dist.barrier() if world > 1 else None
	wall0 = time.perf_counter()
	step_ms, sweep_ms, xchg_ms, sm_mhz = [], [], [], []
	with ClockSampler(local_rank) as clocks:
		for _ in range(args.steps):
			ms, t = one_step()
			step_ms.append(ms); sweep_ms.append(t['sweep_ms']); xchg_ms.append(max(t['exchange_ms'], 0.0)); sm_mhz.append(t['sm_mhz'])
	torch.cuda.synchronize(); dist.barrier() if world > 1 else None
	wall_ms = (time.perf_counter() - wall0) * 1e3
	launches = shard.info()['launches'] - launches0
	total_ms = dist.max_over_ranks(sum(step_ms)) if world > 1 else sum(step_ms)
	info = shard.info()
	rows_rank0 = shard.n_local
	value = interactions * args.steps / (total_ms * 1e-3) / 1e9

	shard.close()

	# ---- end-to-end arm: reference-facing kernel module, host buffers ----------------------------
	kw = dict(T = T_STEP, G = G_SI, scale_off = True, dtype = dtype, eager_host = True)
	if world > 1:
		uid = dist.broadcast_bytes(_shim.nccl_unique_id() if rank == 0 else None)
		kw.update(rank = rank, world = world, nccl_id = uid, device = local_rank)
	else:
		kw.update(device = local_rank)
	u = b200.universe(**kw)
	u.add_objects(r, v, m, scale_off = True)
	u.start()
	esz = np.dtype(dtype).itemsize
	n_local = u._shards[0].n_local
	h2d = (n * 3 + n_local * 3 + n) * esz          # push_host_state: r (all), v (own rows), m
	d2h = (n_local * 3) * esz + (n * 3 + n_local * 3) * esz   # a after stage 1; r (all), v (own rows) after stage 2

	def e2e_step():
		u.push_host_state()   # H2D of the step's inputs from the pinned host mirrors
		u.step()              # stage 1 (+ D2H of a), stage 2 (+ D2H of r, v), eager_host = True
	e2e_warm = 1 if args.quick_e2e else min(args.warmup, 3)
	for _ in range(e2e_warm):
		e2e_step()
	e2e_steps = 1 if args.quick_e2e else max(3, min(args.steps, 5))
	torch.cuda.synchronize(); dist.barrier() if world > 1 else None
	t0 = time.perf_counter()
	for _ in range(e2e_steps):
		e2e_step()
	torch.cuda.synchronize(); dist.barrier() if world > 1 else None
	e2e_s = dist.max_over_ranks(time.perf_counter() - t0) if world > 1 else (time.perf_counter() - t0)
	e2e_value = interactions * e2e_steps / e2e_s / 1e9
	checksum = float(np.abs(u.mass_r_array).sum())
	u.stop()

	if rank != 0:
		return 0

	# ---- roofline: FP32 non-tensor FMA -------------------------------------------------------------
	probe = _shim.peak_probe(local_rank)
	sm_count = info['sm_count']
	sm_max = clocks.summary()['sm_max_mhz'] or 1965
	lanes = 128 if dtype == 'float32' else 64
	peak_nominal = sm_count * lanes * 2 * sm_max * 1e6 / 1e12
	peak_measured = max(probe['fp32x2_tflops'], probe['fp32_tflops']) if dtype == 'float32' else probe['fp64_tflops']
	per_gpu_rate = value * 1e9 / world
	kernel_s = float(np.mean(sweep_ms)) * 1e-3
	achieved = (interactions / world) * FLOP_PER_INTERACTION / kernel_s / 1e12   # dominant kernel, per GPU
	traffic = None
	tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
	if os.path.isfile(tpath):
		try:
			traffic = json.load(open(tpath)).get('%s_2p%d' % (args.dtype, args.bodies))
		except Exception:
			traffic = None
	symmetric = info.get('variant', 0) >= _shim.SYM_BASE
	# executed work per ORDERED interaction (the metric's unit): ordered sweep 12 FP32-pipe lane-ops = 19 FLOP,
	# symmetric sweep (every unordered pair once, both bodies updated) 8 lane-ops = 13 FLOP; fp64: 16 / 10 ops
	lane_ops = (8.0 if symmetric else 12.0) if dtype == 'float32' else (10.0 if symmetric else 16.0)
	flop_exec = (13.0 if symmetric else 19.0) if dtype == 'float32' else (16.0 if symmetric else 25.0)
	pipe_ceiling = sm_count * lanes * sm_max * 1e6 / lane_ops / 1e9   # G interactions/s if the FP pipe never idled
	roofline = {
		'bound': 'fp32_fma' if dtype == 'float32' else 'fp64_fma',
		'achieved': achieved, 'peak': peak_measured, 'unit': 'TFLOP/s', 'frac': achieved / peak_measured,
		'peak_source': 'measured FFMA2/DFMA chain microbenchmark on this GPU (gravb200_peak_probe)',
		'peak_nominal': peak_nominal, 'frac_nominal': achieved / peak_nominal,
		'flop_per_interaction': FLOP_PER_INTERACTION,
		'kernel': 'symmetric (Newton 3rd law, every unordered pair once)' if symmetric else 'ordered (every ordered pair)',
		'executed_flop_per_interaction': flop_exec, 'achieved_executed': achieved * flop_exec / FLOP_PER_INTERACTION,
		'pipe_lane_ops_per_interaction': lane_ops, 'pipe_ceiling_g_inter_s': pipe_ceiling,
		'frac_of_pipe_ceiling': (per_gpu_rate / 1e9) / pipe_ceiling,
		'kernel_ms': kernel_s * 1e3, 'kernel_share_of_step': float(np.sum(sweep_ms) / max(np.sum(step_ms), 1e-9)),
		'traffic': traffic,
		'algorithmic_hbm_bytes': int(n * 4 * esz + (n // world) * 4 * esz * 4),
		'note': 'tensor cores are not applicable (softened 1/r^3 is not a contraction); HBM traffic is negligible',
		}

	# ---- parity of what was measured (inline float64 numpy on a few of rank 0's rows; no oracle here)
	n_rows0 = a0.shape[0]
	rows = np.linspace(0, n_rows0 - 1, 64 if n <= (1 << 22) else 8).astype(np.int64)
	r64, m64 = r.astype(np.float64), m.astype(np.float64)
	parity = 0.0
	for i in rows:
		d = r64 - r64[i]
		d2 = (d * d).sum(1)
		d2[i] = np.inf
		ref = G_SI * (d * (m64 / (d2 * np.sqrt(d2)))[:, None]).sum(0)
		parity = max(parity, float(np.linalg.norm(a0[i].astype(np.float64) - ref) / np.linalg.norm(ref)))

	# ---- CPU baseline (N = 1 only, bounded sample) ---------------------------------------------------
	cpu = None
	if world == 1 and not args.no_cpu_baseline:
		try:
			res = cpu_reference_run(16, 8, 2, os.cpu_count() or 1)
			cpu = {'value': res['g_inter_s'], 'unit': UNIT, 'cores': res['threads'], 'kind': res['kind'], 'sample': res['sample']}
		except Exception as e:
			cpu = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'reference', 'sample': 'failed: %s' % str(e)[:200]}

	# ---- the reference's own GPU kernel (pc2) on this B200: reported baseline, N = 1 only ------------
	gpu_ref = None
	if world == 1 and not args.no_cpu_baseline and dtype == 'float32':
		try:
			from oracle import pc2_bench
			if pc2_bench.available(n, dtype):
				gpu_ref = pc2_bench.run(args.bodies, steps = 2, warmup = 1, dtype = dtype)
		except Exception as e:
			gpu_ref = {'error': str(e)[:300]}

	line = {
		'metric': METRIC if (args.bodies == 20 and dtype == 'float32') else 'G body-interactions/s at N=2^%d %s' % (args.bodies, args.dtype),
		'value': value, 'unit': UNIT,
		'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
		'ms_per_step': total_ms / args.steps,
		'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
		'dtype': args.dtype, 'data': 'synthetic',
		'config': {
			'workload': 'all-pairs step (stage 1 sweep + fused stage 2), uniform universe, N=2^%d %s' % (args.bodies, args.dtype),
			'n_bodies': n, 'parallelism': ('row-sharded x%d, %s' % (world, 'fused peer-store exchange over NVLink (CUDA IPC) + flag barrier' if info['exchange_mode'] == 1 else 'NCCL all-gather of positions')) if world > 1 else 'single GPU',
			'grid': info['grid'], 'threads': info['threads'], 'bodies_per_thread': info['bodies_per_thread'], 'tile': info['tile'],
			'l2': 'flushed between timed steps (256 MiB write); the 16 MiB position array is then re-read from L2 by design',
			},
		'per_gpu': {'g_inter_s': per_gpu_rate / 1e9, 'rows_rank0': int(rows_rank0), 'rows_even_share': -(-n // world), 'sweep_ms': float(np.mean(sweep_ms)), 'exchange_ms': float(np.mean(xchg_ms)),
			'sm_mhz_in_kernel': float(np.median(sm_mhz))},
		'wall_ms_per_step': wall_ms / args.steps,
		'clocks': clocks.summary(),
		'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
			'steps': e2e_steps, 'api': 'gravitation_b200.kernel.b200.universe: push_host_state() + step(), eager host mirrors (pinned)',
			'checksum': checksum},
		'gpu_launches': int(launches),
		'roofline': roofline,
		'parity': {'max_rel_err_vs_float64_oracle': parity, 'rows': int(len(rows)), 'tolerance': 1e-4 if dtype == 'float32' else 1e-11},
		'peak_probe': probe,
		}
	if cpu is not None:
		line['cpu_baseline'] = cpu
	if gpu_ref is not None:
		line['reference_gpu_kernel'] = gpu_ref
	_emit(line)
	return 0


def _shutdown():
	try:
		import torch.distributed as tdist
		if tdist.is_initialized():
			tdist.destroy_process_group()
	except Exception:
		pass


class _StdoutGuard:
	"""The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to
	stdout when NCCL_DEBUG is set on the box), so file descriptor 1 points at stderr while the bench runs
	and is restored for the result line."""

	def __enter__(self):
		sys.stdout.flush()
		self._saved = os.dup(1)
		os.dup2(2, 1)
		return self

	def __exit__(self, *exc):
		sys.stdout.flush()
		os.dup2(self._saved, 1)
		os.close(self._saved)


def _emit(line):
	"""print the result line on the REAL stdout (see _StdoutGuard)"""
	sys.stdout.flush()
	fd = getattr(_emit, 'fd', None)
	payload = (json.dumps(line) + '\n').encode()
	if fd is None:
		sys.stdout.write(payload.decode()); sys.stdout.flush()
	else:
		os.write(fd, payload)


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type = int, default = 1)
	ap.add_argument('--steps', type = int, default = 10)
	ap.add_argument('--warmup', type = int, default = 3)
	ap.add_argument('--impl', default = 'b200', choices = ('b200', 'reference'))
	ap.add_argument('--bodies', type = int, default = 20, help = 'log2 of the number of bodies (default: the north-star 2^20)')
	ap.add_argument('--dtype', default = 'f32', choices = ('f32', 'f64'))
	ap.add_argument('--no-cpu-baseline', action = 'store_true')
	ap.add_argument('--quick-e2e', action = 'store_true', help = 'one warm-up + one timed end-to-end step (very large N)')
	args = ap.parse_args()
	if args.warmup < 3:
		args.warmup = 3
	with _StdoutGuard() as guard:
		_emit.fd = guard._saved
		try:
			if args.impl == 'reference':
				return reference_arm(args)
			rc = own_arm(args)
			_shutdown()
			return rc
		finally:
			_emit.fd = None


if __name__ == '__main__':
	sys.exit(main())
