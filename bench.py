#!/usr/bin/env python3
# -*- coding: utf-8 -*-
"""bench.py — the driver's measurement contract for the hot path (all-pairs step, stage 1 + fused stage 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--bodies LOG2N] [--dtype f32|f64]
  torchrun ... bench.py --gpus N ...        (one rank per GPU)

Workload (BASELINE.json `metric`): synthetic uniform universe, N = 2^20 bodies, float32, the same N at
every GPU count (strong scaling: rows are sharded, positions exchanged once per step).  One "step" = one
sweep over all N*(N-1) ordered interactions + the fused velocity/position update (+ the exchange).

Printed (rank 0, ONE JSON line):
  value        G interactions/s, device-timed (CUDA events on the launching stream, summed over K steps, max
               over ranks), state resident in HBM
  e2e          the same metric through the reference-facing kernel module with HOST buffers (pinned): every
               step uploads the step's inputs and downloads its results inside the timed region; on several
               GPUs every rank moves its own rows only and the device exchange completes the positions
  roofline     against the non-tensor FMA peak (measured FFMA2 / DFMA chain on this GPU, nominal beside it; 20
               FLOP per interaction) — `frac` is that convention number;
               `frac_of_pipe_ceiling` says how busy the FP pipe is with the instructions actually executed and
               is the number to tune against
  parity       max relative error of the accelerations of the measured run against inline float64 numpy on
               >= 4096 rows sampled from EVERY rank's shard (>= 512 above 2^22 bodies)
  configs      short runs of the other BASELINE.json configurations that fit this GPU count (2^16 fp32 and
               2^18 fp64 on one GPU; 2^18 fp64 on several; 2^24 fp32 on eight), each with value, pipe fraction, parity
  cpu_baseline the reference's fastest CPU kernel (c4b, oracle/_ref/lib4.so) on this box's host cores, bounded
               sample; `baselines` = the SURVEY 8d matrix (c4b raw/full at 2^12..2^16, c1a, np2 = configs[0],
               the float64 oracle) and the reference's pc2 GPU kernel compiled for sm_100a at 2^12 / 2^16 / 2^20
`--impl reference` times c4b alone on the largest sample of the workload the step budget allows and prints the
same line shape, with the size it actually ran in `config`.  Nothing here reads /root/reference at run time."""

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'G body-interactions/s at N=2^20 fp32'
UNIT = 'G interactions/s'
FLOP_PER_INTERACTION = 20.0 # SURVEY.md section 8d convention
G_SI, T_STEP = 6.6740831e-11, 2.0e12
DTYPES = {'f32': 'float32', 'f64': 'float64'}
TOLERANCE = {'float32': 1e-4, 'float64': 1e-11}


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

	def __init__(self, index, period = 0.05):
		# long steps are sampled at 4 Hz, short ones at 20 Hz: enough samples in either case, and fewer NVML queries
		# next to the launches of a step (single multi-GPU steps showed millisecond outliers, cause not pinned down)
		self.period = period
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
			self._stop.wait(self.period)

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

def _cpu_subprocess(expr, threads, timeout = 1500):
	"""fresh process (OMP_NUM_THREADS must be set before libgomp loads, c4a.py:63); returns the JSON it prints"""
	env = dict(os.environ, OMP_NUM_THREADS = str(threads))
	code = (
		'import sys, json; sys.path.insert(0, %r)\n'
		'from oracle import cpu_bench\n'
		'print(json.dumps(%s))\n'
		) % (ROOT, expr)
	out = subprocess.run([sys.executable, '-c', code], env = env, capture_output = True, text = True, timeout = timeout)
	if out.returncode != 0:
		raise RuntimeError('cpu reference run failed: %s' % out.stderr[-2000:])
	return json.loads(out.stdout.strip().splitlines()[-1])


def cpu_reference_run(log2n, steps, warmup, threads):
	return _cpu_subprocess('cpu_bench.run(%d, %d, %d)' % (log2n, steps, warmup), threads)


def reference_arm(args):
	rank = int(os.environ.get('RANK', 0))
	if rank != 0:
		return 0
	threads = os.cpu_count() or 1
	# the largest sample of the 2^20 workload whose K + W steps end within ~2.5 minutes (c4b is O(N^2): the
	# rate barely depends on N, the step time does)
	budget = 150.0 / max(1, args.steps + args.warmup)
	log2n = args.bodies if args.bodies_given else int(_cpu_subprocess('cpu_bench.pick_log2n(%r, %d)' % (budget, 20), threads))
	res = cpu_reference_run(log2n, args.steps, args.warmup, threads)
	line = {
		'impl': 'reference',
		'metric': METRIC, 'value': res['g_inter_s'], 'unit': UNIT,
		'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
		'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong',
		'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
		'config': {
			'workload': 'all-pairs step (stage 1 + stage 2), uniform universe, fp32: N=2^%d bodies, a bounded sample of the '
				'N=2^20 workload (O(N^2) kernel: the rate is a proxy, the step time is not)' % log2n,
			'n_bodies': 1 << log2n, 'sampled_from_n_bodies': 1 << 20, 'same_config': log2n == 20,
			'parallelism': 'host threads (%d)' % res['threads'],
			},
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

class Env:
	"""rank / world / torch handles shared by the legs"""

	def __init__(self, args):
		import torch
		from gravitation_b200 import _shim, dist
		self.torch, self.shim, self.dist = torch, _shim, dist
		self.rank, self.world, self.local_rank = dist.init_process_group()
		if self.world != args.gpus and self.world == 1 and args.gpus > 1:
			raise SystemExit('--gpus %d needs torchrun with %d ranks (one process per GPU)' % (args.gpus, args.gpus))
		torch.cuda.set_device(self.local_rank)
		# L2 flush between timed steps: write a buffer larger than the 126 MB L2
		self._flush = torch.empty(256 << 20, dtype = torch.uint8, device = 'cuda')

	def flush_l2(self):
		self._flush.zero_()
		self.torch.cuda.synchronize()

	def barrier(self):
		if self.world > 1:
			self.dist.barrier()

	def sync(self):
		self.torch.cuda.synchronize()
		self.barrier()

	def max_over_ranks(self, x):
		return self.dist.max_over_ranks(x) if self.world > 1 else float(x)

	def sum_over_ranks(self, x):
		if self.world == 1:
			return float(x)
		t = self.torch.tensor([float(x)], dtype = self.torch.float64, device = 'cuda')
		self.torch.distributed.all_reduce(t)
		return float(t.item())


def sampled_parity(env, r, m, a_local, row0, rows_total):
	"""max relative error of `a_local` (this rank's rows) against float64 numpy, on rows spread evenly over
	EVERY rank's shard; the per-rank maxima are combined with a MAX all-reduce.  Inline numpy on purpose: the
	oracle package is test infrastructure and stays out of the measured arm."""
	n_local = a_local.shape[0]
	want = max(1, -(-rows_total // env.world))
	rows = np.unique(np.linspace(0, n_local - 1, min(want, n_local)).astype(np.int64)) if n_local > 0 else np.zeros(0, np.int64)
	r64, m64 = r.astype(np.float64), m.astype(np.float64)

	def one(i_local):
		i = row0 + int(i_local)
		d = r64 - r64[i]
		d2 = np.einsum('ij,ij->i', d, d)
		d2[i] = np.inf
		ref = G_SI * ((m64 / (d2 * np.sqrt(d2))) @ d)
		return float(np.linalg.norm(a_local[i_local].astype(np.float64) - ref) / np.linalg.norm(ref))

	workers = max(1, (os.cpu_count() or 1) // env.world)
	with ThreadPoolExecutor(max_workers = workers) as pool:
		errs = list(pool.map(one, rows))
	worst = env.max_over_ranks(max(errs) if errs else 0.0)
	count = int(round(env.sum_over_ranks(len(rows))))
	return worst, count


def device_leg(env, n, dtype, steps, warmup, parity_rows):
	"""device-resident arm through the C ABI: `steps` timed steps after `warmup`, L2 flushed before each"""
	r, v, m = uniform_universe(n, 1000 + int(np.log2(n)), dtype)
	shard = env.dist.make_shard(n, dtype)
	shard.upload(r, v, m, G_SI, T_STEP)

	def one_step():
		env.flush_l2()
		env.barrier()
		shard.peer_barrier() # several GPUs: the timed events start when every rank's stream has got here (no host launch skew in the step)
		shard.stage1()
		shard.stage2()
		t = shard.timings()
		return t['sweep_ms'] + (t['exchange_ms'] if env.world > 1 else 0.0), t

	# warm-up step 1 doubles as the parity sample: its accelerations belong to the initial positions
	env.flush_l2()
	shard.stage1(); shard.sync()
	_, _, a0 = shard.download(r = False, v = False, a = True) # this rank's rows
	shard.stage2()
	for _ in range(warmup - 1):
		one_step()
	launches0 = shard.info()['launches']
	env.sync()
	wall0 = time.perf_counter()
	step_ms, sweep_ms, xchg_ms, sm_mhz, phases = [], [], [], [], []
	with ClockSampler(env.local_rank, period = 0.25 if n >= (1 << 18) else 0.05) as clocks:
		for _ in range(steps):
			ms, t = one_step()
			step_ms.append(ms); sweep_ms.append(t['sweep_ms']); xchg_ms.append(max(t['exchange_ms'], 0.0)); sm_mhz.append(t['sm_mhz'])
			if 'phases_ms' in t:
				phases.append(t['phases_ms'])
	env.sync()
	wall_ms = (time.perf_counter() - wall0) * 1e3
	info = shard.info()
	out = {
		'n': n, 'dtype': dtype, 'steps': steps, 'warmup': warmup,
		'total_ms': env.max_over_ranks(sum(step_ms)), 'wall_ms': wall_ms,
		'step_ms': step_ms, 'sweep_ms': sweep_ms, 'xchg_ms': xchg_ms, 'sm_mhz': sm_mhz,
		'launches': info['launches'] - launches0, 'info': info, 'rows': shard.n_local, 'row0': shard.row0,
		'clocks': clocks.summary(),
		}
	if phases:
		out['phases_ms'] = {key: float(np.mean([p[key] for p in phases])) for key in phases[0]}
	shard.close()
	out['value'] = float(n) * float(n - 1) * steps / (out['total_ms'] * 1e-3) / 1e9
	out['parity'], out['parity_rows'] = sampled_parity(env, r, m, a0, out['row0'], parity_rows)
	out['state'] = (r, v, m)
	return out


def e2e_leg(env, n, dtype, state, steps, warmup):
	"""end to end through the kernel module: host buffers in, host buffers out, every step"""
	from gravitation_b200.kernel import b200
	r, v, m = state
	kw = dict(T = T_STEP, G = G_SI, scale_off = True, dtype = dtype, eager_host = True, device = env.local_rank)
	if env.world > 1:
		uid = env.dist.broadcast_bytes(env.shim.nccl_unique_id() if env.rank == 0 else None)
		kw.update(rank = env.rank, world = env.world, nccl_id = uid, host_rows = 'own')
	u = b200.universe(**kw)
	u.add_objects(r, v, m, scale_off = True)
	u.start()
	esz = np.dtype(dtype).itemsize
	n_local = u._shards[0].n_local
	if env.world > 1: # own rows only; the device exchange (NVLink) completes the positions on every shard
		h2d = 2 * n_local * 3 * esz # r, v of this rank's rows
		d2h = 3 * n_local * 3 * esz # a after stage 1; r, v after stage 2
	else: # masses ride along with the full re-upload on one GPU
		h2d = (n * 3 + n * 3 + n) * esz
		d2h = 3 * n * 3 * esz

	def step():
		u.push_host_state() # H2D of the step's inputs from the pinned host mirrors
		u.step()            # stage 1 (+ D2H of a), stage 2 (+ D2H of r, v): eager_host = True
	for _ in range(warmup):
		step()
	env.sync()
	t0 = time.perf_counter()
	for _ in range(steps):
		step()
	env.sync()
	seconds = env.max_over_ranks(time.perf_counter() - t0)
	rows = slice(u._shards[0].row0, u._shards[0].row0 + n_local)
	checksum = env.sum_over_ranks(float(np.abs(u.mass_r_array[rows]).sum(dtype = np.float64)))
	fallback = getattr(u, 'exchange_fallback', None)
	u.stop()
	return {
		'value': float(n) * float(n - 1) * steps / seconds / 1e9, 'unit': UNIT,
		'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h), 'bytes_are': 'per rank' if env.world > 1 else 'total',
		'steps': steps, 'warmup': warmup,
		'api': 'gravitation_b200.kernel.b200.universe: push_host_state() + step(), eager host mirrors (pinned)'
			+ ("; host_rows='own': every rank moves its own rows" if env.world > 1 else ''),
		'checksum': checksum, 'exchange_fallback': fallback,
		}


def pipe_model(info, dtype, shim):
	"""executed work per ORDERED interaction (the metric's unit): ordered sweep 12 FP32-pipe lane-ops = 19 FLOP,
	symmetric sweep (every unordered pair once, both bodies updated) 8 lane-ops = 13 FLOP; fp64: 16 / 10 ops"""
	symmetric = shim.SYM_BASE <= info.get('variant', 0) < shim.SMALL_BASE
	lane_ops = (8.0 if symmetric else 12.0) if dtype == 'float32' else (10.0 if symmetric else 16.0)
	flop_exec = (13.0 if symmetric else 19.0) if dtype == 'float32' else (16.0 if symmetric else 25.0)
	return symmetric, lane_ops, flop_exec


def pipe_fraction(run, world, sm_max, shim):
	symmetric, lane_ops, _ = pipe_model(run['info'], run['dtype'], shim)
	lanes = 128 if run['dtype'] == 'float32' else 64
	ceiling = run['info']['sm_count'] * lanes * sm_max * 1e6 / lane_ops / 1e9 # G interactions/s per GPU, FP pipe never idle
	return (run['value'] / world) / ceiling, ceiling, symmetric


def config_entry(env, name, log2n, dtype, steps, warmup, parity_rows, sm_max):
	run = device_leg(env, 1 << log2n, dtype, steps, warmup, parity_rows)
	frac, ceiling, symmetric = pipe_fraction(run, env.world, sm_max, env.shim)
	return {
		'config': name, 'n_bodies': run['n'], 'dtype': dtype, 'n_gpus': env.world,
		'value': run['value'], 'unit': UNIT, 'ms_per_step': run['total_ms'] / steps, 'steps': steps, 'warmup': warmup,
		'frac_of_pipe_ceiling': frac, 'pipe_ceiling_g_inter_s': ceiling,
		'kernel': 'symmetric' if symmetric else 'ordered', 'variant': run['info']['variant'], 'chunk_granular_cta_ranges': bool(run['info'].get('split', 0)), 'grid': run['info']['grid'],
		'parity': {'max_rel_err_vs_float64': run['parity'], 'rows': run['parity_rows'], 'tolerance': TOLERANCE[dtype],
			'ok': bool(run['parity'] <= TOLERANCE[dtype])},
		'clocks': run['clocks'],
		}


def own_arm(args):
	env = Env(args)
	shim, world, rank = env.shim, env.world, env.rank
	dtype = DTYPES[args.dtype]
	n = 1 << args.bodies
	interactions = float(n) * float(n - 1)
	esz = np.dtype(dtype).itemsize
	parity_rows = 4096 if args.bodies <= 22 else 512

	main = device_leg(env, n, dtype, args.steps, args.warmup, parity_rows)
	e2e = e2e_leg(env, n, dtype, main.pop('state'), 1 if args.quick_e2e else args.steps, 1 if args.quick_e2e else min(args.warmup, 3))
	sm_max = main['clocks']['sm_max_mhz'] or 1965

	# the other BASELINE.json configurations that fit this GPU count, as short runs
	configs = []
	if not args.no_configs and args.bodies == 20 and dtype == 'float32':
		todo = []
		if world == 1:
			todo.append(('configs[1]: 2^16 bodies fp32 on 1xB200', 16, 'float32', 50, 5))
		todo.append(('configs[3]: 2^18 bodies fp64 on %dxB200' % world, 18, 'float64', 10, 3))
		if world == 8:
			todo.append(('configs[4]: 2^24 bodies fp32 on 8xB200 (one timed step)', 24, 'float32', 1, 1))
		for name, lg, dt, st, wu in todo:
			try:
				entry = config_entry(env, name, lg, dt, st, wu, 4096 if lg <= 22 else 512, sm_max)
			except Exception as e: # a failed side run must not lose the headline
				entry = {'config': name, 'error': str(e)[:300]}
			configs.append(entry)

	if rank != 0:
		return 0

	# ---- roofline: non-tensor FMA pipe --------------------------------------------------------------
	info = main['info']
	probe = shim.peak_probe(env.local_rank)
	lanes = 128 if dtype == 'float32' else 64
	peak_nominal = info['sm_count'] * lanes * 2 * sm_max * 1e6 / 1e12
	peak_measured = max(probe['fp32x2_tflops'], probe['fp32_tflops']) if dtype == 'float32' else probe['fp64_tflops']
	per_gpu_rate = main['value'] / world
	kernel_s = float(np.mean(main['sweep_ms'])) * 1e-3
	achieved = (interactions / world) * FLOP_PER_INTERACTION / kernel_s / 1e12 # dominant kernel, per GPU
	traffic, traffic_source = None, None
	tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
	if os.path.isfile(tpath):
		try:
			tj = json.load(open(tpath))
			traffic = tj.get('%s_2p%d' % (args.dtype, args.bodies))
			if traffic is not None:
				traffic_source = 'profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch on one GPU: %s); not measured in this run' % tj.get('source', {}).get('%s_2p%d' % (args.dtype, args.bodies), 'see profiles/README.md')
		except Exception:
			traffic = None
	symmetric, lane_ops, flop_exec = pipe_model(info, dtype, shim)
	frac_pipe, pipe_ceiling, _ = pipe_fraction(main, world, sm_max, shim)
	roofline = {
		'bound': 'fp32_fma' if dtype == 'float32' else 'fp64_fma',
		'achieved': achieved, 'peak': peak_measured, 'unit': 'TFLOP/s', 'frac': achieved / peak_measured,
		'frac_is': 'the 20-FLOP-per-interaction convention of SURVEY 8d; it credits work the symmetric sweep never executes — tune against frac_of_pipe_ceiling',
		'frac_of_pipe_ceiling': frac_pipe, 'pipe_ceiling_g_inter_s': pipe_ceiling, 'pipe_lane_ops_per_interaction': lane_ops,
		'peak_source': 'measured FFMA2/DFMA chain microbenchmark on this GPU (gravb200_peak_probe)',
		'peak_nominal': peak_nominal, 'frac_nominal': achieved / peak_nominal,
		'flop_per_interaction': FLOP_PER_INTERACTION,
		'kernel': 'symmetric (Newton 3rd law, every unordered pair once)' if symmetric else 'ordered (every ordered pair)',
		'executed_flop_per_interaction': flop_exec, 'achieved_executed': achieved * flop_exec / FLOP_PER_INTERACTION,
		'kernel_ms': kernel_s * 1e3, 'kernel_share_of_step': float(np.sum(main['sweep_ms']) / max(np.sum(main['step_ms']), 1e-9)),
		'traffic': traffic, 'traffic_source': traffic_source,
		'algorithmic_hbm_bytes': int(n * 4 * esz + (n // world) * 4 * esz * 4),
		'note': 'tensor cores are not applicable (the 1/r^3 interaction is not a contraction); HBM traffic is negligible',
		}

	# ---- CPU baselines (N = 1 only, bounded samples) -------------------------------------------------
	cpu, baselines = None, None
	if world == 1 and not args.no_cpu_baseline:
		threads = os.cpu_count() or 1
		try:
			res = cpu_reference_run(16, 8, 2, threads)
			cpu = {'value': res['g_inter_s'], 'unit': UNIT, 'cores': res['threads'], 'kind': res['kind'], 'sample': res['sample']}
		except Exception as e:
			cpu = {'value': None, 'unit': UNIT, 'cores': threads, 'kind': 'reference', 'sample': 'failed: %s' % str(e)[:200]}
		try:
			baselines = {'cpu': _cpu_subprocess('cpu_bench.matrix(40.0)', threads, timeout = 600)}
		except Exception as e:
			baselines = {'cpu': {'error': str(e)[:300]}}
		# the reference's own GPU kernel (pc2) on this B200
		gpu_rows = []
		if dtype == 'float32':
			try:
				from oracle import pc2_bench
				for lg in (12, 16, 20):
					if pc2_bench.available(1 << lg, 'float32'):
						gpu_rows.append(pc2_bench.run(lg, steps = 3 if lg < 20 else 2, warmup = 1, dtype = 'float32'))
			except Exception as e:
				gpu_rows.append({'error': str(e)[:300]})
		baselines['reference_gpu_kernel_pc2'] = gpu_rows

	exchange = None
	if world > 1:
		if info['exchange_mode'] != 1:
			exchange = 'NCCL all-gather of positions'
		elif symmetric:
			exchange = ('stream-K shares of the universe\'s tile list per GPU; reduce-scatter of the partial sums + peer stores of r\' inside the '
				'integrate kernel over NVLink (CUDA IPC), hand-over flags folded into the sweep / integrate kernels (no barrier launches)')
		else:
			exchange = 'fused peer-store exchange over NVLink (CUDA IPC) in the sweep epilogue + flag barrier'
	line = {
		'metric': METRIC if (args.bodies == 20 and dtype == 'float32') else 'G body-interactions/s at N=2^%d %s' % (args.bodies, args.dtype),
		'value': main['value'], 'unit': UNIT,
		'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
		'ms_per_step': main['total_ms'] / args.steps,
		'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
		'dtype': args.dtype, 'data': 'synthetic',
		'config': {
			'workload': 'all-pairs step (stage 1 sweep + fused stage 2), uniform universe, N=2^%d %s' % (args.bodies, args.dtype),
			'n_bodies': n, 'parallelism': ('row-sharded x%d, %s' % (world, exchange)) if world > 1 else 'single GPU',
			'exchange': exchange, 'exchange_fallback': e2e.get('exchange_fallback'),
			'grid': info['grid'], 'threads': info['threads'], 'bodies_per_thread': info['bodies_per_thread'], 'tile': info['tile'],
			'variant': info['variant'], 'chunk_granular_cta_ranges': bool(info.get('split', 0)),
			'l2': 'flushed between timed steps (256 MiB write); the position array is then re-read from L2 by design',
			'start': 'every timed step starts behind a device-side flag barrier of all ranks (host launch skew is not part of the step)' if world > 1 else 'single stream',
			},
		'per_gpu': {'g_inter_s': per_gpu_rate, 'rows_rank0': int(main['rows']), 'rows_even_share': -(-n // world),
			'sweep_ms': float(np.mean(main['sweep_ms'])), 'exchange_ms': float(np.mean(main['xchg_ms'])),
			'step_ms_rank0': {'min': float(np.min(main['step_ms'])), 'median': float(np.median(main['step_ms'])), 'max': float(np.max(main['step_ms']))},
			'sm_mhz_in_kernel': float(np.median(main['sm_mhz']))},
		'wall_ms_per_step': main['wall_ms'] / args.steps,
		'clocks': main['clocks'],
		'e2e': e2e,
		'gpu_launches': int(main['launches']),
		'roofline': roofline,
		'parity': {'max_rel_err_vs_float64_oracle': main['parity'], 'rows': main['parity_rows'], 'rows_from': 'every rank\'s shard',
			'tolerance': TOLERANCE[dtype], 'ok': bool(main['parity'] <= TOLERANCE[dtype])},
		'peak_probe': probe,
		}
	if 'phases_ms' in main:
		line['per_gpu']['phases_ms_rank0'] = main['phases_ms']
	if configs:
		line['configs'] = configs
	if cpu is not None:
		line['cpu_baseline'] = cpu
	if baselines is not None:
		line['baselines'] = baselines
		pc2 = [g for g in baselines.get('reference_gpu_kernel_pc2', []) if g.get('n') == n]
		if pc2:
			line['reference_gpu_kernel'] = pc2[0]
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
	ap.add_argument('--bodies', type = int, default = None, help = 'log2 of the number of bodies (default: the north-star 2^20)')
	ap.add_argument('--dtype', default = 'f32', choices = ('f32', 'f64'))
	ap.add_argument('--no-cpu-baseline', action = 'store_true')
	ap.add_argument('--no-configs', action = 'store_true', help = 'skip the short runs of the other BASELINE.json configurations')
	ap.add_argument('--quick-e2e', action = 'store_true', help = 'one warm-up + one timed end-to-end step (very large N)')
	args = ap.parse_args()
	args.bodies_given = args.bodies is not None
	if args.bodies is None:
		args.bodies = 20
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
