# -*- coding: utf-8 -*-
"""CPU tests of the host side: front-end contract (mirrors /root/reference/src/gravitation/kernel/_base_.py),
inventory, scenario builders against golden universes, and the C-ABI library's symbol table."""
import os
import re

import numpy as np
import pytest

from gravitation_b200.kernel._base_ import universe_base, _point_mass
from gravitation_b200.lib import load, simulation, timing

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _noop(universe_base):
	def step_stage1(self):
		for pm in self._mass_list:
			pm._a[:] = [1.0, 0.0, 0.0]


# ---- universe_base contract (_base_.py:64-177) ---------------------------------------------------

def test_constructor_scales_G_unless_scale_off():
	u = _noop(G = 2.0, scale_m = 4.0, scale_r = 3.0)
	assert u._G == 2.0 * 27.0 / 4.0
	u = _noop(G = 2.0, scale_m = 4.0, scale_r = 3.0, scale_off = True, extra = 7)
	assert u._G == 2.0 and u._meta == {'extra': 7}
	assert (u._dtype, u._threads, u._t, u._T) == ('float32', 1, 0.0, 1.0e3)


def test_add_object_scales_in_place_on_callers_lists():
	u = _noop(scale_m = 0.5, scale_r = 10.0)
	r, v = [1.0, 2.0, 3.0], [0.1, 0.2, 0.3]
	u.add_object(name = 'a', r = r, v = v, m = 8.0)
	assert r == [10.0, 20.0, 30.0] and v == [1.0, 2.0, 3.0] # caller's lists were modified (_base_.py:114-116)
	pm = next(iter(u))
	assert pm._m == 4.0 and pm._r is r and pm._a == [0.0, 0.0, 0.0] and len(u) == 1
	u.add_object(name = 'b', r = [1.0, 1.0, 1.0], v = [0.0, 0.0, 0.0], m = 1.0, scale_off = True)
	assert list(u)[1]._r == [1.0, 1.0, 1.0]


def test_lifecycle_errors_have_the_reference_messages():
	u = _noop()
	with pytest.raises(SyntaxError, match = 'simulation was not started'):
		u.step()
	with pytest.raises(SyntaxError, match = 'simulation was not started'):
		u.stop()
	u.add_object(name = 'a', r = [0.0, 0.0, 0.0], v = [0.0, 0.0, 0.0], m = 1.0)
	u.start()
	with pytest.raises(SyntaxError, match = 'simulation is running'):
		u.start()
	with pytest.raises(SyntaxError, match = 'simulation was started'):
		u.add_object(name = 'b', r = [0.0] * 3, v = [0.0] * 3, m = 1.0)
	u.step()
	u.stop()
	for call, msg in ((u.step, 'simulation was stopped'), (u.start, 'simulation was stopped'), (u.stop, 'simulation was stopped before')):
		with pytest.raises(SyntaxError, match = msg):
			call()
	with pytest.raises(SyntaxError, match = 'simulation was stopped'):
		u.add_object(name = 'b', r = [0.0] * 3, v = [0.0] * 3, m = 1.0)


def test_step_runs_three_stages_and_default_stage2_is_symplectic_euler():
	u = _noop(T = 2.0)
	u.add_object(name = 'a', r = [0.0, 0.0, 0.0], v = [1.0, 0.0, 0.0], m = 1.0)
	u.start()
	u.step()
	pm = list(u)[0]
	assert pm._v == [3.0, 0.0, 0.0] and pm._r == [6.0, 0.0, 0.0] and pm._a == [0.0, 0.0, 0.0] # new v moves r
	assert u._t == 2.0
	with pytest.raises(NotImplementedError):
		universe_base().step_stage1()
	assert str(pm).startswith('a | 6.0000e+00, 0.0000e+00')


# ---- inventory (lib/load.py:40-107) --------------------------------------------------------------

def test_inventory_lists_non_underscore_kernels_and_reads_meta_without_import():
	assert 'b200' in load.inventory and not any(name.startswith('_') for name in load.inventory)
	k = load.inventory['b200']
	with pytest.raises(SyntaxError, match = 'metadata has not been loaded'):
		k['parallel']
	k.load_meta()
	assert k['parallel'] is True and k['name'] == 'b200' and k['interpreters'] == ['python3']
	assert set(load.META_KEYS) <= set(k.keys())
	fresh = load._kernel(k._path, 'b200', True)
	with pytest.raises(SyntaxError, match = 'module has not been loaded'):
		fresh.get_class()
	k.load_module()
	assert k.get_class().__name__ == 'universe'


def test_read_meta_ignores_non_name_targets():
	meta = load.read_meta("__version__ = '1'\na.b = 3\nx, y = 1, 2\n__parallel__ = False\n")
	assert meta['version'] == '1' and meta['parallel'] is False and meta['longname'] is None


# ---- scenario builders (lib/simulation.py:41-184) ------------------------------------------------

class _recorder(universe_base):
	def step_stage1(self):
		pass


def _state(u):
	r = np.array([list(pm._r) for pm in u]); v = np.array([list(pm._v) for pm in u]); m = np.array([pm._m for pm in u])
	return r, v, m


@pytest.mark.parametrize('case,n', (('galaxy256', 256), ('galaxy4096', 4096)))
def test_seeded_galaxy_equals_the_reference_universe(case, n, golden):
	g = golden[case]
	u = simulation.create_simulation('galaxy', _recorder, {'stars_len': n, 'seed': 42, 'dtype': 'float64'})
	r, v, m = _state(u)
	assert len(u) == n and list(u)[0]._name == 'back hole' and list(u)[1]._name == 'star'
	assert np.array_equal(r, g['r0']) and np.array_equal(v, g['v0']) and np.array_equal(m, g['m'])
	assert u._G == float(g['G']) and u._T == 2.0e12 and u._screen['unit'] == 1e20


def test_unseeded_galaxy_uses_the_global_random_stream_like_the_reference(golden):
	import random
	random.seed(42)
	u = simulation.create_simulation('galaxy', _recorder, {'stars_len': 256, 'dtype': 'float64'})
	assert np.array_equal(_state(u)[0], golden['galaxy256']['r0'])


def test_solarsystem_and_unknown_scenario(golden):
	u = simulation.create_simulation('solarsystem', _recorder, {'dtype': 'float64'})
	r, v, m = _state(u)
	g = golden['solarsystem']
	assert np.array_equal(r, g['r0']) and np.array_equal(v, g['v0']) and np.array_equal(m, g['m'])
	with pytest.raises(ValueError, match = 'Unknown scenario'):
		simulation.create_simulation('nope', _recorder)


class _bulk_recorder(universe_base):
	def add_objects(self, r, v, m, names = None, scale_off = False):
		self.bulk = (np.array(r) * self._scale_r, np.array(v) * self._scale_r, np.array(m) * self._scale_m, names)
	def step_stage1(self):
		pass


def test_vectorised_galaxy_builder_has_the_reference_construction():
	"""above 2^16 bodies the galaxy comes from the numpy restatement of `simulation.py:114-184`
	(own seeded stream, not bit-identical).  Checked: body 0 is the central mass, 80 / 20 split, radius ranges,
	z-jitter envelope, circular-orbit speed from the universe's G, velocity at a right angle to the radius"""
	n = 65536 + 8
	u = simulation.create_simulation('galaxy', _bulk_recorder, {'stars_len': n, 'seed': 7})
	R, V, M, names = u.bulk
	assert R.shape == (n, 3) and len(names) == n and names[0] == 'back hole' and names[5] == 'star'
	assert M[0] == 4e40 * 1e-30 and np.all(M[1:] == 2e30 * 1e-30) and not R[0].any() and not V[0].any()
	rad = 1e20 * 1e-10
	stars, n_disc = n - 1, (n - 1) * 4 // 5
	rho = np.linalg.norm(R[1:], axis = 1)
	disc_rho = np.linalg.norm(R[1:1 + n_disc, :2], axis = 1)
	assert disc_rho.min() >= 0.1 * rad and disc_rho.max() < 4.6 * rad and disc_rho.max() > 4.5 * rad
	assert np.all(np.abs(R[1:1 + n_disc, 2]) <= 0.25 * rad * (4.6 * rad - disc_rho) / (4.6 * rad) * (1 + 1e-12))
	assert rho[n_disc:].min() >= 0.1 * rad * (1 - 1e-12) and rho[n_disc:].max() < 0.85 * rad
	assert np.abs(R[1 + n_disc:, 2]).max() > 0.5 * rad # the bulge is three-dimensional
	# orbit speed as the reference computes it: the universe's scaled G with UNSCALED lengths (`simulation.py:148`)
	speed = np.linalg.norm(V[1:], axis = 1)
	assert np.allclose(speed, np.sqrt(u._G * 4e40 / (rho / 1e-10)) * 1e-10, rtol = 1e-12)
	assert np.abs((R[1:, :2] * V[1:, :2]).sum(1)).max() < 1e-9 * rad * speed.max() and not V[1:, 2].any()
	again = simulation.create_simulation('galaxy', _bulk_recorder, {'stars_len': n, 'seed': 7}).bulk
	other = simulation.create_simulation('galaxy', _bulk_recorder, {'stars_len': n, 'seed': 8}).bulk
	assert np.array_equal(again[0], R) and not np.array_equal(other[0], R)
	# below the switch the reference's own stream is kept (bit-identical universes); either builder can be forced
	small = simulation.create_simulation('galaxy', _bulk_recorder, {'stars_len': 4200, 'seed': 42})
	ref = simulation.create_simulation('galaxy', _recorder, {'stars_len': 4200, 'seed': 42})
	assert np.array_equal(small.bulk[0], _state(ref)[0])
	forced = simulation.create_simulation('galaxy', _bulk_recorder, {'stars_len': 4200, 'seed': 42, 'builder': 'vector'})
	assert not np.array_equal(forced.bulk[0], small.bulk[0])
	with pytest.raises(ValueError, match = 'Unknown galaxy builder'):
		simulation.create_simulation('galaxy', _bulk_recorder, {'stars_len': 64, 'builder': 'nope'})
	# tilt and turn act on positions and velocities alike (rotation about x, then about z)
	R0, V0, _ = simulation.galaxy_arrays(1000, 1.0, [0.0] * 3, [0.0] * 3, 0.0, 0.0, 10.0, 1.0, 1.0, seed = 1)
	R1, V1, _ = simulation.galaxy_arrays(1000, 1.0, [1.0, 2.0, 3.0], [4.0, 5.0, 6.0], 0.3, 0.2, 10.0, 1.0, 1.0, seed = 1)
	cb, sb, ca, sa = np.cos(0.2), np.sin(0.2), np.cos(0.3), np.sin(0.3)
	rot = np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1.0]]) @ np.array([[1.0, 0, 0], [0, cb, -sb], [0, sb, cb]])
	assert np.allclose(R1[1:], R0[1:] @ rot.T + [1.0, 2.0, 3.0]) and np.allclose(V1[1:], V0[1:] @ rot.T + [4.0, 5.0, 6.0])


def test_snapshot_roundtrip(tmp_path):
	u = simulation.create_simulation('galaxy', _recorder, {'stars_len': 32, 'seed': 1})
	fn = str(tmp_path / 'data.npz') # the .npz twin, whatever is installed
	path = simulation.store_simulation(u, fn, 'kernel=x;len=32;step=0')
	assert path == fn
	u2 = simulation.load_simulation(_recorder, path, 'kernel=x;len=32;step=0')
	r, v, m = _state(u); r2, v2, m2 = _state(u2)
	assert np.array_equal(r.astype('f4'), r2.astype('f4')) and np.array_equal(m.astype('f4'), m2.astype('f4'))
	assert u2._G == u._G and u2._T == u._T and u2._dtype == 'float32' and list(u2)[0]._name == 'back hole'
	with pytest.raises(FileNotFoundError):
		simulation.load_simulation(_recorder, str(tmp_path / 'nothing.h5'), 'x')


def _hdf5_roundtrip(tmp_path, h5py):
	"""store -> inspect the file with `h5py` -> load; the layout is the reference's (`simulation.py:216-258`)"""
	u = simulation.create_simulation('galaxy', _recorder, {'stars_len': 32, 'seed': 1})
	fn = str(tmp_path / 'data.h5')
	groups = ['kernel=x;len=32;step=0', 'kernel=x;len=32;step=5']
	assert simulation.store_simulation(u, fn, groups[0]) == fn # what the writer returns is what the reader takes
	u._t = 5 * u._T
	assert simulation.store_simulation(u, fn, groups[1]) == fn # append mode: a second group in the same file
	with h5py.File(fn, 'r') as f:
		assert sorted(f.keys()) == sorted(groups)
		dg = f[groups[1]]
		assert dg['r'].shape == (32, 3) and dg['v'].shape == (32, 3) and dg['m'].shape == (32,) and dg['name'].shape == (32,)
		assert np.dtype(dg['r'].dtype) == np.dtype('<f4') and np.dtype(dg['m'].dtype) == np.dtype('<f4')
		assert np.dtype(dg['name'].dtype) == np.dtype('S9') and bytes(dg['name'][0]) == b'back hole'
		assert sorted(dg.attrs.keys()) == sorted(['scale_m', 'scale_r', 't', 'T', 'G', 'dtype', 'threads'])
	u2 = simulation.load_simulation(_recorder, fn, groups[1], threads = 3)
	r, v, m = _state(u); r2, v2, m2 = _state(u2)
	assert np.array_equal(r.astype('f4'), r2.astype('f4')) and np.array_equal(v.astype('f4'), v2.astype('f4'))
	assert np.array_equal(m.astype('f4'), m2.astype('f4'))
	assert u2._G == u._G and u2._T == u._T and u2._t == u._t and u2._dtype == 'float32' and u2._threads == 3
	assert type(u2._dtype) is str and type(u2._T) is float
	assert [pm._name for pm in u2][:2] == ['back hole', 'star']
	assert simulation.load_simulation(_recorder, fn, groups[0])._t == 0.0


def test_snapshot_hdf5_branch_with_a_stand_in_h5py(tmp_path, monkeypatch):
	"""h5py is absent from this image: the HDF5 branches of store/load run against tests/fake_h5py.py"""
	import sys
	sys.path.insert(0, os.path.join(ROOT, 'tests'))
	import fake_h5py
	monkeypatch.setitem(sys.modules, 'h5py', fake_h5py)
	_hdf5_roundtrip(tmp_path, fake_h5py)
	# an HDF5 file without h5py: a clear error instead of "data.h5.npz not found"
	monkeypatch.setitem(sys.modules, 'h5py', None)
	with pytest.raises(OSError, match = 'h5py is not importable'):
		simulation.load_simulation(_recorder, str(tmp_path / 'data.h5'), 'kernel=x;len=32;step=0')


def test_snapshot_hdf5_layout_with_real_h5py(tmp_path):
	h5py = pytest.importorskip('h5py')
	_hdf5_roundtrip(tmp_path, h5py)


def test_pinned_views_keep_their_memory_alive(monkeypatch):
	"""a row view that outlives its universe must not point at freed page-locked memory.
	The allocation belongs to the numpy base chain; it is released when the LAST view goes."""
	import ctypes
	import gc
	from gravitation_b200 import _shim
	libc = ctypes.CDLL(None)
	libc.malloc.restype = ctypes.c_void_p
	libc.malloc.argtypes = [ctypes.c_size_t]
	libc.free.argtypes = [ctypes.c_void_p]
	freed = []

	class _fake_lib:
		@staticmethod
		def gravb200_host_alloc(nbytes, out):
			out._obj.value = libc.malloc(nbytes)
			return 0
		@staticmethod
		def gravb200_host_free(p):
			freed.append(p.value)
			libc.free(p)
			return 0
	monkeypatch.setattr(_shim, 'load', lambda: _fake_lib)
	holder = _shim.PinnedArray((7, 3), 'float64')
	assert holder.array.shape == (7, 3) and holder.array.dtype == np.float64 and not holder.array.any()
	holder.array[3, :] = (1.0, 2.0, 3.0)
	row = holder.array[3, :]
	del holder
	gc.collect()
	assert freed == [] and list(row) == [1.0, 2.0, 3.0] # still valid, still owned
	del row
	gc.collect()
	assert len(freed) == 1


def test_timers():
	t = timing.best_run_timer()
	for _ in range(3):
		t.start(); t.stop()
	assert len(t) == 3 and t.min() <= t.avg() <= t.sum() and timing.elapsed_timer()() >= 0


# ---- C-ABI library: loads, exports everything the header declares, refuses to compute without a GPU

def test_library_exports_every_symbol_of_the_header(shim):
	header = open(os.path.join(ROOT, 'include', 'gravb200.h')).read()
	declared = set(re.findall(r'\b(gravb200_[a-z0-9_]+)\s*\(', header))
	assert declared == set(shim.SYMBOLS)
	lib = shim.load()
	for name in declared:
		assert hasattr(lib, name)
	assert lib.gravb200_abi_version() == 1
	assert lib.gravb200_variant_count(shim.F32) >= 4 and lib.gravb200_variant_count(shim.F64) >= 3


def test_no_cpu_fallback(shim):
	if shim.device_count() > 0:
		pytest.skip('a GPU is present')
	with pytest.raises(shim.GravB200Error, match = 'no CUDA device'):
		shim.Shard(16)
	from gravitation_b200.kernel import b200
	u = b200.universe()
	u.add_object(name = 'a', r = [0.0] * 3, v = [0.0] * 3, m = 1.0)
	with pytest.raises(shim.GravB200Error):
		u.start()


def test_product_code_never_touches_the_oracle():
	for base, _, files in os.walk(os.path.join(ROOT, 'gravitation_b200')):
		for fn in files:
			if fn.endswith(('.py', '.cu', '.cuh', '.h')):
				src = open(os.path.join(base, fn), errors = 'replace').read()
				assert not re.search(r'^\s*(from|import)\s+oracle|liboracle|oracle/_ref', src, re.M), fn


def test_row_partition_covers_all_rows(shim):
	from gravitation_b200.dist import row_partition
	for dtype in ('float32', 'float64'):
		for n, world in ((1 << 20, 8), (1 << 20, 2), (1 << 18, 8), (1 << 24, 8), (5, 8), (1000, 3), (7, 1), (40000, 2), (100003, 4), (32768, 8), (33000, 8), (43116, 4), (178225, 4), (1 << 20, 3)):
			parts = row_partition(n, world, dtype)
			assert len(parts) == world and sum(c for _, c in parts) == n
			pos = 0
			for row0, cnt in parts:
				if cnt:
					assert row0 == pos
				pos += cnt
			# ceil(n / world) rows per shard, the last one short (the symmetric sweep divides its WORK
			# independently of the rows a shard owns, so nothing is rounded to body-blocks any more)
			chunk = -(-n // world)
			assert parts == [(min(k * chunk, n), max(0, min(chunk, n - k * chunk))) for k in range(world)]
	assert row_partition(1 << 20, 8, 'float32') == [(131072 * k, 131072) for k in range(8)]
	assert row_partition(178225, 4, 'float32')[3] == (3 * 44557, 178225 - 3 * 44557)


def test_small_kernel_geometry(shim):
	"""persistent small-N kernel (csrc/nbody_small.cuh): the host-side geometry covers every row and every
	j-body exactly once, keeps at most 32 rows per row group, an even number of rows per CTA, whole unrolled
	trips per slice, and fits the 227 KiB of shared memory wherever it says so"""
	names = shim.small_variant_names()
	assert len(names) >= 2 and all(nm.startswith('small_t') for nm in names)
	for dtype, esz in (('float32', 4), ('float64', 8)):
		for k, name in enumerate(names):
			threads, unroll, rows = (int(name.split(key)[1].split('_')[0]) for key in ('_t', '_u', '_r'))
			for sms in (148, 132, 8):
				for n in (1, 2, 3, 16, 255, 256, 257, 1000, 4096, 4099, 6000, 8192, 9472, 9473, 14000, 16384):
					try:
						g = shim.small_geometry(n, dtype, sms, shim.SMALL_BASE + k)
					except shim.GravB200Error:
						assert -(-n // (32 * (threads // 32))) > sms # more rows per CTA than the variant has lanes for
						continue
					assert g['grid'] <= sms and g['grid'] * g['rows_per_cta'] >= n and (g['grid'] - 1) * g['rows_per_cta'] < n
					assert g['rows_per_cta'] % 2 == 0 and g['rows_per_cta'] <= 32 * g['row_groups']
					assert g['row_groups'] * g['slices'] == threads // 32
					assert g['slice'] % (rows * unroll) == 0 and g['slices'] * g['slice'] >= n
					assert g['smem_bytes'] >= 128 + g['slices'] * g['slice'] * 4 * esz
					assert g['fits'] == (g['smem_bytes'] <= 227 * 1024)
	# the automatic range on a B200: up to 64 rows per CTA on 148 SMs
	assert shim.small_geometry(9472, 'float32')['row_groups'] == 2 and shim.small_geometry(9473, 'float32')['row_groups'] == 4
	assert shim.small_geometry(4096, 'float32') == dict(grid = 128, rows_per_cta = 32, row_groups = 1, slice = 512, slices = 8,
		smem_bytes = 128 + 4096 * 16 + 8 * 3 * 32 * 8, fits = True)


# ---- worker protocol / analyze (cli/worker.py:115-245, cli/analyze.py:47-108) ----------------------

def _fake_log(steps = 3, exit_ok = True, extra = ''):
	import json
	lines = ['{"log": "START"}', json.dumps({'log': 'INPUT', 'simulation': {'kernel': 'b200', 'threads': 1}}),
		json.dumps({'log': 'PROCEDURE', 'msg': 'x'}), json.dumps({'log': 'SIZE', 'value': 16})]
	for k in range(1, steps + 1):
		lines += [json.dumps({'log': 'STEP', 'runtime': 100 + k, 'gctime': 5, 'counter': k}), json.dumps({'log': 'BEST_TIME', 'value': 101})]
	lines += [json.dumps({'log': 'RATE', 'best_interactions_per_s': 1.0})] # our extra line type is ignored
	if extra:
		lines.append(extra)
	if exit_ok:
		lines.append(json.dumps({'log': 'EXIT', 'msg': 'OK'}))
	return '\n'.join(lines) + '\n'


def test_analyze_accepts_good_logs_and_rejects_bad_ones():
	from gravitation_b200.cli import analyze
	runs = analyze.parse_log(_fake_log() + _fake_log(steps = 2))
	assert len(runs) == 2 and runs[0]['runtime'] == [101, 102, 103] and runs[0]['meta']['simulation']['size'] == 16
	for bad, msg in (
		(_fake_log(exit_ok = False), 'did not exit properly'),
		(_fake_log(extra = 'not json'), 'non-JSON'),
		(_fake_log(extra = '{"log": "ERROR", "msg": "boom"}'), 'has errors'),
		(_fake_log(steps = 0), 'did not run any steps'),
		(_fake_log().replace('"counter": 2', '"counter": 7'), 'unexpected sequence'),
		):
		with pytest.raises(SyntaxError, match = msg):
			analyze.parse_log(bad)


def test_analyze_summary_rate_and_fraction_of_peak(tmp_path):
	"""dtype and GPU count as axes, interactions/s and %-of-peak columns"""
	import json
	from gravitation_b200.cli import analyze
	runs = analyze.parse_log(_fake_log())
	runs[0]['meta']['simulation'].update(scenario_param = {'stars_len': 16, 'dtype': 'float64'}, threads = 2)
	rows = analyze.summarize(runs)
	assert len(rows) == 1 and rows[0]['dtype'] == 'float64' and rows[0]['threads'] == 2 and rows[0]['bodies'] == 16
	rate = 16 * 15 / 101e-9 # best of the three fake steps is 101 ns
	assert rows[0]['g_interactions_per_s'] == pytest.approx(rate / 1e9)
	assert rows[0]['fraction_of_peak'] == pytest.approx(rate * 20 / (2 * 37.22e12))
	assert analyze.summarize(runs, peak_tflops = 10.0)[0]['fraction_of_peak'] == pytest.approx(rate * 20 / (2 * 10e12))
	log = tmp_path / 'b.log'
	log.write_text(_fake_log() + _fake_log(steps = 2))
	analyze.main(['-l', str(log), '-o', str(tmp_path / 'b.json'), '--summary'])
	assert len(json.loads((tmp_path / 'b.json.summary.json').read_text())) == 2
	assert 'G interactions/s' in analyze.format_summary(rows)


def test_benchmark_size_range_matches_reference_rule():
	from gravitation_b200.cli.benchmark import size_range, worker_command
	assert size_range(2, 4) == [4, 6, 8, 12, 16] and size_range(3, 3) == [8]
	cmd = worker_command('b200', 4096, 2, 10, 0)
	assert cmd[1:3] == ['-m', 'gravitation_b200.cli.worker'] and '{"stars_len": 4096}' in cmd and cmd[-1] == '2'


# ---- npnn: the independent float64 numpy kernel of the inventory -------------------

def _npnn(n, dtype, seed = 42, scenario = 'galaxy'):
	from gravitation_b200.kernel import npnn
	param = {'dtype': dtype, 'seed': seed}
	if scenario == 'galaxy':
		param['stars_len'] = n
	return simulation.create_simulation(scenario, npnn.universe, param)


@pytest.mark.parametrize('case,n,scenario', (('solarsystem', 2, 'solarsystem'), ('galaxy256', 256, 'galaxy'), ('galaxy4096', 4096, 'galaxy')))
def test_npnn_is_pinned_to_the_reference_goldens(case, n, scenario, golden, oracle):
	"""accelerations and 10-step trajectories of the N x N numpy kernel at float64 against what the REFERENCE's
	np2 at float64 (and py1) produced (tests/golden/make_golden.py); different summation order, same physics"""
	g = golden[case]
	u = _npnn(n, 'float64', scenario = scenario)
	assert 'npnn' in load.inventory and np.array_equal(u.mass_r_array, g['r0'])
	u.step_stage1()
	assert oracle.max_rel_err(u.mass_a_array, g['acc_np2_f64']) < 1e-12
	if 'acc_py1' in g:
		assert oracle.max_rel_err(u.mass_a_array, g['acc_py1']) < 1e-12
	first = [pm for pm in u][n - 1]
	assert np.array_equal(np.asarray(first._a), u.mass_a_array[n - 1, :]) # point masses are row views
	u.step_stage2(); u.step_stage3()
	for _ in range(9):
		u.step()
	scale = np.abs(g['r10_np2_f64']).max()
	assert np.abs(u.mass_r_array - g['r10_np2_f64']).max() / scale < 1e-12
	assert oracle.max_rel_err(u.mass_v_array, g['v10_np2_f64']) < 1e-10
	assert u._t == 10 * u._T
	u.stop()


def test_npnn_float32_state_is_within_the_reference_fp32_envelope(golden, oracle):
	g = golden['galaxy256']
	u = _npnn(256, 'float32')
	assert u.mass_r_array.dtype == np.float32 and np.array_equal(u.mass_r_array, g['r0_f32'])
	u.step_stage1()
	assert oracle.max_rel_err(u.mass_a_array, g['acc_np2_f64']) < 1e-6 # input rounding only: the arithmetic is float64
	for _ in range(10):
		u.step()
	assert oracle.max_rel_err(u.mass_r_array[1:], g['r10_np2_f64'][1:]) < 5e-6


def test_benchmark_cli_sweeps_dtype_through_worker_and_analyze(tmp_path):
	"""SURVEY 8f rank 4 end to end on the CPU kernel: `benchmark` spawns one worker per (threads, N) with the
	dtype riding in --scenario_param, the log parses, and `analyze --summary` has the dtype / threads / N axes"""
	import json
	from gravitation_b200.cli import analyze, benchmark
	log = str(tmp_path / 'bench.log')
	for dtype in ('float32', 'float64'):
		rc = benchmark.main(['-k', 'npnn', '-b', '5', '6', '-i', '2', '-t', '0', '-p', '1', '-p', '2', '-l', log,
			'--scenario_param', json.dumps({'dtype': dtype, 'seed': 3})])
		assert rc == 0
	analyze.main(['-l', log, '-o', str(tmp_path / 'bench.json'), '--summary'])
	rows = json.loads((tmp_path / 'bench.json.summary.json').read_text())
	# npnn is not parallel: `-p 1 -p 2` collapses to threads = 1 (benchmark.py:188-192 of the reference)
	assert [(r['kernel'], r['dtype'], r['threads'], r['bodies']) for r in rows] == [
		('npnn', dt, 1, n) for dt in ('float32', 'float64') for n in (32, 48, 64)]
	assert all(r['steps'] >= 2 and r['g_interactions_per_s'] > 0 for r in rows)


def test_accuracy_cli_compares_two_kernel_runs(capsys):
	from gravitation_b200.cli import accuracy
	out = accuracy.main(['-k', 'npnn', '--dtype', 'float32', '--ref_kernel', 'npnn', '--ref_dtype', 'float64', '-n', '128', '-s', '3'])
	assert out['test'] == {'kernel': 'npnn', 'dtype': 'float32'} and out['reference'] == {'kernel': 'npnn', 'dtype': 'float64'}
	assert 0 < out['acceleration_max_rel'] < 1e-6 and 0 < out['position_max_rel'] < 5e-6 and out['bodies'] == 128


# ---- bench.py contract: exactly one JSON line on stdout, whatever libraries print ------------------

def test_bench_reference_arm_prints_one_json_line():
	"""`bench.py --impl reference` (the CPU arm; the one place outside tests that may execute oracle/): one
	JSON line on stdout with the contract's keys; everything else (library banners, warnings) is on stderr"""
	import json, subprocess, sys
	root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
	if not os.path.isfile(os.path.join(root, 'oracle', '_ref', 'lib4.so')) and not os.path.isfile(os.path.join(root, 'oracle', 'liboracle.so')):
		pytest.skip('oracle not built')
	out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '3', '--bodies', '14'],
		capture_output = True, text = True, timeout = 600)
	assert out.returncode == 0, out.stderr[-2000:]
	lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
	assert len(lines) == 1
	line = json.loads(lines[0])
	assert line['impl'] == 'reference' and line['unit'] == 'G interactions/s' and line['higher_is_better'] is True
	assert line['value'] > 0 and line['e2e']['value'] == line['value'] and line['e2e']['h2d_bytes_per_step'] == 0
	assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
	# the line names the size it actually ran, not the size it stands in for
	assert line['config']['n_bodies'] == 1 << 14 and line['config']['sampled_from_n_bodies'] == 1 << 20
	assert line['config']['same_config'] is False and 'N=2^14' in line['config']['workload'] and '2^14' in line['cpu_baseline']['sample']


# ---- the symmetric sweep's work decomposition, replayed on the host (csrc/nbody_sym.cuh) ------------

def test_symmetric_schedule_visits_every_block_pair_once(tmp_path):
	"""tests/native/sym_schedule_check.cu includes the kernel header and replays its geometry helpers and its
	tile walker on the CPU: every unordered pair of body-blocks exactly once over all shards, row_start table
	== walked tiles, walker order == nested loops from any flat offset, ragged last block covered, block
	rows balanced to one column block.  nvcc only cross-compiles; nothing runs on a device."""
	import shutil, subprocess
	nvcc = shutil.which('nvcc') or ('/usr/local/cuda/bin/nvcc' if os.path.isfile('/usr/local/cuda/bin/nvcc') else None)
	if nvcc is None:
		pytest.skip('nvcc not available')
	exe = str(tmp_path / 'sym_schedule_check')
	build = subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O1', '-std=c++17', '-o', exe,
		os.path.join(ROOT, 'tests', 'native', 'sym_schedule_check.cu')], capture_output = True, text = True)
	assert build.returncode == 0, build.stderr[-3000:]
	run = subprocess.run([exe], capture_output = True, text = True, timeout = 300)
	assert run.returncode == 0 and run.stdout.strip().splitlines()[-1].startswith('OK:'), run.stdout[-3000:]


def test_row_partition_properties_hypothesis(shim):
	"""random (n, world, dtype): slices are contiguous, cover [0, n), hold ceil(n / world) rows each except the
	short (possibly empty) tail"""
	from hypothesis import given, settings, strategies as st

	@settings(max_examples = 300, deadline = None)
	@given(n = st.integers(min_value = 0, max_value = 1 << 25), world = st.integers(min_value = 1, max_value = 8),
		dtype = st.sampled_from(('float32', 'float64')))
	def check(n, world, dtype):
		parts = shim.partition(n, world, dtype)
		assert len(parts) == world and sum(c for _, c in parts) == n
		chunk = parts[0][1]
		pos = 0
		for row0, cnt in parts:
			assert row0 == min(pos, n) and 0 <= cnt <= chunk
			pos += chunk
		assert chunk == -(-n // world)
		assert all(cnt == chunk for _, cnt in parts[:max(0, (n // chunk if chunk else 0))])
	check()


# ---- static evidence from the compiled library: what the default kernels are made of ---------------

def test_sass_of_the_default_kernels(shim):
	"""cuobjdump -sass of libgravb200.so (sm_100a): the automatic large-N variants stage j-tiles with TMA bulk
	copies (UBLKCP) behind mbarriers probed without blocking (SYNCS.PHASECHK, no TRYWAIT), compute with packed
	FFMA2 / DFMA and MUFU.RSQ, keep everything in registers (no local-memory spills), and the symmetric
	sweeps have no CTA-wide barrier inside their tile loop (BARs only around it: mbarrier initialisation, peer hand-over)"""
	import collections, shutil, subprocess
	tool = shutil.which('cuobjdump') or ('/usr/local/cuda/bin/cuobjdump' if os.path.isfile('/usr/local/cuda/bin/cuobjdump') else None)
	if tool is None:
		pytest.skip('cuobjdump not available')
	text = subprocess.run([tool, '-sass', shim.LIB_PATH], capture_output = True, text = True, check = True).stdout
	assert 'sm_100a' in text
	funcs, cur = {}, None
	for line in text.splitlines():
		m = re.search(r'Function : (\S+)', line)
		if m:
			cur = m.group(1); funcs[cur] = []
			continue
		m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
		if cur is not None and m:
			funcs[cur].append(re.sub(r'^@!?U?P\d+\s+', '', m.group(1)))
	def census(key):
		names = [f for f in funcs if key in f]
		assert len(names) == 1, (key, names)
		ins = funcs[names[0]]
		return collections.Counter(i.split()[0].split('.')[0] for i in ins), ins
	for key, must, bars in (
		('sym_sweep_kernelILi256ELi12ELi512ELi3ELi2ELi0E', ('UBLKCP', 'FFMA2', 'FADD2', 'FMUL2', 'MUFU', 'SHFL', 'REDG', 'VOTE'), 6), # fp32 symmetric, variant 100
		('sym_sweep_kernelILi256ELi12ELi512ELi3ELi2ELi1E', ('UBLKCP', 'FFMA2', 'FADD2', 'FMUL2', 'MUFU', 'SHFL', 'REDG', 'VOTE'), 6), # its twin with chunk-granular CTA ranges (mid-sized N)
		('sym_sweep_kernel_f64ILi256ELi8ELi256ELi3ELi1ELi1ELi0E', ('UBLKCP', 'DFMA', 'MUFU', 'SHFL', 'REDG', 'VOTE'), 6), # fp64 symmetric, variant 101
		('sym_sweep_kernel_f64ILi256ELi8ELi256ELi3ELi1ELi1ELi1E', ('UBLKCP', 'DFMA', 'MUFU', 'SHFL', 'REDG', 'VOTE'), 6), # its chunk-granular twin
		('sweep_kernelIfLi256ELi8ELi512ELi3ELi1ELi1ELi4ELi1E', ('UBLKCP', 'FFMA2', 'FADD2', 'FMUL2', 'MUFU', 'VOTE'), None), # fp32 ordered, variant 0
		('sweep_kernelIdLi256ELi2ELi256ELi3ELi2ELi0ELi4ELi0E', ('UBLKCP', 'DFMA', 'MUFU', 'VOTE'), None), # fp64 ordered, variant 0
		('small_steps_kernelIfLi256ELi4ELi4E', ('UBLKCP', 'FFMA2', 'FADD2', 'FMUL2', 'MUFU', 'VOTE', 'ATOMG', 'SHFL'), None), # fp32 persistent small-N, variant 200
		('small_steps_kernelIdLi256ELi4ELi4E', ('UBLKCP', 'DFMA', 'MUFU', 'VOTE', 'ATOMG', 'SHFL'), None), # fp64 persistent small-N, variant 200
		):
		ops, ins = census(key)
		for op in must:
			# fp64 accumulation into the global accumulator: RED, or ATOM without a result where a release fence follows
			assert ops[op] > 0 or (op == 'REDG' and ops['ATOMG'] > 0), (key, op)
		assert ops['STL'] == 0 and ops['LDL'] == 0, (key, 'spills to local memory')
		assert any('PHASECHK' in i for i in ins) and not any('TRYWAIT' in i for i in ins), (key, 'mbarrier waits must be non-blocking probes')
		assert any(i.startswith('MUFU.RSQ') for i in ins)
		if bars is not None:
			# mbarrier initialisation, the wait for the peers' flags and the share of the tile list at the start, the hand-over to the peers at the
			# end (twice: idle CTAs leave early) — none of them inside the sweep loop
			# (ncu: smsp__average_warps_issue_stalled_barrier = 0 in profiles/r0*_ncu_sym_*_summary.md)
			assert 1 <= ops['BAR'] <= bars, (key, ops['BAR'])
