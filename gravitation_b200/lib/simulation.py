# -*- coding: utf-8 -*-
"""Scenario builders: the input generators of the hot path.

Restates /root/reference/src/gravitation/lib/simulation.py (behaviour, not text):
  create_simulation   `simulation.py:41-83`   scenario name + kernel class -> started universe
  create_solarsystem  `simulation.py:85-97`   sun + earth, SI units
  create_galaxy       `simulation.py:99-184`  central mass + disc (80 %) + bulge (20 %) on circular orbits
  store_simulation / load_simulation  `simulation.py:186-258`  snapshot I/O (HDF5 when h5py is
                      importable, otherwise the same datasets/attributes in an .npz)

Additions: the galaxy scenario accepts `seed` in `scenario_param` (the reference draws from the
unseeded global `random`, `simulation.py:116`), and universes that offer `add_objects` are filled in
bulk, which is what makes N = 2^20 .. 2^24 constructible."""

import math
import os
import random

import numpy as np

_GALAXY_UNIVERSE = dict(T = 2.0e12, scale_m = 1.0e-30, scale_r = 1.0e-10, dtype = 'float32')
_GALAXY_SCREEN = dict(unit = 1e20, unit_size = [16.0, 10.0], average_over_steps = 20, steps_per_frame = 1)
_SOLAR_SCREEN = dict(unit = 149597870700, unit_size = [3.0, 2.25], average_over_steps = 500, steps_per_frame = 20)


def create_simulation(scenario, universe_class, scenario_param = None, threads = 1):
	"""builds, fills and starts a universe of `universe_class` for a named scenario"""
	scenario_param = dict(scenario_param) if scenario_param is not None else {}
	universe_param = {'threads': threads}
	if scenario == 'solarsystem':
		universe_param.update(scenario_param)
		universe_obj = universe_class(**universe_param)
		universe_obj._screen = dict(_SOLAR_SCREEN)
		create_solarsystem(universe_obj)
	elif scenario == 'galaxy':
		universe_param.update(_GALAXY_UNIVERSE)
		universe_param.update(scenario_param)
		universe_obj = universe_class(**universe_param)
		universe_obj._screen = dict(_GALAXY_SCREEN)
		create_galaxy(
			universe_obj = universe_obj,
			stars_len = scenario_param.get('stars_len', 2000) - 1, # `stars_len` counts the central mass
			r = [0.0, 0.0, 0.0], v = [0.0, 0.0, 0.0], g_alpha = 0.0, g_beta = 0.0,
			m_hole = 4e40, m_star = 2e30, radius = 1e20,
			seed = scenario_param.get('seed', None), builder = scenario_param.get('builder', None),
			)
	else:
		raise ValueError('Unknown scenario: "%s"' % scenario)
	universe_obj.start()
	return universe_obj


def create_solarsystem(universe_obj):
	universe_obj.add_object(name = 'sun', r = [0.0, 0.0, 0.0], v = [0.0, 0.0, 0.0], m = 1.98892e30)
	universe_obj.add_object(name = 'earth', r = [0.0, -149597870700.0, 0.0], v = [29777.777, 0.0, 0.0], m = 5.97237e24)


def _star(rnd, n, stars_len, G, r0, v0, g_alpha, g_beta, m_hole, radius):
	"""position and velocity of star n; consumes the RNG in the reference's order
	(alpha, r_abs, then z-jitter for disc stars or beta for bulge stars)"""
	alpha = rnd.random() * 2.0 * math.pi
	r_out = (4.5 + 0.1) * radius
	if n < (stars_len * 4 // 5): # disc
		r_abs = (rnd.random() * 4.5 + 0.1) * radius
		r_s = [
			r_abs * math.cos(alpha),
			r_abs * math.sin(alpha),
			(0.5 * rnd.random() - 0.25) * radius * (r_out - r_abs) / r_out,
			]
	else: # central bulge
		r_abs = (rnd.random() * 0.75 + 0.1) * radius
		beta = math.pi * (rnd.random() - 0.5)
		r_s = [
			r_abs * math.cos(alpha) * math.cos(beta),
			r_abs * math.sin(alpha) * math.cos(beta),
			r_abs * math.sin(beta),
			]
	# circular orbit around the central mass, velocity at a right angle to the radius
	# sum() on purpose: since Python 3.12 it is a compensated sum, and the reference calls it (`simulation.py:148`)
	v_abs = math.sqrt(G * m_hole / math.sqrt(sum([d ** 2 for d in r_s])))
	v_alpha = alpha - (math.pi / 2)
	v_s = [v_abs * math.cos(v_alpha), v_abs * math.sin(v_alpha), 0.0]
	# tilt by g_beta around x, turn by g_alpha around z (velocity, then position)
	v_s[1:] = [v_s[1] * math.cos(g_beta), v_s[1] * math.sin(g_beta)]
	ang = math.atan2(v_s[1], v_s[0]) + g_alpha
	mag = math.sqrt(v_s[0] ** 2 + v_s[1] ** 2)
	v_s[0:2] = [mag * math.cos(ang), mag * math.sin(ang)]
	v_s = [a + b for a, b in zip(v_s, v0)]
	ang = math.atan2(r_s[2], r_s[1]) + g_beta
	mag = math.sqrt(r_s[2] ** 2 + r_s[1] ** 2)
	r_s[1:] = [mag * math.cos(ang), mag * math.sin(ang)]
	ang = math.atan2(r_s[1], r_s[0]) + g_alpha
	mag = math.sqrt(r_s[0] ** 2 + r_s[1] ** 2)
	r_s[0:2] = [mag * math.cos(ang), mag * math.sin(ang)]
	r_s = [a + b for a, b in zip(r_s, r0)]
	return r_s, v_s


_VECTOR_BUILDER_FROM = 65536 + 1 # the reference's own stream up to 2^16 bodies, vectorised above


def galaxy_arrays(stars_len, G, r, v, g_alpha, g_beta, m_hole, m_star, radius, seed = None):
	"""numpy-vectorised restatement of the galaxy builder (`simulation.py:114-184`) for universes the per-star
	Python loop cannot build in reasonable time (2^20 stars: minutes; 2^24: hours).  Same distributions and
	the same construction — central mass, 80 % disc stars with r in [0.1, 4.6) radius and a z-jitter that
	tapers towards the rim, 20 % bulge stars with r in [0.1, 0.85) radius, every star on a circular orbit
	around the central mass, tilt g_beta around x and turn g_alpha around z — but drawn from its OWN stream
	(`numpy.random.default_rng(seed)`), so the bodies are NOT those of `random.seed(seed)` + the reference.
	Returns R, V (stars_len + 1, 3) and M (stars_len + 1,) in the caller's units, body 0 = central mass."""
	rng = np.random.default_rng(seed)
	n = int(stars_len)
	n_disc = n * 4 // 5
	alpha = rng.random(n) * 2.0 * math.pi
	u_r = rng.random(n)
	u_3 = rng.random(n) # z-jitter (disc) or latitude (bulge)
	disc = np.arange(n) < n_disc
	r_out = (4.5 + 0.1) * radius
	r_abs = np.where(disc, u_r * 4.5 + 0.1, u_r * 0.75 + 0.1) * radius
	beta = np.where(disc, 0.0, math.pi * (u_3 - 0.5))
	P = np.empty((n, 3))
	P[:, 0] = r_abs * np.cos(alpha) * np.cos(beta)
	P[:, 1] = r_abs * np.sin(alpha) * np.cos(beta)
	P[:, 2] = np.where(disc, (0.5 * u_3 - 0.25) * radius * (r_out - r_abs) / r_out, r_abs * np.sin(beta))
	v_abs = np.sqrt(G * m_hole / np.sqrt((P * P).sum(axis = 1)))
	W = np.zeros((n, 3))
	W[:, 0] = v_abs * np.cos(alpha - math.pi / 2)
	W[:, 1] = v_abs * np.sin(alpha - math.pi / 2)
	cb, sb, ca, sa = math.cos(g_beta), math.sin(g_beta), math.cos(g_alpha), math.sin(g_alpha)
	tilt = np.array([[1.0, 0.0, 0.0], [0.0, cb, -sb], [0.0, sb, cb]]) # around x
	turn = np.array([[ca, -sa, 0.0], [sa, ca, 0.0], [0.0, 0.0, 1.0]]) # around z
	rot = (turn @ tilt).T
	R = np.empty((n + 1, 3)); V = np.empty((n + 1, 3)); M = np.full(n + 1, float(m_star))
	R[0, :], V[0, :], M[0] = r, v, m_hole
	R[1:, :] = P @ rot + np.asarray(r, dtype = np.float64)
	V[1:, :] = W @ rot + np.asarray(v, dtype = np.float64)
	return R, V, M


def create_galaxy(universe_obj, stars_len, r, v, g_alpha, g_beta, m_hole, m_star, radius, seed = None, builder = None):
	"""central mass 'back hole' (sic, `simulation.py:109`) plus `stars_len` stars.
	builder 'reference' (default up to 2^16 bodies): the reference's per-star construction — seed None draws
	from the global `random` stream like the reference, otherwise a private seeded stream that yields the same
	bodies as `random.seed(seed)` followed by the reference's builder.  builder 'vector' (default above 2^16
	bodies when the universe offers `add_objects`): `galaxy_arrays`, own stream, seconds instead of minutes."""
	can_bulk = hasattr(universe_obj, 'add_objects')
	if builder is None:
		builder = 'vector' if (can_bulk and stars_len + 1 >= _VECTOR_BUILDER_FROM) else 'reference'
	if builder not in ('reference', 'vector'):
		raise ValueError('Unknown galaxy builder: "%s"' % builder)
	G = universe_obj._G # the universe's (already unit-scaled) G, as `simulation.py:148`
	if builder == 'vector':
		if not can_bulk:
			raise ValueError('the vector builder needs a universe with add_objects')
		R, V, M = galaxy_arrays(stars_len, G, r, v, g_alpha, g_beta, m_hole, m_star, radius, seed)
		universe_obj.add_objects(R, V, M, names = _names(stars_len + 1))
		return
	rnd = random if seed is None else random.Random(seed)
	bulk = can_bulk and stars_len >= 4096
	if not bulk:
		universe_obj.add_object(name = 'back hole', r = [d for d in r], v = [d for d in v], m = m_hole)
		for n in range(stars_len):
			r_s, v_s = _star(rnd, n, stars_len, G, r, v, g_alpha, g_beta, m_hole, radius)
			universe_obj.add_object(name = 'star', r = r_s, v = v_s, m = m_star)
		return
	R = np.empty((stars_len + 1, 3)); V = np.empty((stars_len + 1, 3)); M = np.full(stars_len + 1, m_star)
	R[0, :], V[0, :], M[0] = r, v, m_hole
	for n in range(stars_len):
		R[n + 1, :], V[n + 1, :] = _star(rnd, n, stars_len, G, r, v, g_alpha, g_beta, m_hole, radius)
	names = _names(stars_len + 1)
	universe_obj.add_objects(R, V, M, names = names)


class _names:
	"""'back hole', 'star', 'star', ... without a list of N strings"""

	def __init__(self, n):
		self._n = n

	def __len__(self):
		return self._n

	def __getitem__(self, k):
		return 'back hole' if k == 0 else 'star'


# -------------------------------------------------------------------------------------------------
# snapshots
# -------------------------------------------------------------------------------------------------

_ATTRS = ('scale_m', 'scale_r', 't', 'T', 'G', 'dtype', 'threads')


def _snapshot_arrays(universe_obj):
	n = len(universe_obj)
	r = np.empty((n, 3)); v = np.empty((n, 3)); m = np.empty(n)
	names = []
	for k, pm in enumerate(universe_obj): # the documented read path, `simulation.py:234-238`
		names.append(pm._name)
		r[k, :] = pm._r[:]
		v[k, :] = pm._v[:]
		m[k] = pm._m
	return r, v, m, names


def _h5py():
	"""h5py if it is importable (it is an optional dependency here; the reference hard-imports it, `simulation.py:35`)"""
	try:
		import h5py
		return h5py
	except ImportError:
		return None


def _npz_path(fn):
	return fn if fn.endswith('.npz') else fn + '.npz'


def _plain(val):
	"""HDF5 attribute / 0-d array -> plain Python value (str, int, float)"""
	if isinstance(val, bytes):
		return val.decode('utf-8')
	if isinstance(val, np.ndarray) and val.ndim == 0:
		val = val[()]
	if isinstance(val, np.generic):
		val = val.item()
	if isinstance(val, bytes):
		return val.decode('utf-8')
	return val


def store_simulation(universe_obj, fn, gn):
	"""appends snapshot `gn` to file `fn`: datasets r, v (N x dim), m, name and the universe attributes.
	HDF5 with the reference's layout (`simulation.py:216-258`: group `gn`, datasets `r`, `v`, `m` in the run
	dtype, `name` as fixed-width bytes, attributes scale_m, scale_r, t, T, G, dtype, threads) if h5py is
	importable and `fn` does not ask for `.npz`; otherwise `<fn>.npz` with keys `<gn>/<dataset>` and
	`<gn>/attr/<attribute>`.  Returns the path written."""
	r, v, m, names = _snapshot_arrays(universe_obj)
	dtype = {'float32': '<f4', 'float64': '<f8'}[universe_obj._dtype]
	attrs = {a: getattr(universe_obj, '_' + a) for a in _ATTRS}
	name_array = np.array([s.encode('utf-8') for s in names]) # dtype 'S<longest name>', as `simulation.py:239-244`
	h5py = _h5py()
	if h5py is not None and not fn.endswith('.npz'):
		with h5py.File(fn, 'a') as f:
			dg = f.create_group(gn)
			dg.create_dataset('r', data = r.astype(dtype))
			dg.create_dataset('v', data = v.astype(dtype))
			dg.create_dataset('m', data = m.astype(dtype))
			dg.create_dataset('name', data = name_array)
			for key, val in attrs.items():
				dg.attrs[key] = val
		return fn
	path = _npz_path(fn)
	store = {}
	try:
		with np.load(path, allow_pickle = False) as old:
			store.update({k: old[k] for k in old.files})
	except FileNotFoundError:
		pass
	store[gn + '/r'] = r.astype(dtype)
	store[gn + '/v'] = v.astype(dtype)
	store[gn + '/m'] = m.astype(dtype)
	store[gn + '/name'] = name_array
	for key, val in attrs.items():
		store[gn + '/attr/' + key] = np.array(val)
	np.savez(path, **store)
	return path


def _read_snapshot(fn, gn):
	"""(param dict, r, v, m, names) of snapshot `gn`: from the HDF5 file `fn` when there is one and h5py can
	open it (`simulation.py:186-196`), else from the `.npz` twin `store_simulation` writes without h5py"""
	h5py = _h5py()
	if h5py is not None and not fn.endswith('.npz') and os.path.isfile(fn) and h5py.is_hdf5(fn):
		with h5py.File(fn, 'r') as f:
			dg = f[gn]
			param = {str(key): _plain(dg.attrs[key]) for key in dg.attrs.keys()}
			r, v, m, names = (np.array(dg[key]) for key in ('r', 'v', 'm', 'name'))
		return param, r, v, m, names
	path = _npz_path(fn)
	if not os.path.isfile(path):
		if os.path.isfile(fn):
			raise OSError('%s is an HDF5 snapshot file, but h5py is not importable here' % fn)
		raise FileNotFoundError('no snapshot file %s (or %s)' % (fn, path))
	with np.load(path, allow_pickle = False) as f:
		param = {key: _plain(f[gn + '/attr/' + key]) for key in _ATTRS}
		r, v, m, names = f[gn + '/r'], f[gn + '/v'], f[gn + '/m'], f[gn + '/name']
	return param, r, v, m, names


def load_simulation(universe_class, fn, gn, threads = None):
	"""rebuilds a (not yet started) universe from snapshot `gn` of file `fn` (`simulation.py:186-214`): values
	are stored in internal units, hence `scale_off`"""
	param, r, v, m, names = _read_snapshot(fn, gn)
	if isinstance(threads, int):
		param['threads'] = threads
	universe_obj = universe_class(scale_off = True, **param)
	if hasattr(universe_obj, 'add_objects'):
		universe_obj.add_objects(r, v, m, names = [bytes(s).decode('utf-8') for s in names], scale_off = True)
	else:
		for k in range(r.shape[0]):
			universe_obj.add_object(
				scale_off = True, name = bytes(names[k]).decode('utf-8'),
				r = [float(x) for x in r[k, :]], v = [float(x) for x in v[k, :]], m = float(m[k]),
				)
	return universe_obj
