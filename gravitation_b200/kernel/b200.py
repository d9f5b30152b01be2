# -*- coding: utf-8 -*-
"""B200-native gravitation kernel: the drop-in module for `src/gravitation/kernel/` of
pleiszenburg/gravitation (reference path: /root/reference/src/gravitation/kernel/).

Front end unchanged: a class `universe(universe_base)`, the nine literal meta dunders the reference's
inventory parses by AST (`lib/load.py:77-99`), float32 by default and float64 on request, point masses
readable through iteration after every step.  Back end: `libgravb200.so` (hand-written CUDA for
sm_100a, `include/gravb200.h`) through ctypes.  There is no CPU path in this module — if the library or
a B200 is missing, `start()` raises.

Differences to the reference's own GPU kernel pc2 (`pc2.py`):
  * state lives on the device; nothing is copied per step (pc2: 3 H2D + 3 D2H per step, `pc2.py:147-162`)
  * stage 2 is fused into the stage-1 kernel's epilogue; `step_stage2` only commits + synchronises
  * `threads` = number of GPUs (`__parallel__ = True` makes `gravitation benchmark -p 1 -p 2 ...` sweep it)
  * host mirrors (`mass_r_array`, `mass_v_array`, `mass_a_array`) are refreshed lazily, the first time
    the point masses are read after a step (`eager_host=True` restores a download every step)
Extras (additive): `add_objects` (bulk constructor), `steps(k)`, `accelerations()`, `push_host_state()`.
"""

# KERNEL META (literals only: parsed without importing this module, lib/load.py:77-99)

__longname__ = 'b200-backend'
__version__ = '0.1.0'
__description__ = 'hand-written CUDA for sm_100a via C-ABI shim, O(N*(N-1)), device resident, 1-8 GPUs'
__requirements__ = ['numpy']
__externalrequirements__ = ['cuda', 'nccl']
__interpreters__ = ['python3']
__parallel__ = True
__license__ = 'GPLv2'
__authors__ = [
	'gravitation_b200 authors',
	]

import os
import threading
import warnings

import numpy as np

from ._base_ import universe_base, _point_mass, STATE_PREINIT, STATE_STARTED, STATE_STOPPED
try: # inside this repository
	from .. import _shim
except ImportError: # dropped into the reference tree next to an underscore-prefixed `_b200_` package
	from ._b200_ import _shim

# Everything this module needs beyond numpy is `_base_` (only names the reference's own `_base_.py` has:
# universe_base, _point_mass, STATE_*) and `_shim` — so the file works unchanged inside the reference tree
# (tests/test_dropin.py installs it there).


class _synced_list(list):
	"""`_mass_list` after `start()`: a plain list whose reads first bring the host mirrors up to date"""

	def __init__(self, items, sync):
		super().__init__(items)
		self._sync = sync

	def __iter__(self):
		self._sync()
		return super().__iter__()

	def __getitem__(self, key):
		self._sync()
		return super().__getitem__(key)


class _bulk_masses:
	"""`_mass_list` of a universe filled by `add_objects`: point-mass records are created on access as
	row views of the host mirrors, so 2^20 .. 2^24 bodies cost no Python objects up front"""

	def __init__(self, owner, names):
		self._owner = owner
		self._names = names

	def __len__(self):
		return self._owner._bulk_r.shape[0]

	def _make(self, k):
		o = self._owner
		started = o._state != STATE_PREINIT
		r = o.mass_r_array if started else o._bulk_r
		v = o.mass_v_array if started else o._bulk_v
		m = o.mass_m_array if started else o._bulk_m
		pm = _point_mass(
			name = self._names[k] if self._names is not None else 'body',
			r = r[k, :], v = v[k, :], m = float(m[k]),
			)
		if started:
			pm._a = o.mass_a_array[k, :]
		return pm

	def __getitem__(self, key):
		self._owner._sync_host()
		if isinstance(key, slice):
			return [self._make(k) for k in range(*key.indices(len(self)))]
		if key < 0:
			key += len(self)
		if not 0 <= key < len(self):
			raise IndexError('point mass index out of range')
		return self._make(key)

	def __iter__(self):
		self._owner._sync_host()
		return (self._make(k) for k in range(len(self)))

	def append(self, item):
		raise SyntaxError('universe was filled with add_objects; add_object cannot be mixed in')


class universe(universe_base):

	# ---------------------------------------------------------------------------------------------
	# additive front end
	# ---------------------------------------------------------------------------------------------

	def add_objects(self, r, v, m, names = None, scale_off = False):
		"""bulk `add_object`: r, v array-likes (N,3), m (N,).  Same unit scaling as `add_object`
		(reference `_base_.py:114-117`) unless `scale_off`.  Must be the only way this universe is filled."""
		if self._state == STATE_STARTED: # same guards and messages as `add_object` (reference `_base_.py:110-113`)
			raise SyntaxError('simulation was started')
		if self._state == STATE_STOPPED:
			raise SyntaxError('simulation was stopped')
		if len(self._mass_list) != 0:
			raise SyntaxError('add_objects needs an empty universe')
		r = np.array(r, dtype = np.float64)
		v = np.array(v, dtype = np.float64)
		m = np.array(m, dtype = np.float64)
		if r.ndim != 2 or r.shape[1] != 3 or v.shape != r.shape or m.shape != (r.shape[0],):
			raise ValueError('expected r, v of shape (N, 3) and m of shape (N,)')
		if not scale_off:
			r *= self._scale_r
			v *= self._scale_r
			m *= self._scale_m
		self._bulk_r, self._bulk_v, self._bulk_m = r, v, m
		self._mass_list = _bulk_masses(self, names)

	def steps(self, k):
		"""k full steps on the device without returning to Python in between"""
		if self._state == STATE_PREINIT: # same guards and messages as `step` (reference `_base_.py:139-142`)
			raise SyntaxError('simulation was not started')
		if self._state == STATE_STOPPED:
			raise SyntaxError('simulation was stopped')
		if len(self._shards) == 1:
			self._shards[0].steps(k)
		else:
			for _ in range(k):
				self.step_stage1()
				self._commit()
		for _ in range(k): # k roundings, exactly like k calls of step_stage3 (`_base_.py:158-161`)
			self._t += self._T
		self._stale_rv = True
		self._stale_a = True

	def accelerations(self):
		"""(N,3) accelerations of the last `step_stage1` as a host array"""
		self._sync_host()
		return self.mass_a_array

	def push_host_state(self):
		"""re-upload positions/velocities after the caller edited the host mirrors (masses, G, T stay).
		One process per GPU with `host_rows='own'`: every rank sends only its own rows and the device exchange
		(NVLink) completes the position array on all shards — collective, like `start()`."""
		if self._own_rows_only:
			sh = self._shards[0]
			rows = slice(sh.row0, sh.row0 + sh.n_local)
			sh.upload_rows(self.mass_r_array[rows, :], self.mass_v_array[rows, :])
		else:
			for sh in self._shards:
				sh.upload(self.mass_r_array, self.mass_v_array, self.mass_m_array, self._G, self._T, self._eps)
		self._stale_rv = self._stale_a = False

	# ---------------------------------------------------------------------------------------------
	# kernel hooks
	# ---------------------------------------------------------------------------------------------

	def start_kernel(self):
		self.DTYPE = self._dtype
		if self.DTYPE not in ('float32', 'float64'):
			raise ValueError('dtype must be float32 or float64, got %r' % (self.DTYPE,))
		self.MASS_LEN = len(self)
		self.SIM_DIM = 3
		if self.MASS_LEN < 1:
			raise ValueError('empty universe')
		self._eps = float(self._meta.get('eps', 0.0))
		self._eager = bool(self._meta.get('eager_host', False))
		# one process per GPU: 'all' (default) mirrors all positions on every rank, 'own' only this rank's rows
		# (velocities and accelerations exist for own rows only in either case)
		self._own_rows_only = 'world' in self._meta and int(self._meta['world']) > 1 and self._meta.get('host_rows', 'all') == 'own'
		if self._meta.get('host_rows', 'all') not in ('all', 'own'):
			raise ValueError("host_rows must be 'all' or 'own'")
		n = self.MASS_LEN
		# host mirrors, laid out like the reference's numpy kernels (np2.py:63-66)
		# (page-locked, so the on-demand downloads and `push_host_state` run at full PCIe rate)
		self._pinned = [_shim.PinnedArray(shape, self.DTYPE) for shape in ((n, 3), (n, 3), (n, 3), (n,))]
		self.mass_r_array, self.mass_v_array, self.mass_a_array, self.mass_m_array = (p.array for p in self._pinned)
		if isinstance(self._mass_list, _bulk_masses):
			self.mass_r_array[:, :] = self._bulk_r
			self.mass_v_array[:, :] = self._bulk_v
			self.mass_m_array[:] = self._bulk_m
		else:
			for k, pm in enumerate(self._mass_list):
				if len(pm._r) != 3:
					raise ValueError('this kernel is three-dimensional')
				self.mass_m_array[k] = pm._m
				self.mass_r_array[k, :] = pm._r[:]
				self.mass_v_array[k, :] = pm._v[:]
				# point masses become row views of the mirrors (np2.py:70-75)
				pm._r = self.mass_r_array[k, :]
				pm._v = self.mass_v_array[k, :]
				pm._a = self.mass_a_array[k, :]
			self._mass_list = _synced_list(self._mass_list, self._sync_host)
		self._stale_rv = self._stale_a = False
		self._shards = self._make_shards(n)
		for sh in self._shards:
			sh.upload(self.mass_r_array, self.mass_v_array, self.mass_m_array, self._G, self._T, self._eps)

	def _make_shards(self, n):
		meta = self._meta
		if 'world' in meta: # one process per GPU (torchrun): this process owns shard `rank`
			shard = _shim.Shard(
				n, self.DTYPE, device = int(meta.get('device', 0)),
				rank = int(meta['rank']), world = int(meta['world']), nccl_id = meta.get('nccl_id'),
				)
			if int(meta['world']) > 1:
				mode = _shim.connect_peers(shard, want = self._want_peer_exchange())
				self._note_exchange(mode)
			return [shard]
		gpus = int(self._threads)
		if gpus < 1:
			raise ValueError('threads (= number of GPUs) must be >= 1')
		have = _shim.device_count()
		if gpus > have:
			raise _shim.GravB200Error('threads=%d GPUs requested, %d visible' % (gpus, have))
		if gpus == 1:
			return [_shim.Shard(n, self.DTYPE, device = int(meta.get('device', 0)))]
		# several GPUs in this process: NCCL communicators must be created concurrently
		uid = _shim.nccl_unique_id()
		shards, errors = [None] * gpus, []
		def make(rank):
			try:
				shards[rank] = _shim.Shard(n, self.DTYPE, device = rank, rank = rank, world = gpus, nccl_id = uid)
			except Exception as e: # re-raised on the caller's thread below
				errors.append(e)
		workers = [threading.Thread(target = make, args = (rank,)) for rank in range(gpus)]
		for w in workers:
			w.start()
		for w in workers:
			w.join()
		if errors:
			raise errors[0]
		# fused exchange: every shard maps every other shard's position buffers (direct peer access) and
		# the sweep's epilogue stores r' there; all-or-nothing, NCCL all-gather otherwise
		mode, why = _shim.XCHG_NCCL, None
		if self._want_peer_exchange():
			try:
				blobs = [sh.peer_export() for sh in shards]
				for sh in shards:
					sh.peer_connect(blobs)
				mode = _shim.XCHG_PEER
			except _shim.GravB200Error as e:
				mode, why = _shim.XCHG_NCCL, str(e)
		for sh in shards:
			sh.set_exchange_mode(mode)
		self._note_exchange(mode, why)
		return shards

	def _want_peer_exchange(self):
		return str(self._meta.get('exchange', os.environ.get('GRAVB200_EXCHANGE', 'peer'))).lower() != 'nccl'

	def _note_exchange(self, mode, why = None):
		"""records the position exchange in use (`exchange_mode`: 'peer' | 'nccl') and, when the fused
		peer-store exchange was wanted but is not available, says so loudly: the NCCL all-gather also rules out
		the symmetric sweep on several shards, i.e. costs about a quarter of the throughput"""
		self.exchange_mode = 'peer' if mode == _shim.XCHG_PEER else 'nccl'
		self.exchange_fallback = None
		if mode != _shim.XCHG_PEER and self._want_peer_exchange():
			self.exchange_fallback = why or 'a rank could not map its peers (CUDA IPC / peer access)'
			warnings.warn(
				'b200 kernel: fused peer-store exchange unavailable (%s); falling back to the NCCL all-gather '
				'with the ordered sweep (about 28 %% slower)' % self.exchange_fallback, RuntimeWarning, stacklevel = 2,
				)

	def step_stage1(self):
		"""launches the sweep on every shard (asynchronous): accelerations + fused v', r' into back buffers"""
		for sh in self._shards:
			sh.stage1()
		self._stale_a = True
		if self._eager:
			self._sync_host()

	def _commit(self):
		if len(self._shards) > 1:
			# enqueue the exchange on EVERY shard before waiting on any (one host thread drives them all)
			nccl = self._shards[0].info()['exchange_mode'] == _shim.XCHG_NCCL
			if nccl:
				_shim.group_begin()
			for sh in self._shards:
				sh.exchange()
			if nccl:
				_shim.group_end()
		for sh in self._shards:
			sh.stage2()

	def step_stage2(self):
		"""commits the back buffers (after the position exchange on several GPUs) and waits for the device"""
		self._commit()
		self._stale_rv = True
		if self._eager:
			self._sync_host()

	def stop_kernel(self):
		self._sync_host()
		for sh in self._shards:
			sh.close()
		self._shards = []

	# ---------------------------------------------------------------------------------------------
	# host mirrors
	# ---------------------------------------------------------------------------------------------

	def _sync_host(self):
		"""device -> host mirrors for whatever changed since the last read"""
		if self._state == STATE_PREINIT or not getattr(self, '_shards', None):
			return
		if not (self._stale_rv or self._stale_a):
			return
		if self._own_rows_only:
			sh = self._shards[0]
			rows = slice(sh.row0, sh.row0 + sh.n_local)
			sh.download_rows(
				r = self._stale_rv, v = self._stale_rv, a = self._stale_a,
				out_r = self.mass_r_array[rows, :], out_v = self.mass_v_array[rows, :], out_a = self.mass_a_array[rows, :],
				)
			self._stale_rv = self._stale_a = False
			return
		for k, sh in enumerate(self._shards):
			rows = slice(sh.row0, sh.row0 + sh.n_local)
			sh.download(
				r = self._stale_rv and k == 0, v = self._stale_rv, a = self._stale_a,
				out_r = self.mass_r_array,
				out_v = self.mass_v_array[rows, :],
				out_a = self.mass_a_array[rows, :],
				)
		self._stale_rv = self._stale_a = False
