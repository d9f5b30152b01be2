# -*- coding: utf-8 -*-
"""ctypes binding of libgravb200.so (include/gravb200.h).

This is the only way Python reaches the CUDA code.  There is no fallback of any kind: if the shared
library is missing `load()` raises, and if it cannot find an sm_100 device every compute call raises
`GravB200Error` with the library's own message.

The binding mirrors how the reference binds its native kernels
(/root/reference/src/gravitation/kernel/c4b.py:86-93: `ctypes.cdll.LoadLibrary(.../lib.so)` + argtypes).
"""

import ctypes
import os

import numpy as np

F32, F64 = 0, 1
NCCL_ID_BYTES = 128
PEER_BLOB_BYTES = 512
XCHG_NCCL, XCHG_PEER = 0, 1
_DTYPES = {'float32': F32, 'float64': F64}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libgravb200.so')

# every symbol include/gravb200.h declares: (name, restype, argtypes)
_c_ctx = ctypes.c_void_p
_SIGNATURES = [
	('gravb200_abi_version', ctypes.c_int, []),
	('gravb200_device_count', ctypes.c_int, []),
	('gravb200_last_error', ctypes.c_char_p, []),
	('gravb200_ctx_create', ctypes.c_int, [
		ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
		ctypes.c_void_p, ctypes.POINTER(_c_ctx),
		]),
	('gravb200_ctx_destroy', ctypes.c_int, [_c_ctx]),
	('gravb200_nccl_unique_id', ctypes.c_int, [ctypes.c_void_p]),
	('gravb200_upload', ctypes.c_int, [
		_c_ctx, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
		ctypes.c_double, ctypes.c_double, ctypes.c_double,
		]),
	('gravb200_upload_positions', ctypes.c_int, [_c_ctx, ctypes.c_void_p]),
	('gravb200_upload_rows', ctypes.c_int, [_c_ctx, ctypes.c_void_p, ctypes.c_void_p]),
	('gravb200_download_rows', ctypes.c_int, [_c_ctx, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
	('gravb200_stage1', ctypes.c_int, [_c_ctx]),
	('gravb200_stage2', ctypes.c_int, [_c_ctx]),
	('gravb200_exchange', ctypes.c_int, [_c_ctx]),
	('gravb200_peer_barrier', ctypes.c_int, [_c_ctx]),
	('gravb200_group_begin', ctypes.c_int, []),
	('gravb200_group_end', ctypes.c_int, []),
	('gravb200_peer_export', ctypes.c_int, [_c_ctx, ctypes.c_void_p]),
	('gravb200_peer_connect', ctypes.c_int, [_c_ctx, ctypes.c_void_p]),
	('gravb200_set_exchange_mode', ctypes.c_int, [_c_ctx, ctypes.c_int]),
	('gravb200_steps', ctypes.c_int, [_c_ctx, ctypes.c_int]),
	('gravb200_sync', ctypes.c_int, [_c_ctx]),
	('gravb200_download', ctypes.c_int, [_c_ctx, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
	('gravb200_shard', ctypes.c_int, [_c_ctx, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
	('gravb200_partition', ctypes.c_int, [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
	('gravb200_timings', ctypes.c_int, [_c_ctx, ctypes.POINTER(ctypes.c_float), ctypes.c_int]),
	('gravb200_info', ctypes.c_int, [_c_ctx, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
	('gravb200_set_variant', ctypes.c_int, [_c_ctx, ctypes.c_int]),
	('gravb200_set_split', ctypes.c_int, [_c_ctx, ctypes.c_int]),
	('gravb200_variant_count', ctypes.c_int, [ctypes.c_int]),
	('gravb200_sym_variant_count', ctypes.c_int, [ctypes.c_int]),
	('gravb200_small_variant_count', ctypes.c_int, []),
	('gravb200_small_geometry', ctypes.c_int, [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
	('gravb200_variant_name', ctypes.c_char_p, [ctypes.c_int, ctypes.c_int]),
	('gravb200_device_ptr', ctypes.c_void_p, [_c_ctx, ctypes.c_int]),
	('gravb200_host_alloc', ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
	('gravb200_host_free', ctypes.c_int, [ctypes.c_void_p]),
	('gravb200_peak_probe', ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int]),
	]
SYMBOLS = tuple(name for name, _, _ in _SIGNATURES)

_lib = None


class GravB200Error(RuntimeError):
	"""a libgravb200 call returned a negative status"""


def _prefer_bundled_nccl():
	"""libgravb200 dlopens NCCL by soname when a multi-GPU context is created.  If that happened before
	`import torch`, the system libnccl.so.2 would win and torch (which needs the newer NCCL bundled in
	its wheel) could no longer be imported into the process.  Point the library at the bundled copy when
	there is one; GRAVB200_NCCL_LIB set by the user is respected."""
	if os.environ.get('GRAVB200_NCCL_LIB'):
		return
	try:
		import importlib.util
		spec = importlib.util.find_spec('nvidia.nccl')
		for loc in (spec.submodule_search_locations if spec is not None else []):
			cand = os.path.join(loc, 'lib', 'libnccl.so.2')
			if os.path.isfile(cand):
				os.environ['GRAVB200_NCCL_LIB'] = cand
				return
	except Exception:
		pass


def load():
	"""dlopen libgravb200.so and declare every prototype; raises if the library was not built"""
	global _lib
	if _lib is not None:
		return _lib
	if not os.path.isfile(LIB_PATH):
		raise GravB200Error(
			'%s not found: build it with `make -C gravitation_b200/csrc` '
			'(or `python -c "import __graft_entry__ as g; g.build()"`). There is no CPU fallback.' % LIB_PATH
			)
	_prefer_bundled_nccl()
	lib = ctypes.CDLL(LIB_PATH, mode = ctypes.RTLD_GLOBAL)
	for name, restype, argtypes in _SIGNATURES:
		fn = getattr(lib, name) # AttributeError if the symbol is missing
		fn.restype = restype
		fn.argtypes = argtypes
	_lib = lib
	return lib


def _check(rc):
	if rc != 0:
		raise GravB200Error('libgravb200: %s (code %d)' % (_lib.gravb200_last_error().decode('utf-8', 'replace'), rc))


def device_count():
	return load().gravb200_device_count()


def nccl_unique_id():
	buf = ctypes.create_string_buffer(NCCL_ID_BYTES)
	_check(load().gravb200_nccl_unique_id(buf))
	return buf.raw


def group_begin():
	_check(load().gravb200_group_begin())


def group_end():
	_check(load().gravb200_group_end())


def peak_probe(device = 0):
	"""measured non-tensor peaks: dict(fp32_tflops, fp32x2_tflops, fp64_tflops, mufu_gops, sm_mhz)"""
	out = (ctypes.c_double * 5)()
	_check(load().gravb200_peak_probe(device, out, 5))
	return dict(fp32_tflops = out[0], fp32x2_tflops = out[1], fp64_tflops = out[2], mufu_gops = out[3], sm_mhz = out[4])


SYM_BASE = 100 # ids of the symmetric sweeps start here
SMALL_BASE = 200 # ids of the persistent multi-step kernel for small universes start here


def variant_names(dtype = 'float32'):
	"""ordered sweeps of `dtype` (ids 0 .. len-1)"""
	lib = load()
	d = _DTYPES[dtype]
	return [lib.gravb200_variant_name(d, i).decode() for i in range(lib.gravb200_variant_count(d))]


def partition(n, world, dtype = 'float32'):
	"""[(row0, n_local)] of every shard — the partition gravb200_ctx_create applies (needs no device)"""
	lib = load()
	out = []
	for rank in range(world):
		row0, n_local = ctypes.c_int64(), ctypes.c_int64()
		_check(lib.gravb200_partition(int(n), _DTYPES[dtype], int(world), rank, ctypes.byref(row0), ctypes.byref(n_local)))
		out.append((row0.value, n_local.value))
	return out


def sym_variant_names(dtype = 'float32'):
	"""symmetric sweeps of `dtype` (ids SYM_BASE + k)"""
	lib = load()
	d = _DTYPES[dtype]
	return [lib.gravb200_variant_name(d, SYM_BASE + k).decode() for k in range(lib.gravb200_sym_variant_count(d))]


def small_variant_names():
	"""persistent small-N kernels (ids SMALL_BASE + k; the same list for both dtypes)"""
	lib = load()
	return [lib.gravb200_variant_name(_DTYPES['float32'], SMALL_BASE + k).decode() for k in range(lib.gravb200_small_variant_count())]


def small_geometry(n, dtype = 'float32', sm_count = 148, variant = SMALL_BASE):
	"""geometry of the persistent small-N kernel (needs no device): dict(grid, rows_per_cta, row_groups, slice,
	slices, smem_bytes, fits); raises when `n` needs more rows per CTA than the variant has lanes for"""
	out = (ctypes.c_int64 * 6)()
	rc = load().gravb200_small_geometry(int(n), _DTYPES[dtype], int(sm_count), int(variant), out, 6)
	if rc < 0:
		_check(rc)
	return dict(grid = out[0], rows_per_cta = out[1], row_groups = out[2], slice = out[3], slices = out[4], smem_bytes = out[5], fits = rc == 0)


class _PinnedBlock:
	"""owns one cudaHostAlloc allocation and exposes it through the array interface, so every numpy view made
	from it (`np.asarray(block)` and all views of that) keeps the block — and thereby the page-locked memory —
	alive.  The memory is freed when the last view is gone, never earlier (no use-after-free through a row
	view that outlives its universe)."""

	def __init__(self, nbytes):
		self._lib = load()
		p = ctypes.c_void_p()
		_check(self._lib.gravb200_host_alloc(max(int(nbytes), 1), ctypes.byref(p)))
		self._p = p
		self.__array_interface__ = {'shape': (max(int(nbytes), 1),), 'typestr': '|u1', 'data': (p.value, False), 'version': 3}

	def __del__(self):
		try:
			if self._p is not None and self._p.value:
				self._lib.gravb200_host_free(self._p)
				self._p = None
		except Exception:
			pass


class PinnedArray:
	"""a page-locked host buffer as a numpy array (`.array`).  The allocation belongs to the ARRAY (its base
	chain ends in a `_PinnedBlock`), not to this wrapper: views handed out to callers stay valid after the
	universe that created them is gone, like the reference's plain numpy arrays."""

	def __init__(self, shape, dtype):
		dt = np.dtype(dtype)
		count = int(np.prod(shape))
		block = _PinnedBlock(count * dt.itemsize)
		self.array = np.asarray(block)[:count * dt.itemsize].view(dt).reshape(shape)
		self.array[...] = 0


def _ptr(a):
	return None if a is None else ctypes.c_void_p(a.ctypes.data)


class Shard:
	"""one GPU's share of a universe: thin object wrapper over gravb200_ctx"""

	def __init__(self, n_total, dtype = 'float32', device = 0, rank = 0, world = 1, nccl_id = None):
		self._lib = load()
		self._np = np.dtype(dtype)
		self.n_total = int(n_total)
		self.dtype = dtype
		self.rank, self.world = rank, world
		ctx = _c_ctx()
		idbuf = None
		if nccl_id is not None:
			idbuf = ctypes.create_string_buffer(bytes(nccl_id), NCCL_ID_BYTES)
		_check(self._lib.gravb200_ctx_create(
			self.n_total, _DTYPES[dtype], device, rank, world, idbuf, ctypes.byref(ctx),
			))
		self._ctx = ctx
		row0, n_local = ctypes.c_int64(), ctypes.c_int64()
		_check(self._lib.gravb200_shard(self._ctx, ctypes.byref(row0), ctypes.byref(n_local)))
		self.row0, self.n_local = row0.value, n_local.value

	def _arr(self, a, shape):
		a = np.ascontiguousarray(a, dtype = self._np)
		if a.shape != shape:
			raise ValueError('expected shape %s, got %s' % (shape, a.shape))
		return a

	def upload(self, r, v, m, G, T, eps = 0.0):
		r = self._arr(r, (self.n_total, 3))
		v = self._arr(v, (self.n_total, 3))
		m = self._arr(m, (self.n_total,))
		_check(self._lib.gravb200_upload(self._ctx, _ptr(r), _ptr(v), _ptr(m), G, T, eps))

	def upload_positions(self, r):
		r = self._arr(r, (self.n_total, 3))
		_check(self._lib.gravb200_upload_positions(self._ctx, _ptr(r)))

	def upload_rows(self, r_own, v_own = None):
		"""this shard's rows only ([n_local, 3] each); collective over the shards of a universe"""
		r_own = self._arr(r_own, (self.n_local, 3))
		if v_own is not None:
			v_own = self._arr(v_own, (self.n_local, 3))
		_check(self._lib.gravb200_upload_rows(self._ctx, _ptr(r_own), _ptr(v_own)))

	def download_rows(self, r = True, v = True, a = False, out_r = None, out_v = None, out_a = None):
		"""this shard's rows only: (r, v, a), each [n_local, 3] or None"""
		outs = []
		for want, arr in ((r, out_r), (v, out_v), (a, out_a)):
			if want and arr is None:
				arr = np.empty((self.n_local, 3), dtype = self._np)
			if want and not (arr.flags.c_contiguous and arr.dtype == self._np and arr.shape == (self.n_local, 3)):
				raise ValueError('download targets must be C-contiguous (n_local, 3) arrays of dtype %s' % self.dtype)
			outs.append(arr if want else None)
		_check(self._lib.gravb200_download_rows(self._ctx, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2])))
		return tuple(outs)

	def upload_raw(self, r_ptr, v_ptr, m_ptr, G, T, eps = 0.0):
		"""host pointers (ints) of C-contiguous buffers in the context dtype, e.g. pinned memory"""
		_check(self._lib.gravb200_upload(self._ctx, r_ptr, v_ptr, m_ptr, G, T, eps))

	def download_raw(self, r_ptr, v_ptr, a_ptr):
		_check(self._lib.gravb200_download(self._ctx, r_ptr, v_ptr, a_ptr))

	def stage1(self):
		_check(self._lib.gravb200_stage1(self._ctx))

	def peer_export(self):
		buf = ctypes.create_string_buffer(PEER_BLOB_BYTES)
		_check(self._lib.gravb200_peer_export(self._ctx, buf))
		return buf.raw

	def peer_connect(self, blobs):
		"""blobs: the `peer_export()` of every shard, in rank order"""
		joined = b''.join(bytes(b) for b in blobs)
		if len(joined) != PEER_BLOB_BYTES * self.world:
			raise ValueError('need %d blobs of %d bytes' % (self.world, PEER_BLOB_BYTES))
		_check(self._lib.gravb200_peer_connect(self._ctx, ctypes.create_string_buffer(joined, len(joined))))

	def set_exchange_mode(self, mode):
		_check(self._lib.gravb200_set_exchange_mode(self._ctx, int(mode)))

	def exchange(self):
		_check(self._lib.gravb200_exchange(self._ctx))

	def stage2(self):
		_check(self._lib.gravb200_stage2(self._ctx))

	def steps(self, k):
		_check(self._lib.gravb200_steps(self._ctx, int(k)))

	def sync(self):
		_check(self._lib.gravb200_sync(self._ctx))

	def download(self, r = True, v = True, a = False, out_r = None, out_v = None, out_a = None):
		"""returns (r[n_total,3] | None, v[n_local,3] | None, a[n_local,3] | None)"""
		if r and out_r is None:
			out_r = np.empty((self.n_total, 3), dtype = self._np)
		if v and out_v is None:
			out_v = np.empty((self.n_local, 3), dtype = self._np)
		if a and out_a is None:
			out_a = np.empty((self.n_local, 3), dtype = self._np)
		for arr in (out_r, out_v, out_a):
			if arr is not None and not (arr.flags.c_contiguous and arr.dtype == self._np):
				raise ValueError('download targets must be C-contiguous arrays of dtype %s' % self.dtype)
		_check(self._lib.gravb200_download(
			self._ctx, _ptr(out_r if r else None), _ptr(out_v if v else None), _ptr(out_a if a else None),
			))
		return (out_r if r else None, out_v if v else None, out_a if a else None)

	def timings(self):
		ms = (ctypes.c_float * 10)()
		_check(self._lib.gravb200_timings(self._ctx, ms, 10))
		out = dict(sweep_ms = ms[0], exchange_ms = ms[1], steps_ms = ms[2], sm_mhz = ms[3], cta0_ms = ms[4])
		if ms[5] >= 0: # several shards, symmetric sweep: where the step's time went (the integrate kernel starts by waiting for every shard's sweep, the tail by waiting for every shard's integrate)
			out['phases_ms'] = dict(sweep_kernel = ms[5], integrate_incl_wait_for_sweeps = ms[6], tail_wait = ms[7])
			if ms[9] >= 0: # of the integrate kernel's time: the part after its wait for the peers' sweeps
				out['phases_ms']['integrate_work'] = ms[9]
			if ms[8] >= 0: # speed-proportional shares: this shard's part of the tile list against the equal share
				out['phases_ms']['share_of_equal'] = ms[8]
		return out

	def info(self):
		v = (ctypes.c_int64 * 13)()
		_check(self._lib.gravb200_info(self._ctx, v, 13))
		keys = ('grid', 'threads', 'bodies_per_thread', 'tile', 'stages', 'smem_bytes', 'launches', 'sm_count', 'packed', 'ctas_per_sm', 'exchange_mode', 'variant', 'split')
		return dict(zip(keys, [int(x) for x in v]))

	def peer_barrier(self):
		"""enqueue a flag barrier with all peer shards (collective; no-op on one shard / in NCCL mode)"""
		_check(self._lib.gravb200_peer_barrier(self._ctx))

	def set_variant(self, variant):
		_check(self._lib.gravb200_set_variant(self._ctx, int(variant)))

	def set_split(self, mode):
		"""symmetric sweeps: CTA ranges cut at chunk (1) or tile (0) granularity, -1 = automatic"""
		_check(self._lib.gravb200_set_split(self._ctx, int(mode)))

	def device_ptr(self, which):
		return self._lib.gravb200_device_ptr(self._ctx, which)

	def close(self):
		if self._ctx is not None and self._ctx.value:
			self._lib.gravb200_ctx_destroy(self._ctx)
			self._ctx = None

	def __del__(self):
		try:
			self.close()
		except Exception:
			pass


def connect_peers(shard, want = None):
	"""One process per GPU: switches `shard` to the fused peer-store exchange if EVERY rank of the
	torch.distributed group can map every peer's buffers (CUDA IPC over NVLink); otherwise all ranks stay on
	the NCCL all-gather.  The decision is collective — a mixed world would deadlock.  `want` False (or
	GRAVB200_EXCHANGE=nccl) keeps NCCL.  torch.distributed is used for the rendezvous only (blob gather +
	one MIN all-reduce), never on the data path.  Returns the mode in use (XCHG_PEER / XCHG_NCCL)."""
	import torch
	import torch.distributed as tdist
	if want is None:
		want = os.environ.get('GRAVB200_EXCHANGE', 'peer').lower() != 'nccl'
	ok = 1
	try:
		blobs = [None] * shard.world
		tdist.all_gather_object(blobs, shard.peer_export())
		if want:
			shard.peer_connect(blobs)
	except GravB200Error:
		ok = 0
	t = torch.tensor([ok if want else 0], dtype = torch.int32)
	if tdist.get_backend() == 'nccl':
		t = t.cuda()
	tdist.all_reduce(t, op = tdist.ReduceOp.MIN)
	mode = XCHG_PEER if int(t.item()) == 1 else XCHG_NCCL
	shard.set_exchange_mode(mode)
	return mode
