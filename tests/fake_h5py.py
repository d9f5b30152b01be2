# -*- coding: utf-8 -*-
"""A tiny stand-in for the part of the h5py API the snapshot code uses (File / group / dataset / attrs /
is_hdf5), persisted with pickle.  h5py is not installed in this image, so without it the HDF5 branch of
`lib/simulation.py` would never execute under test; with it the branch logic (layout, attribute handling,
append mode, type conversion on the way back) runs on every CPU test pass.  The real-h5py twin of the test
is guarded by `pytest.importorskip('h5py')`."""
import os
import pickle

import numpy as np

MAGIC = b'FAKEHDF5'


class _Attrs(dict):
	def __setitem__(self, key, val):
		# h5py hands attributes back as numpy scalars / str
		if isinstance(val, str):
			super().__setitem__(key, val)
		else:
			super().__setitem__(key, np.array(val)[()])


class _Group:
	def __init__(self):
		self.datasets, self.attrs = {}, _Attrs()

	def create_dataset(self, name, shape = None, dtype = None, data = None):
		if name in self.datasets:
			raise ValueError('dataset exists')
		arr = np.array(data) if data is not None else np.zeros(shape, dtype = dtype)
		self.datasets[name] = arr
		return arr

	def __getitem__(self, name):
		return self.datasets[name]


class File:
	def __init__(self, fn, mode = 'r'):
		self._fn, self._mode, self._groups = fn, mode, {}
		if os.path.isfile(fn):
			with open(fn, 'rb') as f:
				assert f.read(len(MAGIC)) == MAGIC
				self._groups = pickle.load(f)
		elif mode == 'r':
			raise OSError('no such file')

	def create_group(self, name):
		if self._mode == 'r':
			raise ValueError('read-only')
		if name in self._groups:
			raise ValueError('group exists')
		self._groups[name] = _Group()
		return self._groups[name]

	def __getitem__(self, name):
		return self._groups[name]

	def keys(self):
		return self._groups.keys()

	def close(self):
		if self._mode != 'r':
			with open(self._fn, 'wb') as f:
				f.write(MAGIC)
				pickle.dump(self._groups, f)

	def __enter__(self):
		return self

	def __exit__(self, *exc):
		self.close()


def is_hdf5(fn):
	try:
		with open(fn, 'rb') as f:
			return f.read(len(MAGIC)) == MAGIC
	except OSError:
		return False
