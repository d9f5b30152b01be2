# -*- coding: utf-8 -*-
"""numpy N x N kernel — an INDEPENDENT CPU implementation of the same time step, shipped as a second entry of
the inventory so that `gravitation accuracy -k b200 --ref_kernel npnn` compares the CUDA path with something
that is not itself (/root/reference/TODO.md:4 asks for such a tool).

It is its own kernel, never a fallback: `b200` does not import it and fails without a GPU.  It is also not a
copy of the reference's numpy kernels: np1/np2 walk the unique pairs row by row (`np2.py:89-108`); this one
evaluates the per-body N x N form the CUDA kernel uses (`pc2.py:66-89`: a_i = G sum_{j != i} m_j d / |d|^3,
self pair excluded by index) for blocks of rows at once, always in float64, whatever `dtype` the state has.
Stage 2 is the array form of `np2.py:110-115` in the state dtype (four separately rounded operations).
Pinned to the reference's golden vectors in tests/test_host.py."""

# KERNEL META (literals only: parsed without importing this module, lib/load.py:77-99)

__longname__ = 'numpy-nxn-float64-backend'
__version__ = '0.1.0'
__description__ = 'numpy, row-blocked N x N form, float64 accumulation, O(N*(N-1)); reference for accuracy checks'
__requirements__ = ['numpy']
__externalrequirements__ = []
__interpreters__ = ['python3']
__parallel__ = False
__license__ = 'GPLv2'
__authors__ = [
	'gravitation_b200 authors',
	]

import numpy as np

from ._base_ import universe_base

_BLOCK_BYTES = 48 << 20 # budget for the (rows, N, 3) float64 difference block


class universe(universe_base):

	def start_kernel(self):
		n = len(self)
		if n < 1:
			raise ValueError('empty universe')
		self.SIM_DIM = len(self._mass_list[0]._r)
		self.MASS_LEN = n
		dt = np.dtype(self._dtype)
		self.mass_r_array = np.zeros((n, self.SIM_DIM), dtype = dt)
		self.mass_v_array = np.zeros((n, self.SIM_DIM), dtype = dt)
		self.mass_a_array = np.zeros((n, self.SIM_DIM), dtype = dt)
		self.mass_m_array = np.zeros((n,), dtype = dt)
		for k, pm in enumerate(self._mass_list):
			self.mass_r_array[k, :] = pm._r[:]
			self.mass_v_array[k, :] = pm._v[:]
			self.mass_m_array[k] = pm._m
			pm._r = self.mass_r_array[k, :]
			pm._v = self.mass_v_array[k, :]
			pm._a = self.mass_a_array[k, :]
		self._rows = max(1, min(n, _BLOCK_BYTES // (n * self.SIM_DIM * 8)))
		self._eps2 = float(self._meta.get('eps', 0.0)) ** 2

	def step_stage1(self):
		r = self.mass_r_array.astype(np.float64)
		m = self.mass_m_array.astype(np.float64)
		n = self.MASS_LEN
		for i0 in range(0, n, self._rows):
			i1 = min(n, i0 + self._rows)
			d = r[None, :, :] - r[i0:i1, None, :] # (rows, N, dim): r_j - r_i
			d2 = np.einsum('ijk,ijk->ij', d, d) + self._eps2
			d2[np.arange(i1 - i0), np.arange(i0, i1)] = np.inf # the self pair contributes nothing
			w = m[None, :] / (d2 * np.sqrt(d2))
			self.mass_a_array[i0:i1, :] = self._G * np.einsum('ij,ijk->ik', w, d)

	def step_stage2(self):
		T = self.mass_r_array.dtype.type(self._T)
		np.multiply(self.mass_a_array, T, out = self.mass_a_array)
		np.add(self.mass_v_array, self.mass_a_array, out = self.mass_v_array)
		np.add(self.mass_r_array, self.mass_v_array * T, out = self.mass_r_array)
