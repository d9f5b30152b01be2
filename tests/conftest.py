# -*- coding: utf-8 -*-
"""pytest configuration: registers the `gpu` marker, puts the repo root on sys.path and provides the
golden fixtures.  `-m "not gpu"` must pass on a CPU-only box; `-m gpu` needs a B200."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
	config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def load_golden(name):
	with np.load(os.path.join(GOLDEN, name + '.npz')) as f:
		return {k: f[k] for k in f.files}


@pytest.fixture(scope = 'session')
def golden():
	return {name: load_golden(name) for name in ('solarsystem', 'galaxy256', 'galaxy4096')}


@pytest.fixture(scope = 'session')
def oracle():
	from oracle import oracle as o
	o.lib() # builds liboracle.so on first use
	return o


@pytest.fixture(scope = 'session')
def shim():
	"""the ctypes binding; builds libgravb200.so first if this checkout has not been built yet
	(nvcc cross-compiles for sm_100a without a GPU; normally __graft_entry__.build() did that already)"""
	from gravitation_b200 import _shim
	if not os.path.isfile(_shim.LIB_PATH):
		import subprocess
		subprocess.run(['make', '-C', os.path.join(ROOT, 'gravitation_b200', 'csrc'), 'all'], check = True, capture_output = True)
	_shim.load()
	return _shim
