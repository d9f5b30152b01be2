#!/usr/bin/env python3
"""Generates tests/golden/*.npz by RUNNING THE REFERENCE (pleiszenburg/gravitation, /root/reference) in
the dev container.  The reference cannot travel to the GPU box, the vectors can.

What is recorded (the reference has no tests or golden vectors of its own, so the
pins are outputs of its kernels on seeded universes):
  solarsystem.npz   2 bodies (lib/simulation.py:85-97): state, accelerations of py1 / np2@f64 after one
                    step_stage1, state after 10 steps (np2@f64)
  galaxy256.npz     galaxy scenario, random.seed(42), N = 256: initial state as the reference built it,
                    accelerations of py1, np2@f64, np2@f32, c1a (oracle/_ref/lib1.so = the reference's
                    _lib1_/lib.c compiled unmodified), state after 10 steps of np2@f64 and np2@f32
  galaxy4096.npz    same at N = 4096 without py1 (24 s/step)
Usage: python tests/golden/make_golden.py   (needs /root/reference and oracle/_ref built)"""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference/src')
sys.path.insert(0, ROOT)
sys.modules.setdefault('h5py', types.ModuleType('h5py')) # lib/simulation.py:35 imports it at module level

from gravitation.lib.simulation import create_simulation # noqa: E402
from gravitation.kernel import np2, py1 # noqa: E402
from oracle import oracle # noqa: E402


def state(u):
	n = len(u)
	r = np.array([[float(x) for x in pm._r] for pm in u]).reshape(n, 3)
	v = np.array([[float(x) for x in pm._v] for pm in u]).reshape(n, 3)
	a = np.array([[float(x) for x in pm._a] for pm in u]).reshape(n, 3)
	m = np.array([float(pm._m) for pm in u])
	return r, v, a, m


def build(scenario, cls, n, dtype, seed = 42):
	random.seed(seed)
	param = {'dtype': dtype}
	if scenario == 'galaxy':
		param['stars_len'] = n
	return create_simulation(scenario, cls, param, threads = 1)


def run(scenario, n, with_py1):
	out = {}
	u64 = build(scenario, np2.universe, n, 'float64')
	r0, v0, _, m = state(u64)
	out.update(r0 = r0, v0 = v0, m = m, G = np.float64(u64._G), T = np.float64(u64._T))
	u64.step_stage1()
	out['acc_np2_f64'] = state(u64)[2]
	u64.step_stage2(); u64.step_stage3()
	for _ in range(9):
		u64.step()
	r10, v10, _, _ = state(u64)
	out.update(r10_np2_f64 = r10, v10_np2_f64 = v10)

	u32 = build(scenario, np2.universe, n, 'float32')
	r0_32, v0_32, _, m32 = state(u32)
	out.update(r0_f32 = r0_32.astype(np.float32), v0_f32 = v0_32.astype(np.float32), m_f32 = m32.astype(np.float32))
	u32.step_stage1()
	out['acc_np2_f32'] = state(u32)[2].astype(np.float32)
	u32.step_stage2(); u32.step_stage3()
	for _ in range(9):
		u32.step()
	r10, v10, _, _ = state(u32)
	out.update(r10_np2_f32 = r10.astype(np.float32), v10_np2_f32 = v10.astype(np.float32))

	if with_py1:
		up = build(scenario, py1.universe, n, 'float64')
		up.step_stage1()
		out['acc_py1'] = state(up)[2]
	# the reference's scalar C kernel on the float32 state (c1a.py:85-91 feeds it exactly these values)
	out['acc_c1a'] = oracle.ref_c1a_stage1(out['r0_f32'], out['m_f32'], float(u32._G))
	return out


if __name__ == '__main__':
	assert oracle.have_ref(), 'build oracle/_ref first: make -C oracle'
	for name, scenario, n, with_py1 in (
		('solarsystem', 'solarsystem', 2, True),
		('galaxy256', 'galaxy', 256, True),
		('galaxy4096', 'galaxy', 4096, False),
		):
		data = run(scenario, n, with_py1)
		path = os.path.join(HERE, name + '.npz')
		np.savez_compressed(path, **data)
		print(name, {k: (v.shape, str(v.dtype)) for k, v in data.items()}, os.path.getsize(path))
