# -*- coding: utf-8 -*-
"""`gravitation accuracy`: run two kernel/dtype combinations on ONE seeded universe and report how far the
accelerations after the first `step_stage1()` and the state after k steps are apart — the tool the
reference lists as wanted (/root/reference/TODO.md:4) and the parity harness of SURVEY.md section 8d as a
command.  Metrics: max_i |a_i - b_i| / |b_i| for accelerations, positions and velocities (bodies at the
origin are compared relative to the universe's extent).

  python -m gravitation_b200.cli.accuracy -k b200 --dtype float32 --ref_kernel b200 --ref_dtype float64 -n 4096 -s 10

Inside the reference tree any two kernels of the inventory can be compared (e.g. `-k b200 --ref_kernel np2`).
"""

import argparse
import json
import sys

import numpy as np

from ..lib.load import inventory
from ..lib.simulation import create_simulation


def _rel(x, ref):
	num = np.linalg.norm(x - ref, axis = 1)
	den = np.linalg.norm(ref, axis = 1)
	den = np.where(den < 1e-6 * den.max(), den.max(), den)
	return float(np.max(num / den))


def _run(kernel, dtype, scenario, n, seed, steps, threads):
	inventory[kernel].load_module()
	u = create_simulation(scenario, inventory[kernel].get_class(), {'stars_len': n, 'seed': seed, 'dtype': dtype}, threads = threads)
	u.step_stage1()
	a = np.array([[float(c) for c in pm._a] for pm in u])
	u.step_stage2(); u.step_stage3()
	for _ in range(steps - 1):
		u.step()
	r = np.array([[float(c) for c in pm._r] for pm in u])
	v = np.array([[float(c) for c in pm._v] for pm in u])
	u.stop()
	return a, r, v


def main(argv = None):
	ap = argparse.ArgumentParser(description = 'compare two kernels / dtypes on one seeded universe')
	names = sorted(inventory.keys())
	ap.add_argument('--kernel', '-k', required = True, choices = names)
	ap.add_argument('--dtype', default = 'float32')
	ap.add_argument('--ref_kernel', choices = names)
	ap.add_argument('--ref_dtype', default = 'float64')
	ap.add_argument('--scenario', default = 'galaxy')
	ap.add_argument('--bodies', '-n', type = int, default = 4096)
	ap.add_argument('--steps', '-s', type = int, default = 10)
	ap.add_argument('--seed', type = int, default = 42)
	ap.add_argument('--threads', '-p', type = int, default = 1)
	a = ap.parse_args(argv)
	ref_kernel = a.ref_kernel or a.kernel
	acc, r, v = _run(a.kernel, a.dtype, a.scenario, a.bodies, a.seed, a.steps, a.threads)
	acc0, r0, v0 = _run(ref_kernel, a.ref_dtype, a.scenario, a.bodies, a.seed, a.steps, a.threads)
	out = {
		'test': {'kernel': a.kernel, 'dtype': a.dtype}, 'reference': {'kernel': ref_kernel, 'dtype': a.ref_dtype},
		'scenario': a.scenario, 'bodies': len(r), 'steps': a.steps, 'seed': a.seed,
		'acceleration_max_rel': _rel(acc, acc0), 'position_max_rel': _rel(r, r0), 'velocity_max_rel': _rel(v, v0),
		}
	sys.stdout.write(json.dumps(out) + '\n')
	return out


if __name__ == '__main__':
	main()
