# -*- coding: utf-8 -*-
"""Isolated single-kernel benchmark worker — the timed loop around the hot path.

Keeps the JSON-lines protocol of the reference's worker (/root/reference/src/gravitation/cli/worker.py:
115-245) byte-compatible, so its `analyze`/`plot` keep working on our logs: one JSON object per line with
`log` in START, INPUT, PROCEDURE, SIZE, STEP{runtime, gctime, counter}, BEST_TIME{value}, ERROR, EXIT.
Timing semantics are the reference's: garbage collection disabled for the run, `gc.collect()` before and
(timed separately) after every step, wall-clock ns of `universe.step()` only, minimum over steps
(`worker.py:119-136,216-245`).  Additional keys (interactions/s) ride on extra `log` types that the
reference's analyze ignores.

  python -m gravitation_b200.cli.worker --kernel b200 --scenario_param '{"stars_len": 65536}' -p 1
"""

import argparse
import gc
import json
import platform
import sys
import traceback

from ..lib.load import inventory
from ..lib.simulation import create_simulation, store_simulation
from ..lib.timing import best_run_timer, elapsed_timer


def _msg(**d):
	sys.stdout.write(json.dumps(d) + '\n')
	sys.stdout.flush()


def _bail():
	_msg(log = 'ERROR', msg = traceback.format_exc())
	_msg(log = 'EXIT', msg = 'BAD')
	sys.exit()


def worker(kernel, scenario, scenario_param, data_out_file, save_after_iteration, min_iterations, min_total_runtime, threads):
	_msg(log = 'START')
	scenario_param = json.loads(scenario_param)
	counter = [0]
	_msg(
		log = 'INPUT',
		simulation = dict(
			kernel = kernel, scenario = scenario, scenario_param = scenario_param,
			min_iterations = min_iterations, min_total_runtime = min_total_runtime, threads = threads,
			),
		python = dict(
			build = list(platform.python_build()), compiler = platform.python_compiler(),
			implementation = platform.python_implementation(), version = list(sys.version_info),
			),
		platform = dict(
			system = platform.system(), release = platform.release(), version = platform.version(),
			machine = platform.machine(), processor = platform.processor(),
			),
		)
	min_total_runtime_ns = min_total_runtime * 10 ** 9
	inventory[kernel].load_module()

	_msg(log = 'PROCEDURE', msg = 'Creating simulation ...')
	try:
		s = create_simulation(
			scenario = scenario, universe_class = inventory[kernel].get_class(),
			scenario_param = scenario_param, threads = threads,
			)
	except Exception:
		_bail()
	_msg(log = 'PROCEDURE', msg = 'Simulation created.')
	_msg(log = 'SIZE', value = len(s))

	rt, gt, et = best_run_timer(), best_run_timer(), elapsed_timer()

	def _store():
		_msg(log = 'PROCEDURE', msg = 'Saving data after step %d ...' % counter[0])
		try:
			store_simulation(s, data_out_file, 'kernel=%s;len=%d;step=%d' % (kernel, len(s), counter[0]))
		except Exception:
			_bail()
		_msg(log = 'PROCEDURE', msg = 'Data saved after step %d.' % counter[0])

	def _step():
		try:
			gc.collect()
			rt.start()
			s.step()
			rt_ = rt.stop()
			gt.start()
			gc.collect()
			gt_ = gt.stop()
		except Exception:
			_bail()
		counter[0] += 1
		if counter[0] in save_after_iteration:
			_store()
		_msg(log = 'STEP', runtime = rt_, gctime = gt_, counter = counter[0])
		_msg(log = 'BEST_TIME', value = rt.min())

	gc.disable()
	if 0 in save_after_iteration:
		_store()
	for _ in range(min_iterations):
		_step()
	elapsed = et()
	if elapsed < min_total_runtime_ns:
		_msg(log = 'PROCEDURE', msg = 'Extra steps required.')
		remaining = min_total_runtime_ns - elapsed
		for _ in range(remaining // elapsed * min_iterations):
			_step()
	else:
		_msg(log = 'PROCEDURE', msg = 'Minimum steps sufficient.')
	n = len(s)
	_msg(log = 'RATE', interactions_per_step = n * (n - 1), best_interactions_per_s = n * (n - 1) / (rt.min() * 1e-9))
	_msg(log = 'EXIT', msg = 'OK')


def main(argv = None):
	ap = argparse.ArgumentParser(description = 'isolated single-kernel benchmark worker')
	ap.add_argument('--kernel', '-k', required = True, choices = sorted(inventory.keys()))
	ap.add_argument('--scenario', default = 'galaxy')
	ap.add_argument('--scenario_param', default = '{}', help = 'JSON string with scenario parameters')
	ap.add_argument('--data_out_file', '-o', default = 'data.h5')
	ap.add_argument('--save_after_iteration', '-s', type = int, action = 'append', default = [])
	ap.add_argument('--min_iterations', '-i', type = int, default = 10)
	ap.add_argument('--min_total_runtime', '-t', type = int, default = 10)
	ap.add_argument('--threads', '-p', type = int, default = 1, help = 'b200 kernel: number of GPUs')
	a = ap.parse_args(argv)
	worker(a.kernel, a.scenario, a.scenario_param, a.data_out_file, a.save_after_iteration,
		a.min_iterations, a.min_total_runtime, a.threads)


if __name__ == '__main__':
	main()
