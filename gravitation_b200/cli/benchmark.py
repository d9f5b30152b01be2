# -*- coding: utf-8 -*-
"""Benchmark driver: kernels x GPU counts x sizes, one isolated worker subprocess per measurement.

Same sweep semantics as the reference (/root/reference/src/gravitation/cli/benchmark.py:94-104,187-205):
sizes run from 2^a to 2^b with the x1.5 midpoints in between, `threads` is swept only for kernels whose
meta data say `__parallel__ = True` (for the b200 kernel: threads = number of GPUs), every worker's
JSON lines are appended to one log that `analyze` turns into data.  No plotting dependency.

  python -m gravitation_b200.cli.benchmark -k b200 -b 12 16 -p 1 -l benchmark.log
"""

import argparse
import json
import subprocess
import sys
import tempfile

from ..lib.load import inventory


def size_range(start, end):
	"""2^start, midpoint, 2^(start+1), midpoint, ... 2^end (benchmark.py:94-104)"""
	powers = [2 ** i for i in range(start, end + 1)]
	out = []
	for lo, hi in zip(powers[:-1], powers[1:]):
		out.extend([lo, (lo + hi) // 2])
	out.append(powers[-1])
	return out


def worker_command(kernel, bodies, threads, min_iterations, min_total_runtime, scenario_param = None, interpreter = None):
	param = {'stars_len': bodies}
	param.update(scenario_param or {})
	return [
		interpreter or sys.executable, '-m', 'gravitation_b200.cli.worker',
		'--kernel', kernel, '--scenario', 'galaxy', '--scenario_param', json.dumps(param),
		'--min_iterations', str(min_iterations), '--min_total_runtime', str(min_total_runtime),
		'--threads', str(threads),
		]


def main(argv = None):
	ap = argparse.ArgumentParser(description = 'run a benchmark across kernels')
	ap.add_argument('--logfile', '-l', default = 'benchmark.log')
	ap.add_argument('--kernel', '-k', action = 'append', choices = sorted(inventory.keys()))
	ap.add_argument('--n_body_power_boundaries', '-b', type = int, nargs = 2, default = [2, 16])
	ap.add_argument('--min_iterations', '-i', type = int, default = 10)
	ap.add_argument('--min_total_runtime', '-t', type = int, default = 10)
	ap.add_argument('--threads', '-p', type = int, action = 'append')
	ap.add_argument('--scenario_param', default = '{}', help = 'extra scenario parameters, e.g. {"dtype": "float64", "seed": 42}')
	a = ap.parse_args(argv)
	kernels = a.kernel or sorted(inventory.keys())
	extra = json.loads(a.scenario_param)
	best = {}
	with open(a.logfile, 'a') as log:
		for name in kernels:
			inventory[name].load_meta()
			thread_list = (a.threads or [1]) if inventory[name]['parallel'] is True else [1]
			for threads in thread_list:
				for bodies in size_range(*a.n_body_power_boundaries):
					cmd = worker_command(name, bodies, threads, a.min_iterations, a.min_total_runtime, extra)
					# stderr goes to a temp file, not a pipe: a chatty worker (NCCL_DEBUG, CUDA warnings) must never
					# block on a full pipe while this loop is still reading its stdout
					with tempfile.TemporaryFile('w+') as errfile:
						proc = subprocess.Popen(cmd, stdout = subprocess.PIPE, stderr = errfile, text = True)
						for line in proc.stdout:
							log.write(line)
							try:
								msg = json.loads(line)
							except ValueError:
								continue
							if msg.get('log') == 'BEST_TIME':
								best[(name, threads, bodies)] = msg['value']
						proc.wait()
						errfile.seek(0)
						err = errfile.read()
					log.flush()
					ns = best.get((name, threads, bodies))
					print('%s@%d N=%d best=%s ns%s' % (name, threads, bodies, ns, ' [stderr: %s]' % err.strip()[-200:] if proc.returncode else ''))
	return 0


if __name__ == '__main__':
	sys.exit(main())
