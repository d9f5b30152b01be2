# -*- coding: utf-8 -*-
"""Benchmark log -> JSON, with the validation rules of the reference's analyze
(/root/reference/src/gravitation/cli/analyze.py:47-108): a log is a concatenation of worker runs, each
starting with `{"log": "START"}`; a run must be pure JSON lines, carry no ERROR, exactly one INPUT and one
SIZE line, a gap-free STEP counter sequence and end with `{"log": "EXIT", "msg": "OK"}`.  Output per run:
`{"meta": INPUT (+ simulation.size), "runtime": [...], "gctime": [...]}`.  Line types this repo adds
(e.g. RATE) are ignored, exactly as the reference's analyze ignores unknown types.

  python -m gravitation_b200.cli.analyze -l benchmark.log -o benchmark.json
  python -m gravitation_b200.cli.analyze -l benchmark.log --summary      # table on stdout as well

`--summary` adds what the reference's plot cannot show (/root/reference/TODO.md:10, cli/plot.py:67-72 only
knows time per step): one row per (kernel, dtype, threads = GPUs, N) with the best step time, the rate in
G body-interactions/s (N(N-1) ordered interactions per step) and that rate as a
fraction of the non-tensor FMA peak by the 20-FLOP convention (`--peak_tflops`, default: nominal B200,
74.45 fp32 / 37.22 fp64, times the number of GPUs).
"""

import argparse
import copy
import json

START_LINE = '{"log": "START"}\n'


def parse_run(text):
	lines = []
	for raw in text.split('\n'):
		if raw.strip() == '':
			continue
		try:
			lines.append(json.loads(raw))
		except json.decoder.JSONDecodeError:
			raise SyntaxError('benchmark log has non-JSON components, likely errors')
	def of(kind):
		return [ln for ln in lines if ln.get('log') == kind]
	if of('ERROR'):
		raise SyntaxError('benchmark has errors')
	inputs, sizes, steps = of('INPUT'), of('SIZE'), of('STEP')
	if len(inputs) > 1:
		raise SyntaxError('more than one INPUT log per benchmark worker run')
	if len(inputs) < 1:
		raise SyntaxError('INPUT log missing in benchmark worker run')
	if len(sizes) > 1:
		raise SyntaxError('more than one SIZE log per benchmark worker run')
	if len(sizes) < 1:
		raise SyntaxError('SIZE log missing in benchmark worker run')
	meta = copy.deepcopy(inputs[0])
	meta.pop('log')
	meta['simulation']['size'] = sizes[0]['value']
	counters = [ln['counter'] for ln in steps]
	if not counters:
		raise SyntaxError('benchmark did not run any steps')
	if counters != list(range(counters[0], counters[0] + len(counters))):
		raise SyntaxError('benchmark has unexpected sequence of steps')
	if lines[-1] != {'log': 'EXIT', 'msg': 'OK'}:
		raise SyntaxError('benchmark did not exit properly')
	return {'meta': meta, 'runtime': [ln['runtime'] for ln in steps], 'gctime': [ln['gctime'] for ln in steps]}


def parse_log(text):
	return [parse_run(chunk) for chunk in text.split(START_LINE) if chunk.strip() != '']


FLOP_PER_INTERACTION = 20.0
NOMINAL_PEAK_TFLOPS = {'float32': 74.45, 'float64': 37.22} # 148 SMs x 128 (64) lanes x 2 x 1.965 GHz


def summarize(runs, peak_tflops = None):
	"""one row per run: kernel, dtype, threads, bodies, best step time, rate, fraction of the FMA peak"""
	rows = []
	for run in runs:
		sim = run['meta']['simulation']
		n = int(sim['size'])
		dtype = sim.get('scenario_param', {}).get('dtype', 'float32')
		threads = int(sim.get('threads', 1))
		best_ns = min(run['runtime'])
		rate = n * (n - 1) / (best_ns * 1e-9)
		peak = (peak_tflops if peak_tflops is not None else NOMINAL_PEAK_TFLOPS.get(dtype, NOMINAL_PEAK_TFLOPS['float32'])) * threads
		rows.append({
			'kernel': sim['kernel'], 'dtype': dtype, 'threads': threads, 'bodies': n, 'steps': len(run['runtime']),
			'best_s': best_ns * 1e-9, 'g_interactions_per_s': rate / 1e9,
			'fraction_of_peak': rate * FLOP_PER_INTERACTION / (peak * 1e12),
			})
	rows.sort(key = lambda r: (r['kernel'], r['dtype'], r['threads'], r['bodies']))
	return rows


def format_summary(rows):
	head = '%-8s %-8s %7s %10s %6s %12s %16s %9s' % ('kernel', 'dtype', 'threads', 'bodies', 'steps', 'best [s]', 'G interactions/s', '% of peak')
	lines = [head, '-' * len(head)]
	for r in rows:
		lines.append('%-8s %-8s %7d %10d %6d %12.6f %16.2f %9.2f' % (
			r['kernel'], r['dtype'], r['threads'], r['bodies'], r['steps'], r['best_s'], r['g_interactions_per_s'], 100.0 * r['fraction_of_peak']))
	return '\n'.join(lines)


def main(argv = None):
	ap = argparse.ArgumentParser(description = 'analyze benchmark logfile')
	ap.add_argument('--logfile', '-l', default = 'benchmark.log')
	ap.add_argument('--data', '-o', default = 'benchmark.json')
	ap.add_argument('--summary', action = 'store_true', help = 'print kernel x dtype x threads x N table with interactions/s and %% of peak')
	ap.add_argument('--peak_tflops', type = float, default = None, help = 'per-GPU FMA peak the fraction refers to (default: nominal B200)')
	a = ap.parse_args(argv)
	with open(a.logfile, 'r') as f:
		runs = parse_log(f.read())
	with open(a.data, 'w') as f:
		f.write(json.dumps(runs, indent = '\t', sort_keys = True))
	if a.summary:
		rows = summarize(runs, a.peak_tflops)
		with open(a.data + '.summary.json', 'w') as f:
			f.write(json.dumps(rows, indent = '\t', sort_keys = True))
		print(format_summary(rows))


if __name__ == '__main__':
	main()
