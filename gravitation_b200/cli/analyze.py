# -*- coding: utf-8 -*-
"""Benchmark log -> JSON, with the validation rules of the reference's analyze
(/root/reference/src/gravitation/cli/analyze.py:47-108): a log is a concatenation of worker runs, each
starting with `{"log": "START"}`; a run must be pure JSON lines, carry no ERROR, exactly one INPUT and one
SIZE line, a gap-free STEP counter sequence and end with `{"log": "EXIT", "msg": "OK"}`.  Output per run:
`{"meta": INPUT (+ simulation.size), "runtime": [...], "gctime": [...]}`.  Line types this repo adds
(e.g. RATE) are ignored, exactly as the reference's analyze ignores unknown types.

  python -m gravitation_b200.cli.analyze -l benchmark.log -o benchmark.json
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


def main(argv = None):
	ap = argparse.ArgumentParser(description = 'analyze benchmark logfile')
	ap.add_argument('--logfile', '-l', default = 'benchmark.log')
	ap.add_argument('--data', '-o', default = 'benchmark.json')
	a = ap.parse_args(argv)
	with open(a.logfile, 'r') as f:
		runs = parse_log(f.read())
	with open(a.data, 'w') as f:
		f.write(json.dumps(runs, indent = '\t', sort_keys = True))


if __name__ == '__main__':
	main()
