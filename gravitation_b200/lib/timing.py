# -*- coding: utf-8 -*-
"""Wall-clock timers with the semantics of the reference's `lib/timing.py:41-99`: nanosecond integers
from `time.time_ns`, a start/stop timer that remembers every run (best = min), and an elapsed timer.

The reference notes that "clocking GPU time is tricky" (`lib/timing.py:43`) and never synchronises a
device; the b200 kernel's `step_stage2` blocks until the device is idle, so these host timers measure
the real step."""

import time


class best_run_timer:
	"""start()/stop() pairs; stop() returns the run time in ns; min()/avg()/sum() over all runs"""

	def __init__(self):
		self._runs = []
		self._t0 = None

	def start(self):
		self._t0 = time.time_ns()

	def stop(self):
		dt = time.time_ns() - self._t0
		self._runs.append(dt)
		return dt

	def min(self):
		return min(self._runs)

	def avg(self):
		return sum(self._runs) // len(self._runs)

	def sum(self):
		return sum(self._runs)

	def __len__(self):
		return len(self._runs)


class elapsed_timer:
	"""callable: ns since construction"""

	def __init__(self):
		self._t0 = time.time_ns()

	def __call__(self):
		return time.time_ns() - self._t0
