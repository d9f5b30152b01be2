# -*- coding: utf-8 -*-
"""Host-side mirror of the reference's kernel API.

Restates the contract of /root/reference/src/gravitation/kernel/_base_.py (read for behaviour, not
copied): the three lifecycle states (`_base_.py:31-33`), the point-mass record (`:39-62`) and the
`universe_base` front end (`:64-177`) with identical method names, argument meaning, attribute names
and error types/messages, so that a kernel written against this file is a drop-in for the reference
package and the parity tests read like tests of the reference.

A kernel derives from `universe_base`, must provide `step_stage1`, and may provide `start_kernel`,
`step_stage2` and `stop_kernel`.  Everything else is part of the fixed front end.
"""

STATE_PREINIT = 0
STATE_STARTED = 1
STATE_STOPPED = 2

_STR_TEMPLATE = '{name} | {x:.4e}, {y:.4e}, {z:.4e} | {vx:.4e}, {vy:.4e}, {vz:.4e}'

# Lifecycle rules of the front end as one table: (call, state it must NOT be made in) -> message of the
# SyntaxError the reference raises in that situation (`_base_.py:110-113,124-127,139-142,167-170`).
_LIFECYCLE_ERRORS = {
	('add_object', STATE_STARTED): 'simulation was started',
	('add_object', STATE_STOPPED): 'simulation was stopped',
	('start', STATE_STARTED): 'simulation is running',
	('start', STATE_STOPPED): 'simulation was stopped',
	('step', STATE_PREINIT): 'simulation was not started',
	('step', STATE_STOPPED): 'simulation was stopped',
	('stop', STATE_PREINIT): 'simulation was not started',
	('stop', STATE_STOPPED): 'simulation was stopped before',
	}


class _point_mass:
	"""one body: `_name`, position `_r`, velocity `_v`, acceleration `_a` (3-sequences) and mass `_m`
	(reference `_base_.py:39-45`; kernels may re-bind `_r/_v/_a` to array row views, np2.py:70-75)"""

	def __init__(self, name, r, v, m):
		self._name = name
		self._r = r
		self._v = v
		self._a = [0.0] * len(r)
		self._m = m

	def __str__(self):
		x, y, z = self._r[0], self._r[1], self._r[2]
		vx, vy, vz = self._v[0], self._v[1], self._v[2]
		return _STR_TEMPLATE.format(name = self._name, x = x, y = y, z = z, vx = vx, vy = vy, vz = vz)

	def move(self, T):
		"""base stage 2 for one body (reference `_base_.py:58-62`): the NEW velocity moves the body
		(symplectic Euler), then the acceleration is cleared"""
		for k in range(len(self._r)):
			self._v[k] = self._a[k] * T + self._v[k]
			self._r[k] = self._v[k] * T + self._r[k]
			self._a[k] = 0.0


class universe_base:
	"""kernel base class — derive from it and implement at least `step_stage1`"""

	def __init__(
		self,
		t = 0.0, # simulation start time (s)
		T = 1.0e3, # time step (s)
		G = 6.6740831e-11, # gravitational constant
		scale_m = 1.0, # mass unit scaling (kg -> internal)
		scale_r = 1.0, # length unit scaling (m -> internal)
		dtype = 'float32', # numerical dtype of the kernel
		threads = 1, # degree of parallelism the kernel may use
		**kwargs # scenario / kernel specific extras, kept in `_meta`
		):
		"""fixed front end (reference `_base_.py:69-93`): do not override"""
		self._scale_m, self._scale_r = scale_m, scale_r
		self._t, self._T = t, T
		scale_off = kwargs.pop('scale_off', False)
		# G has units m^3 kg^-1 s^-2, so internal G = G * scale_r^3 / scale_m (`_base_.py:85-88`)
		self._G = G if scale_off else G * (scale_r ** 3) / scale_m
		self._mass_list = []
		self._state = STATE_PREINIT
		self._dtype = dtype
		self._threads = threads
		self._meta = kwargs

	def _allow(self, call):
		"""lifecycle guard: raises the reference's SyntaxError if `call` is not legal in the current state"""
		message = _LIFECYCLE_ERRORS.get((call, self._state))
		if message is not None:
			raise SyntaxError(message)

	def __iter__(self):
		"""fixed front end: iterates the point masses"""
		return (pm for pm in self._mass_list)

	def __len__(self):
		"""fixed front end: number of point masses"""
		return len(self._mass_list)

	def __str__(self):
		return '\n'.join(str(pm) for pm in self._mass_list)

	def add_object(self, **kwargs):
		"""adds one point mass (keywords name, r, v, m); only before `start`.
		Unless `scale_off` is given, r and v are scaled by `scale_r` IN PLACE on the caller's lists and
		m by `scale_m` (reference `_base_.py:107-118`).  Fixed front end."""
		self._allow('add_object')
		if not kwargs.pop('scale_off', False):
			for key in ('r', 'v'):
				kwargs[key][:] = [component * self._scale_r for component in kwargs[key]]
			kwargs['m'] *= self._scale_m
		self._mass_list.append(_point_mass(**kwargs))

	def start(self):
		"""once, after adding objects and before stepping (reference `_base_.py:120-129`)"""
		self._allow('start')
		self._state = STATE_STARTED
		self.start_kernel()

	def start_kernel(self):
		"""kernel hook: allocate / upload / compile"""

	def step(self):
		"""one time step = stage 1, stage 2, stage 3 in this order (reference `_base_.py:136-145`).
		Fixed front end."""
		self._allow('step')
		self.step_stage1()
		self.step_stage2()
		self.step_stage3()

	def step_stage1(self):
		"""kernel hook, mandatory: accelerations of all bodies, O(N^2)"""
		raise NotImplementedError()

	def step_stage2(self):
		"""kernel hook, optional: velocities and positions from accelerations, O(N)"""
		for pm in self._mass_list:
			pm.move(self._T)

	def step_stage3(self):
		"""advance simulation time.  Fixed front end."""
		self._t += self._T

	def stop(self):
		"""once, after stepping (reference `_base_.py:163-172`)"""
		self._allow('stop')
		self._state = STATE_STOPPED
		self.stop_kernel()

	def stop_kernel(self):
		"""kernel hook: release resources"""
