# -*- coding: utf-8 -*-
"""Kernel inventory with the reference's auto-detection contract (`lib/load.py:40-107`):

  * every entry of `kernel/` whose name does not start with `_` is a kernel — `name.py` or a package
    `name/__init__.py` (`lib/load.py:44-51`);
  * meta data are nine module-level dunder assignments of literals, read from the SOURCE with `ast`
    so that listing kernels never imports (and never needs the dependencies of) any of them
    (`lib/load.py:77-99`);
  * the module is imported lazily and must expose a class named `universe` (`lib/load.py:66,76`).

Written against `ast.literal_eval` instead of the reference's per-node walker (which relies on
`ast.Str/Num/NameConstant`, removed in Python 3.14)."""

import ast
import importlib
import os

META_KEYS = (
	'longname', 'version', 'description', 'requirements', 'externalrequirements',
	'interpreters', 'parallel', 'license', 'authors',
	)

_KERNEL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'kernel')
_KERNEL_PKG = __name__.rsplit('.', 2)[0] + '.kernel'


def read_meta(src):
	"""{key: literal} for the `__key__ = literal` assignments at module top level; missing keys -> None"""
	wanted = {'__%s__' % key: key for key in META_KEYS}
	meta = {key: None for key in META_KEYS}
	for node in ast.parse(src).body:
		if not isinstance(node, ast.Assign):
			continue
		for target in node.targets:
			if isinstance(target, ast.Name) and target.id in wanted:
				meta[wanted[target.id]] = ast.literal_eval(node.value)
	return meta


class _kernel:
	"""lazy descriptor of one kernel: meta data without import, module/class on demand"""

	def __init__(self, path, name, isfile):
		self._path, self._name, self._isfile = path, name, isfile
		self._module = None
		self._meta = None

	def _source_file(self):
		if self._isfile:
			return os.path.join(self._path, self._name + '.py')
		return os.path.join(self._path, self._name, '__init__.py')

	def load_meta(self):
		with open(self._source_file(), 'r') as f:
			self._meta = read_meta(f.read())
		self._meta['name'] = self._name

	def load_module(self):
		self._module = importlib.import_module('%s.%s' % (_KERNEL_PKG, self._name))

	def get_class(self):
		if self._module is None:
			raise SyntaxError('kernel module has not been loaded')
		return self._module.universe

	def __call__(self, *args, **kwargs):
		return self.get_class()(*args, **kwargs)

	def __getitem__(self, key):
		if self._meta is None:
			raise SyntaxError('kernel metadata has not been loaded')
		return self._meta[key]

	def keys(self):
		if self._meta is None:
			raise SyntaxError('kernel metadata has not been loaded')
		return self._meta.keys()


class _inventory(dict):
	"""name -> `_kernel` for everything in `kernel/` that does not start with an underscore"""

	def __init__(self, path = _KERNEL_DIR):
		super().__init__()
		for item in sorted(os.listdir(path)):
			if item.startswith('_'):
				continue
			full = os.path.join(path, item)
			if item.lower().endswith('.py'):
				self[item[:-3]] = _kernel(path, item[:-3], True)
			elif os.path.isdir(full) and os.path.isfile(os.path.join(full, '__init__.py')):
				self[item] = _kernel(path, item, False)


inventory = _inventory()
