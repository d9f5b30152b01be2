# -*- coding: utf-8 -*-
"""The kernel module as a drop-in for the REFERENCE tree.

Copies /root/reference/src/gravitation to a scratch directory, installs exactly the files INTEGRATION.md
section 1 lists, and drives the module through the reference's OWN `lib/load.py` inventory and `_base_.py`
front end in a fresh interpreter.  Everything up to `start()` must behave like any reference kernel;
`start()` itself needs a B200 and must fail with the library's error — never with an AttributeError /
ImportError of something that only exists in this repository.  Skipped where /root/reference is absent
(the GPU box)."""

import json
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = '/root/reference/src/gravitation'

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_PKG), reason = 'reference tree not present')


def integration_file_list():
	"""(destination relative to src/gravitation/kernel, source relative to the repo root | None for an empty
	file) parsed from the fenced file list of INTEGRATION.md section 1 — the test installs what the document says"""
	text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
	block = text.split('## 1.', 1)[1].split('```', 2)[1]
	out = []
	for line in block.strip().splitlines():
		m = re.match(r'\s*src/gravitation/kernel/(\S+)\s+(?:<-\s+(\S+))?', line)
		assert m, line
		dest, src = m.group(1), m.group(2)
		if src == 'empty':
			src = None
		out.append((dest, src))
	return out


@pytest.fixture(scope = 'module')
def ref_tree(tmp_path_factory, shim):
	base = tmp_path_factory.mktemp('reftree')
	pkg = os.path.join(str(base), 'gravitation')
	shutil.copytree(REF_PKG, pkg)
	for dest, src in integration_file_list():
		target = os.path.join(pkg, 'kernel', dest)
		os.makedirs(os.path.dirname(target), exist_ok = True)
		if src is None:
			open(target, 'w').close()
		elif dest.endswith('.so'):
			os.symlink(shim.LIB_PATH, target) # "built by make": the in-tree build of this checkout
		else:
			shutil.copyfile(os.path.join(ROOT, src), target)
	return str(base)


def run_in_tree(tree, code):
	"""fresh interpreter with ONLY the scratch tree importable (not this repository)"""
	env = {k: v for k, v in os.environ.items() if k != 'PYTHONPATH'}
	prog = 'import sys, json, warnings\nwarnings.simplefilter("ignore")\nsys.path.insert(0, %r)\n%s' % (tree, code)
	out = subprocess.run([sys.executable, '-c', prog], capture_output = True, text = True, cwd = tree, env = env, timeout = 300)
	assert out.returncode == 0, out.stderr[-3000:]
	return json.loads(out.stdout.strip().splitlines()[-1])


def test_file_list_is_what_the_document_says():
	files = dict(integration_file_list())
	assert set(files) == {'b200.py', '_b200_/__init__.py', '_b200_/_shim.py', '_b200_/libgravb200.so'}
	assert files['b200.py'] == 'gravitation_b200/kernel/b200.py'
	assert files['_b200_/_shim.py'] == 'gravitation_b200/_shim.py'


def test_reference_inventory_lists_and_loads_the_kernel(ref_tree):
	res = run_in_tree(ref_tree, '''
from gravitation.lib.load import inventory
assert 'b200' in inventory and '_b200_' not in inventory
k = inventory['b200']
k.load_meta()                      # the reference's AST walker (lib/load.py:77-99, 113-144)
meta = {key: k[key] for key in k.keys()}
k.load_module()                    # importlib, lib/load.py:100-102
cls = k.get_class()
import gravitation.kernel._base_ as base
print(json.dumps(dict(meta = meta, is_sub = issubclass(cls, base.universe_base), module = cls.__module__,
	repo_imported = any(name.startswith('gravitation_b200') for name in sys.modules))))
''')
	assert res['is_sub'] and res['module'] == 'gravitation.kernel.b200'
	assert res['repo_imported'] is False
	assert res['meta']['parallel'] is True and res['meta']['name'] == 'b200'
	assert all(res['meta'][key] is not None for key in (
		'longname', 'version', 'description', 'requirements', 'externalrequirements', 'interpreters', 'license', 'authors'))


def test_front_end_in_the_reference_tree(ref_tree):
	"""add_object / add_objects / lifecycle errors exactly as the reference's base class raises them, and
	start() fails only because there is no GPU here (GravB200Error from the library, no fallback)"""
	res = run_in_tree(ref_tree, '''
import numpy as np
from gravitation.lib.load import inventory
k = inventory['b200']; k.load_meta(); k.load_module()
out = {}
u = k(T = 2.0e12, scale_m = 1.0e-30, scale_r = 1.0e-10, threads = 1)
r = [1.0e20, 2.0e20, 3.0e20]; v = [1.0, 2.0, 3.0]
u.add_object(name = 'a', r = r, v = v, m = 2.0e30)
u.add_object(name = 'b', r = [0.0, 0.0, 0.0], v = [0.0, 0.0, 0.0], m = 4.0e40)
out['scaled_in_place'] = r
out['len'] = len(u)
out['names'] = [pm._name for pm in u]
out['G'] = u._G
def err(f, *a):
	try:
		f(*a)
	except Exception as e:
		return [type(e).__name__, str(e)]
	return None
out['step_before_start'] = err(u.step)
out['steps_before_start'] = err(u.steps, 3)
out['stop_before_start'] = err(u.stop)
out['start'] = err(u.start)
out['state_after_failed_start'] = u._state
out['add_after_start'] = err(lambda: u.add_object(name = 'c', r = [0.0] * 3, v = [0.0] * 3, m = 1.0))

b = k(scale_r = 2.0, scale_m = 3.0)
out['bulk_mixed'] = None
b.add_objects(np.ones((5, 3)), np.zeros((5, 3)), np.ones(5), names = ['x%d' % i for i in range(5)])
out['bulk_len'] = len(b)
out['bulk_r0'] = [float(c) for c in b._mass_list[0]._r]
out['bulk_m0'] = b._mass_list[0]._m
out['bulk_names'] = [pm._name for pm in b]
out['bulk_append'] = err(lambda: b.add_object(name = 'c', r = [0.0] * 3, v = [0.0] * 3, m = 1.0))
out['bulk_twice'] = err(b.add_objects, np.ones((2, 3)), np.zeros((2, 3)), np.ones(2))
out['bulk_shape'] = err(k().add_objects, np.ones((2, 2)), np.zeros((2, 2)), np.ones(2))
out['bulk_start'] = err(b.start)
out['bulk_after_start'] = err(b.add_objects, np.ones((2, 3)), np.zeros((2, 3)), np.ones(2))

w = k(rank = 0, world = 2, nccl_id = b'0' * 128)     # one process per GPU: needs nothing outside _b200_
w.add_object(name = 'a', r = [1.0, 0.0, 0.0], v = [0.0] * 3, m = 1.0)
out['world_start'] = err(w.start)
print(json.dumps(out))
''')
	assert res['scaled_in_place'] == [1.0e10, 2.0e10, 3.0e10]
	assert res['len'] == 2 and res['names'] == ['a', 'b']
	assert abs(res['G'] - 6.6740831e-11) < 1e-24
	assert res['step_before_start'] == ['SyntaxError', 'simulation was not started']
	assert res['steps_before_start'] == ['SyntaxError', 'simulation was not started']
	assert res['stop_before_start'] == ['SyntaxError', 'simulation was not started']
	assert res['start'][0] == 'GravB200Error', res['start']
	assert res['state_after_failed_start'] == 1 # the reference sets STARTED before start_kernel (_base_.py:128-129)
	assert res['add_after_start'] == ['SyntaxError', 'simulation was started']
	assert res['bulk_len'] == 5 and res['bulk_r0'] == [2.0, 2.0, 2.0] and res['bulk_m0'] == 3.0
	assert res['bulk_names'] == ['x%d' % i for i in range(5)]
	assert res['bulk_append'][0] == 'SyntaxError'
	assert res['bulk_twice'] == ['SyntaxError', 'add_objects needs an empty universe']
	assert res['bulk_shape'][0] == 'ValueError'
	assert res['bulk_start'][0] == 'GravB200Error'
	assert res['bulk_after_start'] == ['SyntaxError', 'simulation was started']
	assert res['world_start'][0] == 'GravB200Error', res['world_start']


def test_kernel_module_uses_only_names_the_reference_base_has():
	"""static guard: every name b200.py takes from `_base_` exists in the reference's `_base_.py`, and the
	module imports nothing else from this repository but the binding"""
	import ast
	src = open(os.path.join(ROOT, 'gravitation_b200', 'kernel', 'b200.py')).read()
	ref_src = open(os.path.join(REF_PKG, 'kernel', '_base_.py')).read()
	ref_names = set()
	for node in ast.parse(ref_src).body:
		if isinstance(node, (ast.ClassDef, ast.FunctionDef)):
			ref_names.add(node.name)
		elif isinstance(node, ast.Assign):
			ref_names.update(t.id for t in node.targets if isinstance(t, ast.Name))
	ref_methods = {n.name for c in ast.parse(ref_src).body if isinstance(c, ast.ClassDef) and c.name == 'universe_base'
		for n in c.body if isinstance(n, ast.FunctionDef)}
	relative = []
	for node in ast.walk(ast.parse(src)):
		if isinstance(node, ast.ImportFrom) and node.level > 0:
			relative.append((node.level, node.module, [a.name for a in node.names]))
			if node.module == '_base_':
				assert set(a.name for a in node.names) <= ref_names
	assert sorted(relative) == sorted([(1, '_base_', ['universe_base', '_point_mass', 'STATE_PREINIT', 'STATE_STARTED', 'STATE_STOPPED']),
		(2, None, ['_shim']), (1, '_b200_', ['_shim'])])
	# self.<method>() calls on the base class must exist there (the round-1 break was `self._allow`)
	own = {n.name for c in ast.parse(src).body if isinstance(c, ast.ClassDef) and c.name == 'universe' for n in c.body if isinstance(n, ast.FunctionDef)}
	cls = [c for c in ast.parse(src).body if isinstance(c, ast.ClassDef) and c.name == 'universe'][0]
	for node in ast.walk(cls):
		if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name) and node.func.value.id == 'self':
			name = node.func.attr
			assert name in own or name in ref_methods, 'self.%s() exists neither in b200.universe nor in the reference base class' % name
