#!/usr/bin/env python3
"""build_pc2.py — TEST/BENCH INFRASTRUCTURE (dev container only).

Builds the reference's OWN GPU kernel for B200 so bench.py can report it next to ours: the CUDA-C source
string `SM` of /root/reference/src/gravitation/kernel/pc2.py:59-91 is taken from the file as it lies there
(parsed with `ast`, pycuda is not importable here), formatted exactly as `pc2.py:133-139` formats it
(dtype, itype, rsqrt, G, MASS_LEN are baked in as literals) and compiled with nvcc for sm_100a.  Only the
compiled cubins are kept, in the git-ignored oracle/_ref/; no reference source enters the repository.

  python oracle/build_pc2.py            -> oracle/_ref/pc2_float32_<N>.cubin for N in 2^12, 2^16, 2^20
"""
import ast
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/src/gravitation/kernel/pc2.py'
OUT = os.path.join(HERE, '_ref')
G_GALAXY = 6.6740831e-11 # the uniform benchmark universe runs with scale_off (bench.py)


def reference_source_string():
	tree = ast.parse(open(REF).read())
	for node in tree.body:
		if isinstance(node, ast.Assign) and getattr(node.targets[0], 'id', None) == 'SM':
			return ast.literal_eval(node.value)
	raise RuntimeError('SM not found in pc2.py')


def build(n, dtype = 'float32', G = G_GALAXY):
	src = reference_source_string().format( # pc2.py:133-139
		dtype = {'float32': 'float', 'float64': 'double'}[dtype],
		itype = 'int32',
		rsqrt = {'float32': 'rsqrtf', 'float64': 'rsqrt'}[dtype],
		G = G,
		MASS_LEN = n,
		)
	os.makedirs(OUT, exist_ok = True)
	out = os.path.join(OUT, 'pc2_%s_%d.cubin' % (dtype, n))
	with tempfile.TemporaryDirectory() as tmp:
		cu = os.path.join(tmp, 'pc2.cu')
		with open(cu, 'w') as f:
			f.write('#include <stdint.h>\nextern "C" {\n' + src + '\n}\n') # pycuda's SourceModule wraps in extern "C" too
		subprocess.run(['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-cubin', '-o', out, cu], check = True)
	return out


if __name__ == '__main__':
	if not os.path.isfile(REF):
		print('reference not present: keeping prebuilt cubins (if any)')
		sys.exit(0)
	for log2n in (12, 16, 20):
		print(build(1 << log2n))
