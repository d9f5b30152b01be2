# -*- coding: utf-8 -*-
"""One-process-per-GPU plumbing (torchrun): row partition, rendezvous, NCCL id exchange, row gathers.

The data path itself never goes through torch: the per-step position exchange is enqueued by
libgravb200 on its own stream (NCCL all-gather over NVLink).  torch.distributed is used only for the
rendezvous (unique-id broadcast, barriers, max-over-ranks timing) and for gathering per-shard rows
(velocities, accelerations) onto every rank when a caller asks for them.

Row partition = SURVEY.md section 8e: contiguous slices of ceil(N / P) rows,
the last one short; the library is the single source of it."""

import os

import numpy as np


def row_partition(n, world, dtype = 'float32'):
	"""[(row0, n_local)] per rank, asked from the library so that it IS the partition gravb200_ctx_create
	applies (csrc/gravb200.cu shard_chunk: ceil(n / world) rows, the last slice short)"""
	from . import _shim
	return _shim.partition(n, world, dtype)


def env_world():
	"""(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when not launched by it"""
	return (
		int(os.environ.get('RANK', 0)),
		int(os.environ.get('WORLD_SIZE', 1)),
		int(os.environ.get('LOCAL_RANK', 0)),
		)


def init_process_group(backend = None):
	"""joins the torchrun rendezvous if WORLD_SIZE > 1; returns (rank, world, local_rank)"""
	rank, world, local_rank = env_world()
	if world > 1:
		import torch
		import torch.distributed as dist
		if not dist.is_initialized():
			if backend is None:
				backend = 'nccl' if torch.cuda.is_available() else 'gloo'
			if backend == 'nccl':
				torch.cuda.set_device(local_rank)
			dist.init_process_group(backend = backend, rank = rank, world_size = world)
	return rank, world, local_rank


def broadcast_bytes(payload, src = 0):
	"""rank `src` passes bytes, everyone gets them back"""
	import torch.distributed as dist
	box = [payload if dist.get_rank() == src else None]
	dist.broadcast_object_list(box, src = src)
	return box[0]


def make_shard(n, dtype = 'float32'):
	"""this rank's shard of an n-body universe inside the current process group"""
	from . import _shim
	rank, world, local_rank = env_world()
	if world == 1:
		return _shim.Shard(n, dtype, device = local_rank)
	uid = broadcast_bytes(_shim.nccl_unique_id() if rank == 0 else None)
	shard = _shim.Shard(n, dtype, device = local_rank, rank = rank, world = world, nccl_id = uid)
	connect_peers(shard)
	return shard


def connect_peers(shard):
	"""collective switch to the fused peer-store exchange (all ranks or none); lives next to the binding so the
	kernel module needs nothing but `_shim` inside the reference tree.  Returns the mode in use."""
	from . import _shim
	return _shim.connect_peers(shard)


def gather_rows(local_rows, n, dtype = None):
	"""all ranks contribute their (n_local, k) slice and receive the assembled (n, k) array"""
	import torch
	import torch.distributed as dist
	world = dist.get_world_size()
	parts = row_partition(n, world, dtype or ('float64' if local_rows.dtype == np.float64 else 'float32'))
	chunk = max(cnt for _, cnt in parts)
	k = local_rows.shape[1]
	send = np.zeros((chunk, k), dtype = local_rows.dtype)
	send[:local_rows.shape[0], :] = local_rows
	t = torch.from_numpy(send)
	use_cuda = dist.get_backend() == 'nccl'
	if use_cuda:
		t = t.cuda()
	out = [torch.empty_like(t) for _ in range(world)]
	dist.all_gather(out, t)
	full = torch.cat([o[:cnt] for o, (_, cnt) in zip(out, parts)], dim = 0)
	return full.cpu().numpy()


def max_over_ranks(value):
	"""float max across the group (timing rule: a multi-GPU step takes as long as its slowest rank)"""
	import torch
	import torch.distributed as dist
	if not dist.is_initialized():
		return float(value)
	t = torch.tensor([float(value)], dtype = torch.float64)
	if dist.get_backend() == 'nccl':
		t = t.cuda()
	dist.all_reduce(t, op = dist.ReduceOp.MAX)
	return float(t.item())


def barrier():
	import torch.distributed as dist
	if dist.is_initialized():
		dist.barrier()
