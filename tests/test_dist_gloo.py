# -*- coding: utf-8 -*-
"""world_size-2 `gloo` tests (CPU) of the one-process-per-GPU host logic in gravitation_b200/dist.py:
rendezvous, id broadcast, row partition, row gathers, max-over-ranks timing."""
import os
import socket

import numpy as np
import pytest


def _free_port():
	s = socket.socket()
	s.bind(('127.0.0.1', 0))
	port = s.getsockname()[1]
	s.close()
	return port


def _worker(rank, world, port, n, out_dir):
	os.environ.update(RANK = str(rank), WORLD_SIZE = str(world), LOCAL_RANK = str(rank),
		MASTER_ADDR = '127.0.0.1', MASTER_PORT = str(port))
	import torch.distributed as tdist
	from gravitation_b200 import dist
	r, w, lr = dist.init_process_group(backend = 'gloo')
	assert (r, w, lr) == (rank, world, rank)
	# the NCCL unique id travels as opaque bytes from rank 0
	payload = dist.broadcast_bytes(bytes(range(128)) if rank == 0 else None)
	assert payload == bytes(range(128))
	# every rank owns a contiguous slice; gathering the slices rebuilds the whole array on every rank
	full = np.arange(n * 3, dtype = np.float32).reshape(n, 3)
	row0, cnt = dist.row_partition(n, world)[rank]
	got = dist.gather_rows(full[row0:row0 + cnt], n)
	assert np.array_equal(got, full)
	assert dist.max_over_ranks(float(rank + 1)) == float(world)
	dist.barrier()
	np.save(os.path.join(out_dir, 'ok%d.npy' % rank), np.array([row0, cnt]))
	tdist.destroy_process_group()


@pytest.mark.parametrize('n', (10, 7, 1, 1 << 18)) # 2^18: a size the symmetric sweep runs on several shards
def test_two_rank_gloo_plumbing(n, tmp_path, shim):
	import torch.multiprocessing as mp
	port = _free_port()
	mp.spawn(_worker, args = (2, port, n, str(tmp_path)), nprocs = 2, join = True)
	parts = [np.load(str(tmp_path / ('ok%d.npy' % k))) for k in range(2)]
	assert parts[0][0] == 0 and parts[0][1] + parts[1][1] == n
