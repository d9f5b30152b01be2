# -*- coding: utf-8 -*-
"""Exhaustive interleaving check of the deferred j-side combine of the symmetric sweeps
(gravitation_b200/csrc/nbody_sym.cuh, `combine_pending`).

compute-sanitizer's racecheck does not model mbarrier arrive/wait as synchronisation (it reports the
intended write -> arrive -> wait -> read order as a potential hazard, profiles/r01_sym_divergence.md), so the
protocol is checked here instead: a small model of W warps walking a sequence of tiles, every interleaving
of their events explored, with the two properties the kernel needs:

  * a combine of tile k reads, from buffer k % 2, exactly the partials every warp wrote for tile k
    (nothing missing, nothing already overwritten by tile k + 2),
  * nobody deadlocks,
with the mbarrier modelled as the hardware provides it: an arrival COUNTER that completes a phase every W
arrivals (whoever makes them) and parity waits — so phase overrun / parity aliasing would show up as well.

Per warp and SYMMETRIC tile k the kernel does, in program order:
    ring round of chunk 0                      (no shared-memory effect)
    if a combine is pending: wait(jbar)  ->  read buffer of the pending tile (all warps' rows)
    write own row of buffer k % 2              (chunk 0 ... last chunk)
    arrive(jbar)                               -> tile k is now pending, buffers flip
per DIAGONAL tile (no j-partials):
    if a combine is pending: wait(jbar)  ->  read
and after the last tile the pending combine once more.  The model also runs two broken variants (write before
the wait; a single buffer) to show that the checker finds the hazards the real protocol avoids."""

import itertools

import pytest


def warp_program(tiles, write_before_wait = False, buffers = 2):
	"""events of ONE warp for the tile sequence `tiles` ('s' symmetric / 'd' diagonal)"""
	prog, pending, nbuf, n_sym = [], None, 0, 0
	for kind in tiles:
		if kind == 's':
			k, buf = n_sym, nbuf
			combine = [('wait', pending[0]), ('read', pending[0], pending[1])] if pending is not None else []
			write = [('write', k, buf)]
			prog += (write + combine) if write_before_wait else (combine + write)
			prog.append(('arrive', k))
			pending = (k, buf)
			nbuf = (nbuf + 1) % buffers
			n_sym += 1
		else:
			if pending is not None:
				prog += [('wait', pending[0]), ('read', pending[0], pending[1])]
				pending = None
	if pending is not None:
		prog += [('wait', pending[0]), ('read', pending[0], pending[1])]
	return prog


def explore(tiles, warps = 3, **kw):
	"""every interleaving (DFS over program counters, memoised); returns None or a description of the first
	violation.  State: program counters; the shared state (buffer contents, arrivals) is a function of them."""
	prog = warp_program(tiles, **kw)
	n = len(prog)

	def shared(pcs):
		content, arrivals = {}, 0
		for w, pc in enumerate(pcs):
			for ev in prog[:pc]:
				if ev[0] == 'write':
					content[(ev[2], w)] = ev[1] # buffer row of warp w now holds tile ev[1]
				elif ev[0] == 'arrive':
					arrivals += 1 # the mbarrier counts arrivals, whoever makes them and for whichever tile
		return content, arrivals

	# what the hardware gives: an arrival counter that completes a phase every `warps` arrivals, and
	# `test_wait.parity p`, true when the current phase's parity differs from p.  A warp's p is its own
	# count of finished waits modulo 2 (`j_parity`).
	waits_before = []
	for pc in range(n + 1):
		waits_before.append(sum(1 for ev in prog[:pc] if ev[0] == 'wait'))

	seen, stack = set(), [tuple([0] * warps)]
	while stack:
		pcs = stack.pop()
		if pcs in seen:
			continue
		seen.add(pcs)
		if all(pc == n for pc in pcs):
			continue
		content, arrivals = shared(pcs)
		moved = False
		for w, pc in enumerate(pcs):
			if pc == n:
				continue
			ev = prog[pc]
			if ev[0] == 'wait' and (arrivals // warps) % 2 == waits_before[pc] % 2:
				continue # blocked: the phase with this warp's parity has not completed
			if ev[0] == 'read':
				for v in range(warps):
					if content.get((ev[2], v)) != ev[1]:
						return 'warp %d combines tile %d from buffer %d, but warp %d\'s row holds tile %r (state %r)' % (
							w, ev[1], ev[2], v, content.get((ev[2], v)), pcs)
			moved = True
			stack.append(pcs[:w] + (pc + 1,) + pcs[w + 1:])
		if not moved:
			return 'deadlock at %r' % (pcs,)
	return None


SEQUENCES = ['s', 'ss', 'sss', 'ssss', 'sssss', 'ds', 'dss', 'sds', 'ssds', 'sdsds', 'dsssd', 'ssdss', 'sddss', 'dd', 'ssssd']


@pytest.mark.parametrize('warps', (2, 3))
def test_deferred_combine_is_race_free_in_every_interleaving(warps):
	for tiles in SEQUENCES:
		assert explore(tiles, warps = warps) is None, tiles
	# all sequences of four tiles, so no transition is missed
	for tiles in itertools.product('sd', repeat = 4):
		assert explore(''.join(tiles), warps = warps) is None, tiles


def test_the_checker_finds_the_hazards_the_protocol_avoids():
	# chunk 0's partials written BEFORE the wait: tile k + 2 can overwrite rows a slow warp has not combined yet
	assert 'holds tile' in explore('ssss', warps = 2, write_before_wait = True)
	# a single buffer cannot hold tile k + 1 while tile k is still being combined
	assert 'holds tile' in explore('sss', warps = 2, buffers = 1)


# ---- the TMA tile ring (full / empty mbarriers, producer = first warp), same exhaustive treatment ---------

def ring_programs(ntiles, warps, stages = 3, producer_skips_empty_wait = False):
	"""per-thread event lists: warps 0..W-1 (warp 0 also produces, as in the kernels) and the TMA engine.
	Mirrors the sweep loops: prefetch of STAGES-1 tiles, then per tile k: [producer: wait empty slot of tile
	k+STAGES-1 if it was used before, issue that tile] ; wait full(k) ; read slot ; arrive empty."""
	progs = [[] for _ in range(warps)]
	issued = []
	for k in range(min(ntiles, stages - 1)):
		progs[0].append(('issue', k))
		issued.append(k)
	for k in range(ntiles):
		kk = k + stages - 1
		if kk < ntiles:
			if kk >= stages and not producer_skips_empty_wait:
				progs[0].append(('wait_empty', kk % stages, ((kk - stages) // stages) % 2))
			progs[0].append(('issue', kk))
			issued.append(kk)
		for w in range(warps):
			progs[w] += [('wait_full', k % stages, (k // stages) % 2), ('read', k % stages, k), ('arrive_empty', k % stages)]
	progs.append([('land', k) for k in issued]) # the copy engine completes the bulk copies in issue order, whenever
	return progs


def explore_ring(ntiles, warps = 2, stages = 3, **kw):
	progs = ring_programs(ntiles, warps, stages, **kw)
	lens = [len(p) for p in progs]
	seen, stack = set(), [tuple([0] * len(progs))]
	while stack:
		pcs = stack.pop()
		if pcs in seen:
			continue
		seen.add(pcs)
		if all(pc == n for pc, n in zip(pcs, lens)):
			continue
		# shared state from the executed prefixes
		issued, content = set(), {}
		full = [0] * stages # completed phases of full[s] (one per landed copy)
		empty_arrivals = [0] * stages
		for t, pc in enumerate(pcs):
			for ev in progs[t][:pc]:
				if ev[0] == 'issue':
					issued.add(ev[1])
				elif ev[0] == 'land':
					content[ev[1] % stages] = ev[1]; full[ev[1] % stages] += 1
				elif ev[0] == 'arrive_empty':
					empty_arrivals[ev[1]] += 1
		moved = False
		for t, pc in enumerate(pcs):
			if pc == lens[t]:
				continue
			ev = progs[t][pc]
			if ev[0] == 'land' and ev[1] not in issued:
				continue
			if ev[0] == 'wait_full' and full[ev[1]] % 2 == ev[2]:
				continue
			if ev[0] == 'wait_empty' and (empty_arrivals[ev[1]] // warps) % 2 == ev[2]:
				continue
			if ev[0] == 'read' and content.get(ev[1]) != ev[2]:
				return 'warp %d reads tile %d from slot %d, which holds tile %r (state %r)' % (t, ev[2], ev[1], content.get(ev[1]), pcs)
			moved = True
			stack.append(pcs[:t] + (pc + 1,) + pcs[t + 1:])
		if not moved:
			return 'deadlock at %r' % (pcs,)
	return None


@pytest.mark.parametrize('warps', (1, 2))
def test_tile_ring_is_race_free_in_every_interleaving(warps):
	for ntiles in (1, 2, 3, 4, 5, 7, 8):
		assert explore_ring(ntiles, warps = warps) is None, ntiles


def test_the_ring_checker_finds_a_producer_that_does_not_wait():
	assert 'holds tile' in explore_ring(6, warps = 2, producer_skips_empty_wait = True)


# ---- the multi-GPU step: peer stores of r' + one flag barrier per step (csrc: exchange_barrier_kernel) -------

def rank_program(rank, ranks, steps, symmetric, skip_second_barrier = False):
	"""events of one GPU.  Ordered sweep: read front (whole sweep), store own rows of r' into the back buffer of
	every GPU (epilogue, over NVLink), barrier.  Symmetric sweep: accumulate partial sums for ALL bodies into the
	own accumulator, barrier, integrate (read every GPU's partial sums of the own rows, store r' everywhere),
	barrier, clear the own accumulator."""
	prog, epoch = [], 0
	def barrier():
		nonlocal epoch
		epoch += 1
		for peer in range(ranks):
			prog.append(('flag', peer, epoch)) # st.release.sys into peer's flag array (own slot)
		prog.append(('wait', epoch)) # ld.acquire.sys until every slot of the own array shows this epoch
	for s in range(steps):
		front, back = s % 2, (s + 1) % 2
		prog.append(('read_pos', front, s))
		if symmetric:
			prog.append(('acc_fill', s)) # own accumulator now holds the partial sums of step s
			prog.append(('read_pos', front, s)) # ... the sweep read positions until here
			barrier()
			for peer in range(ranks):
				prog.append(('read_acc', peer, s)) # integrate: partial sums of my rows from every GPU
			for peer in range(ranks):
				prog.append(('store_pos', peer, back, s + 1))
			if not skip_second_barrier:
				barrier()
			prog.append(('acc_clear', s))
		else:
			for peer in range(ranks):
				prog.append(('store_pos', peer, back, s + 1))
			prog.append(('read_pos', front, s)) # the sweep reads the front buffer until its last tile
			barrier()
	return prog


def explore_ranks(ranks, steps, symmetric, **kw):
	progs = [rank_program(r, ranks, steps, symmetric, **kw) for r in range(ranks)]
	lens = [len(p) for p in progs]
	seen, stack = set(), [tuple([0] * ranks)]
	while stack:
		pcs = stack.pop()
		if pcs in seen:
			continue
		seen.add(pcs)
		if all(pc == n for pc, n in zip(pcs, lens)):
			continue
		pos = {(g, b, src): 0 for g in range(ranks) for b in range(2) for src in range(ranks)} # upload: both buffers
		flags = {(g, src): 0 for g in range(ranks) for src in range(ranks)}
		acc = {g: ('clear', -1) for g in range(ranks)}
		for r, pc in enumerate(pcs):
			for ev in progs[r][:pc]:
				if ev[0] == 'store_pos':
					pos[(ev[1], ev[2], r)] = ev[3]
				elif ev[0] == 'flag':
					flags[(ev[1], r)] = ev[2]
				elif ev[0] == 'acc_fill':
					acc[r] = ('full', ev[1])
				elif ev[0] == 'acc_clear':
					acc[r] = ('clear', ev[1])
		moved = False
		for r, pc in enumerate(pcs):
			if pc == lens[r]:
				continue
			ev = progs[r][pc]
			if ev[0] == 'wait' and any(flags[(r, src)] < ev[1] for src in range(ranks)):
				continue
			if ev[0] == 'read_pos':
				for src in range(ranks):
					if pos[(r, ev[1], src)] != ev[2]:
						return 'GPU %d sweeps step %d over buffer %d whose rows of GPU %d are at step %d (state %r)' % (r, ev[2], ev[1], src, pos[(r, ev[1], src)], pcs)
			if ev[0] == 'read_acc' and acc[ev[1]] != ('full', ev[2]):
				return 'GPU %d integrates step %d with the accumulator of GPU %d in state %r (state %r)' % (r, ev[2], ev[1], acc[ev[1]], pcs)
			if ev[0] == 'acc_fill' and acc[r][0] != 'clear':
				return 'GPU %d accumulates step %d into an accumulator that was not cleared' % (r, ev[1])
			moved = True
			stack.append(pcs[:r] + (pc + 1,) + pcs[r + 1:])
		if not moved:
			return 'deadlock at %r' % (pcs,)
	return None


@pytest.mark.parametrize('symmetric', (False, True))
def test_multi_gpu_step_is_race_free_in_every_interleaving(symmetric):
	assert explore_ranks(2, 4, symmetric) is None
	assert explore_ranks(3, 2, symmetric) is None


def test_the_rank_checker_finds_a_missing_barrier():
	# without the barrier after the integrate kernel a fast GPU clears (and refills) partial sums a peer still needs
	assert explore_ranks(2, 3, True, skip_second_barrier = True) is not None
