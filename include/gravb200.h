/* gravb200.h — C ABI of the B200-native gravitation kernel (libgravb200.so).
 *
 * This is the drop-in boundary for ONE path of pleiszenburg/gravitation: the per-step hot loop
 *     universe_base.step() -> step_stage1() [O(N^2)] -> step_stage2() [O(N)]
 *     (reference: src/gravitation/kernel/_base_.py:136-161).
 * A kernel module `src/gravitation/kernel/b200.py` binds these entry points with ctypes exactly the
 * way the reference's own native kernels bind theirs (c1a.py:72-83 / c4b.py:86-93 bind
 * `step_stage1(struct univ*)` from _lib1_/lib.c:52 and _lib4_/lib.c:376).  See INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success and a negative
 * GRAVB200_E* code on failure, with a human readable message in gravb200_last_error().  The library
 * never falls back to the CPU: without a usable CUDA device every compute entry point fails.
 *
 * One context = one shard = one GPU.  Multi-GPU runs use one context per GPU (one process per GPU
 * under torchrun, or several contexts in one process); rows are partitioned contiguously and
 * positions are exchanged once per step.
 */
#ifndef GRAVB200_H
#define GRAVB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRAVB200_F32 0 /* reference dtype='float32' (default, _base_.py:75) */
#define GRAVB200_F64 1 /* reference dtype='float64' */

#define GRAVB200_OK 0
#define GRAVB200_EINVAL (-1)  /* bad argument / wrong call order */
#define GRAVB200_ECUDA (-2)   /* CUDA runtime failure (message has the CUDA error string) */
#define GRAVB200_ENCCL (-3)   /* NCCL failure or NCCL not loadable */
#define GRAVB200_ENODEV (-4)  /* no CUDA device */

#define GRAVB200_NCCL_ID_BYTES 128
#define GRAVB200_PEER_BLOB_BYTES 512
#define GRAVB200_XCHG_NCCL 0 /* per-step in-place ncclAllGather of the new positions */
#define GRAVB200_XCHG_PEER 1 /* fused: the sweep's epilogue stores r' into every peer over NVLink + flag barrier */

typedef struct gravb200_ctx gravb200_ctx;

/* Library / device probes (no compute). */
int gravb200_abi_version(void);
int gravb200_device_count(void);
const char* gravb200_last_error(void);

/* Replaces: pc2.py:96-145 `start_kernel` (allocation of device arrays, launch geometry).
 * n_total: number of bodies; dtype: GRAVB200_F32/F64; device: CUDA ordinal;
 * rank/world: this shard's index and the number of shards (world == 1: single GPU);
 * nccl_id: GRAVB200_NCCL_ID_BYTES bytes from gravb200_nccl_unique_id() on rank 0 (NULL iff world == 1). */
int gravb200_ctx_create(int64_t n_total, int dtype, int device, int rank, int world,
                        const void* nccl_id, gravb200_ctx** out);
int gravb200_ctx_destroy(gravb200_ctx* ctx); /* replaces stop_kernel(), _base_.py:174-177 */

/* rank 0 fills `id` (GRAVB200_NCCL_ID_BYTES); the caller distributes it to the other ranks. */
int gravb200_nccl_unique_id(void* id);

/* Replaces: the per-step 3x memcpy_htod of pc2.py:149-151 plus the mass upload of pc2.py:131.
 * r, v: [n_total][3] C-contiguous, m: [n_total], in the context dtype. Caller-owned, copied, never
 * retained. G, T as universe_base holds them (_base_.py:83-88); eps = softening length (reference: 0). */
int gravb200_upload(gravb200_ctx* ctx, const void* r, const void* v, const void* m, double G,
                    double T, double eps);
/* Positions only (a caller that moved bodies on the host between steps, e.g. pc2's host-side stage 2). */
int gravb200_upload_positions(gravb200_ctx* ctx, const void* r);

/* Several shards, each fed by its own host buffer (replaces the same per-step copies, pc2.py:149-151, without
 * every shard re-sending ALL bodies): r_own, v_own = [n_local][3] rows [row0, row0 + n_local) of this shard
 * (v_own may be NULL: positions only); masses, G, T stay as gravb200_upload left them.  The device exchange
 * (NVLink peer copies + flag barrier, or the NCCL all-gather) completes the position array on every shard.
 * COLLECTIVE: every shard of the universe must call it, like gravb200_upload.  Returns when the host buffers
 * may be reused; the device-side exchange is enqueued (NCCL mode with several shards in one host thread:
 * bracket the calls with gravb200_group_begin/end, as for gravb200_exchange). */
int gravb200_upload_rows(gravb200_ctx* ctx, const void* r_own, const void* v_own);
/* The mirror image for reads (pc2.py:160-162): r_own, v_own, a_own = [n_local][3] of this shard's rows only. */
int gravb200_download_rows(gravb200_ctx* ctx, void* r_own, void* v_own, void* a_own);

/* Replaces: step_stage1() (pc2.py:147-162 kernel launch).  Asynchronous.  Computes accelerations of
 * this shard's rows AND, in the same kernel's epilogue, v' and r' into back buffers; front state is
 * untouched, so accelerations can be read between stage1 and stage2 as with every reference kernel. */
int gravb200_stage1(gravb200_ctx* ctx);
/* Replaces: step_stage2() (np2.py:110-115 / pc2.py:164-168).  Commits the back buffers (multi-GPU:
 * after the all-gather of new positions) and blocks until the device is idle. */
int gravb200_stage2(gravb200_ctx* ctx);
/* Enqueue a flag barrier with all peer shards on the context's stream (peer-store mode; a no-op on one shard and in
 * NCCL mode).  Collective: every shard calls it, between steps.  What follows on the stream starts when every shard
 * has got here — bench.py aligns the shards with it before a timed step, so host launch skew between the ranks'
 * processes is not charged to the step. */
int gravb200_peer_barrier(gravb200_ctx* ctx);

/* Multi-GPU only: enqueue the all-gather of the new positions of the pending step without waiting
 * (gravb200_stage2 does it itself if it was not called).  A host thread that drives SEVERAL contexts must
 * bracket their gravb200_exchange calls with gravb200_group_begin/end (ncclGroupStart/End semantics). */
int gravb200_exchange(gravb200_ctx* ctx);
int gravb200_group_begin(void);
int gravb200_group_end(void);

/* Fused position exchange (replaces the NCCL all-gather; no reference counterpart).
 * Each shard publishes a blob (GRAVB200_PEER_BLOB_BYTES: process id, device, raw pointers and CUDA IPC
 * handles of its two position buffers and its flag array); the caller gathers the blobs of all `world`
 * shards in rank order and hands them to every shard.  Shards of the same process use direct peer
 * access, shards of other processes CUDA IPC.  Call before gravb200_upload.  The mode must be the same
 * on all shards: switch only after every shard connected successfully. */
int gravb200_peer_export(gravb200_ctx* ctx, void* blob);
int gravb200_peer_connect(gravb200_ctx* ctx, const void* blobs);
int gravb200_set_exchange_mode(gravb200_ctx* ctx, int mode);
/* k fused steps without host involvement.  One GPU and N <= 32768 (launch-bound): groups of 8 steps are
 * captured once in a CUDA graph and replayed, the remainder is launched step by step. */
int gravb200_steps(gravb200_ctx* ctx, int k);
int gravb200_sync(gravb200_ctx* ctx);

/* Replaces: the 3x memcpy_dtoh of pc2.py:160-162 and the row views of np2.py:70-75.
 * r: [n_total][3] (all bodies); v, a: [n_local][3] rows [row0, row0+n_local) of this shard
 * (world == 1: all bodies).  Any pointer may be NULL. */
int gravb200_download(gravb200_ctx* ctx, void* r, void* v, void* a);
int gravb200_shard(const gravb200_ctx* ctx, int64_t* row0, int64_t* n_local);
/* The row partition gravb200_ctx_create applies, without a context or a device (hosts that lay out their
 * mirrors before creating shards): contiguous slices of ceil(n_total / world) rows, the last slice short
 * (slices are empty only when n_total is small against world).  The rows a shard owns are the
 * rows it integrates; the symmetric sweep's WORK is divided independently of them (equal shares of the flat
 * tile list of the whole universe), so no block alignment is involved. */
int gravb200_partition(int64_t n_total, int dtype, int world, int rank, int64_t* row0, int64_t* n_local);

/* Device-side timings (cudaEvent): ms[0] = last stage1 sweep kernel, ms[1] = last exchange,
 * ms[2] = total of the last gravb200_steps() call, ms[3] = SM clock (MHz) that CTA 0 of the last sweep
 * observed over its lifetime (clock64 / globaltimer), ms[4] = that lifetime in ms; several shards with the
 * symmetric sweep, last step that was followed by stage2 / ended a gravb200_steps call: ms[5] = sweep kernel,
 * ms[6] = integrate kernel (begins by waiting for every shard's sweep; peer loads of the partial sums, peer
 * stores of r' and of the cleared sums), ms[7] = tail wait for every shard's integrate, ms[8] = this shard's
 * share of the universe's tile list in that sweep relative to the equal share (speed-proportional shares: the
 * shards publish items / ns of every sweep and cut the next one accordingly; GRAVB200_BALANCE=0 disables),
 * ms[9] = the part of ms[6] after the wait (the integrate kernel's own work);
 * entries that do not apply are -1; n = capacity of ms. */
int gravb200_timings(gravb200_ctx* ctx, float* ms, int n);

/* Introspection used by bench.py / tests: launch geometry and counters.
 * info[0]=grid, [1]=threads, [2]=i-bodies per thread, [3]=j tile, [4]=stages, [5]=dynamic smem bytes,
 * [6]=kernel launches so far, [7]=SM count, [8]=packed f32x2 (1/0), [9]=resident CTAs per SM,
 * [10]=exchange mode (GRAVB200_XCHG_*), [11]=variant id in use, [12]=1 if the symmetric sweep in use cuts its CTA
 * ranges at chunk granularity (gravb200_set_split). */
int gravb200_info(const gravb200_ctx* ctx, int64_t* info, int n);
/* Force a kernel variant (tests / ncu A-B): variant < 0 restores the automatic choice.  Ids 0 .. count-1 are
 * the ordered sweeps (bit-reproducible), ids 100 + k, k < gravb200_sym_variant_count(dtype), the symmetric sweeps
 * (every unordered pair once, fp64 atomics: reproducible up to fp64 rounding of the cross-tile sum), ids 200 + k,
 * k < gravb200_small_variant_count(), the persistent multi-step kernel for universes that fit one SM's shared
 * memory (one shard, bit-reproducible; the automatic choice up to 9 472 bodies in float32, 4 736 in float64 on 148 SMs —
 * there gravb200_steps(k) is ONE cooperative launch with a grid barrier between the steps). */
int gravb200_set_variant(gravb200_ctx* ctx, int variant);
/* Symmetric sweeps only: granularity of the stream-K cut of the flat (block row, j-tile) list into CTA ranges.
 * mode 0: whole j-tiles (256 / 512 bodies); 1: chunks of 32 j-bodies (two CTAs may share a tile), for the variants
 * built with that twin; -1 (default): chunks when whole tiles would leave the slowest CTA more than 3 % above the
 * average — mid-sized universes and small shards, where one tile more or less is a large part of a CTA's work.
 * Same pairs, same arithmetic per pair; only the grouping of the fp64 atomic adds changes.  With chunks the ranges
 * carry equal COST (a chunk of a diagonal tile, evaluated ordered, weighs 3/4 of a symmetric one in float32, 4/5 in
 * float64 with 8 rows per thread; GRAVB200_SPLIT_WEIGHTED=0: equal chunk counts).  Env GRAVB200_SPLIT=0|1 sets the
 * initial mode. */
int gravb200_set_split(gravb200_ctx* ctx, int mode);
int gravb200_variant_count(int dtype);
int gravb200_sym_variant_count(int dtype);
int gravb200_small_variant_count(void);
/* Geometry the persistent kernel `variant` (>= 200) would run `n_total` bodies with on a device of `sm_count`
 * SMs (needs no device): out[0]=CTAs, [1]=rows per CTA, [2]=row groups of 32 per CTA, [3]=j-slots per slice,
 * [4]=slices, [5]=dynamic shared memory in bytes.  Returns 0 if it fits, 1 if the positions do not fit in
 * shared memory, < 0 on bad arguments or too many rows per CTA. */
int gravb200_small_geometry(int64_t n_total, int dtype, int sm_count, int variant, int64_t* out, int n);
const char* gravb200_variant_name(int dtype, int variant);

/* Raw device pointers of the shard state (for peer access / torch interop in tests).
 * which: 0 = pos front, 1 = pos back, 2 = vel front, 3 = acc. */
void* gravb200_device_ptr(gravb200_ctx* ctx, int which);

/* Page-locked host memory for the caller's mirrors (so uploads/downloads run at full PCIe rate). */
int gravb200_host_alloc(size_t bytes, void** out);
int gravb200_host_free(void* p);

/* FP32/FP64 FMA-chain microbenchmark on `device` (measured non-tensor peak).
 * out[0] = fp32 FFMA TFLOP/s, out[1] = packed FFMA2 TFLOP/s, out[2] = fp64 DFMA TFLOP/s,
 * out[3] = MUFU.RSQ G op/s, out[4] = SM clock MHz observed during the fp32 run. */
int gravb200_peak_probe(int device, double* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* GRAVB200_H */
