// gravitation_b200 — all-pairs gravity sweep + fused symplectic-Euler integrate for sm_100a.
//
// What this restates (reference = pleiszenburg/gravitation, paths relative to its repo root):
//   * stage 1, per-body N x N form:  a_i = G * sum_{j != i} m_j (r_j - r_i) / |r_j - r_i|^3
//     src/gravitation/kernel/pc2.py:59-91 (i == j skipped by index, G applied once at the end),
//     physics spec src/gravitation/kernel/py1.py:57-69.  No softening in the reference; eps2 = 0
//     reproduces it, eps2 > 0 is an extension.
//   * stage 2, array form:           a *= T; v += a; r += v*T   (each op rounded separately)
//     src/gravitation/kernel/np2.py:110-115, pc2.py:164-168.
//
// How it is built for B200 (nothing here is a translation of pc2/pc3):
//   * state is device resident: pos = {x,y,z,m} (float4 / double4), vel, acc; pos/vel are
//     double-buffered so the fused integrate never overwrites what other CTAs still read.
//   * j-tiles are staged in shared memory by TMA 1-D bulk copies (cp.async.bulk + mbarrier
//     complete_tx) through a STAGES-deep full/empty ring; the consumer warps never execute a
//     CTA-wide barrier inside the sweep.
//   * each thread owns R i-bodies in registers; fp32 arithmetic is issued as packed f32x2
//     (FADD2/FMUL2/FFMA2) over PAIRS of i-bodies, the j-body operand is a scalar register that the
//     packed instruction broadcasts, so one LDS.128 (uniform address) feeds R interactions.
//   * accumulation is hierarchical: fp32 inside one j-tile, fp64 across tiles (SURVEY finding 3).
//   * work decomposition is stream-K over the flat (i-block, j-tile) space: every CTA gets the
//     same number of tiles (+-1) whatever N is; i-blocks split between CTAs are combined by the
//     last-arriving contributor in a fixed order (deterministic), which then runs the fused
//     integrate epilogue.  No second launch, no host round trip.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

// A/B switches of the ordered fp32 sweep (scripts/dbg/ab.py builds).  Scalar loads of the i-bodies take every register
// move out of its j-loop (569 -> 487 instructions per unrolled trip of variant 0, see ld_cg_f32) and negating the
// i-positions once per i-block takes out 24 FADD more (-> 469) — and the kernel gets SLOWER: 2.625 -> 2.589 /
// 2.607 / 2.575 T interactions/s at N = 2^20 (profiles/r02_ordered_ab.jsonl).  It is bound by the operand fetch of
// the packed FFMA2, not by issue slots; the copies ptxas makes happen to sit in friendlier
// register banks.  Both stay off; the persistent small-N kernel (two warps per scheduler, issue bound) needs them.
#ifndef ORDERED_SCALAR_LOADS
#define ORDERED_SCALAR_LOADS 0
#endif
#ifndef ORDERED_NEG_ONCE
#define ORDERED_NEG_ONCE 0
#endif

namespace gravb200 {

constexpr int kMaxPeers = 15;   // one 16-GPU NVSwitch domain at most

struct SweepParams {
    const void* pos_front;    // [n_pad] {x,y,z,m} of ALL bodies (read)
    void* pos_back;           // [n_pad] same layout; local rows written by the epilogue
    const void* vel_front;    // [n_local_pad] {vx,vy,vz,0} local rows (read)
    void* vel_back;           // [n_local_pad] (written)
    void* acc;                // [n_local_pad] {ax,ay,az,0} (written)
    double* partial;          // [2*grid][IBLK][4] fp64 partial sums of split i-blocks
    unsigned int* counters;   // [n_iblocks] tiles-arrived counters (self-resetting)
    long long n_total;        // number of j-bodies (all bodies)
    long long row0;           // global index of local row 0
    long long n_local;        // local rows (i-bodies)
    int n_iblocks;            // ceil(n_local / IBLK)
    int n_jtiles;             // ceil(n_total / TILE)
    double G, T;              // gravitational constant, time step
    float eps2_f;             // softening^2 (fp32 kernel)
    double eps2_d;            // softening^2 (fp64 kernel)
    int integrate;            // 1: epilogue also writes vel_back / pos_back
    unsigned long long* clk;  // optional [2]: CTA 0 writes {SM cycles, ns} of its lifetime (clock evidence)
    // fused position exchange (multi-GPU, peer-store mode): the epilogue also stores r' into the back
    // position buffer of every peer GPU through NVLink-mapped pointers (no collective launch)
    int n_peers;
    void* peer_back[kMaxPeers];
};

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (PTX ISA 8.x, sm_90+; SASS: SYNCS.* and UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Waiting on an mbarrier: a non-blocking probe in a spin loop the compiler can see, left on a warp vote.
// The first version spun on the blocking `mbarrier.try_wait` inside one asm block (branches hidden from the
// compiler): lanes of one warp could get their result at different times, left the loop separately and the
// warp stayed split, which sends every following __shfl_sync down its slow path (3-8x slower symmetric
// sweeps, profiles/r01_sym_divergence.md).
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Warp-uniform wait: every lane of a converged warp polls and the warp leaves the loop TOGETHER (vote).
// No lane is ever left behind in a spin loop while the rest of its warp runs on: a warp split that way
// (seen with a single producer lane polling the `empty` barrier) can stay split for the rest of the kernel.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    while (!__all_sync(0xffffffffu, mbar_test(bar, parity) ? 1 : 0)) {}
}
// global -> shared bulk copy, completion signalled on `bar` as transaction bytes
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // one MUFU.RSQ, 2 ulp
    return y;
}
// m * d2^(-3/2) in fp64 without the math library's rsqrt(): MUFU.RSQ64H seed y0 (relative error
// <= 2^-22), residual e = 1 - d2*y0^2, then d2^(-3/2) = y0^3 * (1 - e)^(-3/2) = y0^3 * (1 + 3/2 e +
// 15/8 e^2 + O(e^3)); the dropped term is < 2^-60.  7 FP64-pipe operations, branch free (rsqrt() costs 5
// plus a range-check branch, and the cube and mass 3 more).  d2 = 0 gives NaN, like inf * 0 in the reference.
__device__ __forceinline__ double mass_over_r3(double m, double d2) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d2));
    const double q = y0 * y0;
    const double e = fma(-d2, q, 1.0);
    const double c = fma(fma(e, 1.875, 1.5), e, 1.0);
    return ((m * y0) * q) * c;
}
__device__ __forceinline__ double ld_cg_f64(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// One scalar, never merged into a vector load (asm): coordinates that feed packed f32x2 instructions as the pair
// (x of body 2c, x of body 2c + 1) must sit in an aligned register pair.  Components that arrive as one 128-bit
// load result stay in that register quad and ptxas re-assembles every pair with two MOVs before EVERY use
// (2.6 - 4.75 extra instructions per pair of interactions in the ordered sweeps, profiles/r02_small_n.md).
__device__ __forceinline__ float ld_cg_f32(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ double4 ld_cg_d4(const double4* p) {   // L2-coherent 32-byte load
    const double2 a = __ldcg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// stream-K partition helpers: CTA c owns flat tiles [lo(c), lo(c+1))
__host__ __device__ __forceinline__ long long sk_lo(long long total, long long c, long long S) {
    return total * c / S;
}
// the CTA that owns flat tile t
__host__ __device__ __forceinline__ long long sk_owner(long long total, long long t, long long S) {
    return ((t + 1) * S - 1) / total;
}

// Programmatic dependent launch (one shard, symmetric step): the integrate kernel and the next sweep are launched
// while their predecessor still runs; everything up to griddep_wait() — barrier initialisation, the CTA's range of
// the tile list, the walk to its first tile — overlaps the predecessor's tail, nothing that depends on the
// predecessor's results is touched before.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async_all() {
    asm volatile("fence.proxy.async;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// Cross-GPU hand-over folded INTO the compute kernels of the symmetric multi-shard step (no barrier launches):
// a kernel may begin by waiting until every peer has published `wait_value` in this GPU's flag array
// (peer_wait), and end by publishing `signal_value` into every peer's flag array once ALL its CTAs are done
// (peer_signal: per-kernel arrival counter, the last CTA releases the flags at system scope).  The sweep waits
// for "integrate of the previous step finished everywhere" and signals "my sweep is done"; the integrate
// kernel waits for that and signals "my integrate is done" — so the exchange of step s overlaps whatever the
// peers are still doing, and consecutive steps need two launches each.
// ------------------------------------------------------------------------------------------------
struct PeerSync {
    const unsigned long long* wait_flags;             // local [world], written by the peers; nullptr: no wait
    unsigned long long wait_value;
    unsigned long long* signal_flags[kMaxPeers + 1];  // flag arrays of all ranks (entry [rank] of each is ours)
    unsigned long long signal_value;
    unsigned int* done_ctr;                           // CTAs of this launch that are done (self-resetting); nullptr: no signal
    int rank, world;
    int* error;                                       // set to 1 when a peer does not show up within 60 s
};

// all threads of the CTA call it; returns after a CTA barrier.  Warp 0 polls (lane q watches peer q) and leaves
// the loop on a vote, so no warp is ever split by the wait.
__device__ __forceinline__ void peer_wait(const PeerSync& s) {
    if (threadIdx.x < 32) {
        const int q = threadIdx.x;
        const bool mine = q < s.world && q != s.rank;
        unsigned long long t0, t1, seen;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        bool ok;
        do {
            ok = true;
            if (mine) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(s.wait_flags + q) : "memory");
                ok = seen >= s.wait_value;
                if (!ok) {
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > 60ull * 1000 * 1000 * 1000) { *s.error = 1; ok = true; }
                }
            }
        } while (!__all_sync(0xffffffffu, ok ? 1 : 0));
    }
    __syncthreads();
    fence_proxy_async_all();   // what the peers stored is read by TMA bulk copies (async proxy) too
}

// all threads of the CTA call it, after their last store of the launch
__device__ __forceinline__ void peer_signal(const PeerSync& s) {
    if (s.done_ctr == nullptr) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int old = atomicAdd(s.done_ctr, 1u);
        if (old == gridDim.x - 1) {
            *s.done_ctr = 0;   // the next launch that uses it is stream-ordered behind this one
            __threadfence_system();   // fence + relaxed stores = one release for all peers (a st.release each would fence again)
            for (int q = 0; q < s.world; ++q)
                if (q != s.rank)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(s.signal_flags[q] + s.rank), "l"(s.signal_value) : "memory");
        }
    }
}

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

// ------------------------------------------------------------------------------------------------
// Epilogue shared by both precisions: a = G * sum ; v' = v + a*T ; r' = r + v'*T
// Each op rounded separately in the state dtype (np2.py:110-115) -- intrinsics forbid contraction.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void finalize_body(const SweepParams& p, long long i_local, double sx,
                                              double sy, double sz, float4 ri, float /*tag*/) {
    float ax = (float)(sx * p.G), ay = (float)(sy * p.G), az = (float)(sz * p.G);
    ((float4*)p.acc)[i_local] = make_float4(ax, ay, az, 0.f);
    if (p.integrate) {
        const float T = (float)p.T;
        float4 v = ((const float4*)p.vel_front)[i_local];
        v.x = __fadd_rn(v.x, __fmul_rn(ax, T));
        v.y = __fadd_rn(v.y, __fmul_rn(ay, T));
        v.z = __fadd_rn(v.z, __fmul_rn(az, T));
        ri.x = __fadd_rn(ri.x, __fmul_rn(v.x, T));
        ri.y = __fadd_rn(ri.y, __fmul_rn(v.y, T));
        ri.z = __fadd_rn(ri.z, __fmul_rn(v.z, T));
        ((float4*)p.vel_back)[i_local] = v;
        ((float4*)p.pos_back)[p.row0 + i_local] = ri;
        for (int q = 0; q < p.n_peers; ++q) ((float4*)p.peer_back[q])[p.row0 + i_local] = ri;   // NVLink peer store
    }
}
__device__ __forceinline__ void finalize_body(const SweepParams& p, long long i_local, double sx,
                                              double sy, double sz, double4 ri, double /*tag*/) {
    double ax = __dmul_rn(sx, p.G), ay = __dmul_rn(sy, p.G), az = __dmul_rn(sz, p.G);
    ((double4*)p.acc)[i_local] = make_double4(ax, ay, az, 0.0);
    if (p.integrate) {
        const double T = p.T;
        double4 v = ((const double4*)p.vel_front)[i_local];
        v.x = __dadd_rn(v.x, __dmul_rn(ax, T));
        v.y = __dadd_rn(v.y, __dmul_rn(ay, T));
        v.z = __dadd_rn(v.z, __dmul_rn(az, T));
        ri.x = __dadd_rn(ri.x, __dmul_rn(v.x, T));
        ri.y = __dadd_rn(ri.y, __dmul_rn(v.y, T));
        ri.z = __dadd_rn(ri.z, __dmul_rn(v.z, T));
        ((double4*)p.vel_back)[i_local] = v;
        ((double4*)p.pos_back)[p.row0 + i_local] = ri;
        for (int q = 0; q < p.n_peers; ++q) ((double4*)p.peer_back[q])[p.row0 + i_local] = ri;   // NVLink peer store
    }
}

// ------------------------------------------------------------------------------------------------
// The sweep kernel.
//   REAL    float | double
//   THREADS threads per CTA;  R i-bodies per thread (even for float);  IBLK = THREADS*R rows per
//   i-block;  TILE j-bodies per shared-memory stage;  STAGES ring depth;  MINB min CTAs per SM.
//   PACK    fp32 only: 1 = packed f32x2 over i-body pairs, 0 = scalar FFMA (kept for the ncu A/B).
//   UNROLL  j-bodies per trip of the inner loop.
//   SS      fp32 packed path: 1 = the per-thread fp64 sums live in shared memory instead of registers.
// Shared memory (dynamic): STAGES*TILE*sizeof(vec4) tile ring, 2*STAGES mbarriers, then (SS) 3*R*THREADS doubles.
// ------------------------------------------------------------------------------------------------
template <typename REAL, int THREADS, int R, int TILE, int STAGES, int MINB, int PACK, int UNROLL, int SS>
__global__ void __launch_bounds__(THREADS, MINB) sweep_kernel(const SweepParams p) {
    using V4 = typename Vec4<REAL>::type;
    constexpr int IBLK = THREADS * R;
    constexpr int NWARPS = THREADS / 32;
    constexpr bool F32 = sizeof(REAL) == 4;
    static_assert(!F32 || (R % 2 == 0), "fp32 path packs pairs of i-bodies");
    static_assert(!SS || (F32 && PACK), "shared-memory sums exist for the packed fp32 path only");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    V4* tiles = reinterpret_cast<V4*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * TILE * sizeof(V4));
    uint64_t* empty_bar = full_bar + STAGES;
    double* ssum = reinterpret_cast<double*>(empty_bar + STAGES);   // [3][R][THREADS], SS only
    (void)ssum;
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const long long S = gridDim.x;
    const long long nj = p.n_jtiles;
    const long long total = (long long)p.n_iblocks * nj;
    const long long lo = sk_lo(total, blockIdx.x, S);
    const long long hi = sk_lo(total, blockIdx.x + 1, S);
    if (lo >= hi) return;   // CTA-uniform
    const int ntiles = (int)(hi - lo);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NWARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const V4* __restrict__ posf = reinterpret_cast<const V4*>(p.pos_front);

    unsigned long long clk0 = 0, ns0 = 0;
    if (p.clk && blockIdx.x == 0 && tid == 0) {
        clk0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    }

    // flat tile lo = (i-block ib0, j-tile jt0); both indices are advanced incrementally from here
    // on (no 64-bit division on the per-tile path)
    const int ib0 = (int)(lo / nj);
    const int jt0 = (int)(lo - (long long)ib0 * nj);
    const int njt = p.n_jtiles;

    // producer: one thread issues the TMA bulk copies in flat-tile order into ring slot k % STAGES
    int p_jt = jt0, p_slot = 0;   // only meaningful in thread 0
    auto issue_next = [&]() {
        const long long j0 = (long long)p_jt * TILE;
        long long cnt = p.n_total - j0;
        if (cnt > TILE) cnt = TILE;
        const uint32_t bytes = (uint32_t)(cnt * sizeof(V4));
        mbar_expect_tx(&full_bar[p_slot], bytes);
        tma_bulk_g2s(tiles + (size_t)p_slot * TILE, posf + j0, bytes, &full_bar[p_slot]);
        if (++p_jt == njt) p_jt = 0;
        if (++p_slot == STAGES) p_slot = 0;
    };
    if (tid == 0) {
        const int pre = ntiles < (STAGES - 1) ? ntiles : (STAGES - 1);
        for (int k = 0; k < pre; ++k) issue_next();
    }

    // per-thread i-body state
    REAL xi[R], yi[R], zi[R];
    double sx[R], sy[R], sz[R];
    int ib = ib0, jt = jt0;    // current flat tile
    int seg_tiles = 0;         // tiles accumulated into the current segment
    int seg_index = 0;         // 0 for this CTA's first segment
    bool new_block = true;
    int c_slot = 0;            // consumer ring slot
    uint32_t c_parity = 0;     // parity of the full barrier for the current ring lap
    uint32_t e_parity = 0;     // (thread 0) parity of the empty barrier the producer waits on next
    int e_slot = 0;

    // --- segment completion: combine split i-blocks, then run the fused epilogue ---------------
    auto finish_segment = [&]() {
        const long long t0 = (long long)ib * nj, t1 = t0 + nj;
        bool do_final = true;
        if constexpr (SS) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                sx[r] = ssum[(0 * R + r) * THREADS + tid];
                sy[r] = ssum[(1 * R + r) * THREADS + tid];
                sz[r] = ssum[(2 * R + r) * THREADS + tid];
            }
        }
        if (seg_tiles != njt) {
            // split i-block: publish my partial, last arriver reduces all contributors in order
            const long long slot = 2 * (long long)blockIdx.x + (seg_index == 0 ? 0 : 1);
            double* mine = p.partial + slot * (long long)IBLK * 4;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double4 v = make_double4(sx[r], sy[r], sz[r], 0.0);
                reinterpret_cast<double4*>(mine)[r * THREADS + tid] = v;
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned int mytiles = (unsigned int)seg_tiles;
                const unsigned int old = atomicAdd(&p.counters[ib], mytiles);
                const int last = (old + mytiles == (unsigned int)nj);
                if (last) p.counters[ib] = 0;   // self-reset for the next launch
                s_last = last;
            }
            __syncthreads();
            do_final = (s_last != 0);
            if (do_final) {
                __threadfence();
                const long long c_first = sk_owner(total, t0, S), c_last = sk_owner(total, t1 - 1, S);
#pragma unroll
                for (int r = 0; r < R; ++r) { sx[r] = 0.0; sy[r] = 0.0; sz[r] = 0.0; }
                // contributors in CTA order (fixed order = deterministic sum).  lo(c) is carried from one
                // contributor to the next (one division each), in 32 bits when the products fit.
                const bool small = (unsigned long long)total * (unsigned long long)(S + 1) < 0xffffffffull;
                auto lo_of = [&](long long c) -> long long {
                    return small ? (long long)((unsigned int)total * (unsigned int)c / (unsigned int)S) : sk_lo(total, c, S);
                };
                long long clo = lo_of(c_first);
                for (long long c = c_first; c <= c_last; ++c) {
                    const long long chi = lo_of(c + 1);
                    const bool empty = clo >= chi;
                    const long long cs = 2 * c + ((clo >= t0) ? 0 : 1);
                    clo = chi;
                    if (empty) continue;
                    const double4* src = reinterpret_cast<const double4*>(p.partial + cs * (long long)IBLK * 4);
                    double4 part[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) part[r] = ld_cg_d4(src + r * THREADS + tid);   // all loads in flight first
#pragma unroll
                    for (int r = 0; r < R; ++r) { sx[r] += part[r].x; sy[r] += part[r].y; sz[r] += part[r].z; }
                }
            }
            __syncthreads();   // s_last may be rewritten by the next segment
        }
        if (do_final) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long il = (long long)ib * IBLK + r * THREADS + tid;
                if (il < p.n_local) {
                    V4 ri = posf[p.row0 + il];   // .w = mass (positions equal xi/yi/zi)
                    finalize_body(p, il, sx[r], sy[r], sz[r], ri, REAL(0));
                }
            }
        }
    };

    for (int k = 0; k < ntiles; ++k) {
        if (new_block) {
            new_block = false;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long il = (long long)ib * IBLK + r * THREADS + tid;
                V4 b;
                b.x = 0; b.y = 0; b.z = 0; b.w = 0;
                if constexpr (F32 && PACK && ORDERED_SCALAR_LOADS) {
                    if (il < p.n_local) {   // scalar loads: see ld_cg_f32
                        const float* src = reinterpret_cast<const float*>(posf + p.row0 + il);
                        b.x = ld_cg_f32(src); b.y = ld_cg_f32(src + 1); b.z = ld_cg_f32(src + 2);
                    }
                } else {
                    if (il < p.n_local) b = posf[p.row0 + il];
                }
                if constexpr (F32 && PACK && ORDERED_NEG_ONCE) { xi[r] = -b.x; yi[r] = -b.y; zi[r] = -b.z; }   // packed path: d = r_j + (-r_i), negated once per i-block
                else { xi[r] = b.x; yi[r] = b.y; zi[r] = b.z; }
                if constexpr (SS) {
                    ssum[(0 * R + r) * THREADS + tid] = 0.0;
                    ssum[(1 * R + r) * THREADS + tid] = 0.0;
                    ssum[(2 * R + r) * THREADS + tid] = 0.0;
                } else {
                    sx[r] = 0.0; sy[r] = 0.0; sz[r] = 0.0;
                }
            }
        }
        // producer step: refill the slot freed by tile k-1 with tile k+STAGES-1
        if (tid < 32) {   // the whole first warp waits for the free slot (no lane is left behind polling), one lane issues
            const int kk = k + STAGES - 1;
            if (kk < ntiles) {
                if (kk >= STAGES) {
                    mbar_wait_warp(&empty_bar[e_slot], e_parity);
                    if (++e_slot == STAGES) { e_slot = 0; e_parity ^= 1; }
                }
                if (tid == 0) issue_next();
                __syncwarp();
            }
        }
        const int s = c_slot;
        mbar_wait_warp(&full_bar[s], c_parity);
        if (++c_slot == STAGES) { c_slot = 0; c_parity ^= 1; }
        const V4* __restrict__ tile = tiles + (size_t)s * TILE;

        const long long j0 = (long long)jt * TILE;
        long long cntl = p.n_total - j0;
        const int jn = cntl > TILE ? TILE : (int)cntl;
        const long long ib_g0 = p.row0 + (long long)ib * IBLK;   // global index of the i-block's first row
        const bool special = (jn < TILE) || (j0 < ib_g0 + IBLK && j0 + TILE > ib_g0);

        if constexpr (F32) {
            if constexpr (PACK) {
                constexpr int P = R / 2;
                float2 ax[P], ay[P], az[P];
#pragma unroll
                for (int q = 0; q < P; ++q) ax[q] = ay[q] = az[q] = make_float2(0.f, 0.f);
                const float e2 = p.eps2_f;
                // one interaction of the j-body b with every i-pair of this thread.  MASKED: the self pair is
                // taken out by index — dj = j - (row of this thread's slot 0), slot r is the self pair at
                // dj == r*THREADS; its dx is exactly 0, so a zero factor keeps inf*0 out.  The compares and
                // selects go to the ALU pipe, whose issue slots are idle here.
                auto interact = [&](const float4 b, auto masked, const int dj) {
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        constexpr float sg = ORDERED_NEG_ONCE ? 1.f : -1.f;   // folds away
                        const float2 dx = __fadd2_rn(make_float2(b.x, b.x), make_float2(sg * xi[2 * q], sg * xi[2 * q + 1]));
                        const float2 dy = __fadd2_rn(make_float2(b.y, b.y), make_float2(sg * yi[2 * q], sg * yi[2 * q + 1]));
                        const float2 dz = __fadd2_rn(make_float2(b.z, b.z), make_float2(sg * zi[2 * q], sg * zi[2 * q + 1]));
                        float2 d2 = __ffma2_rn(dx, dx, make_float2(e2, e2));
                        d2 = __ffma2_rn(dy, dy, d2);
                        d2 = __ffma2_rn(dz, dz, d2);
                        const float2 ri = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
                        const float2 ri2 = __fmul2_rn(ri, ri);
                        const float2 mr = __fmul2_rn(make_float2(b.w, b.w), ri);
                        float2 sc = __fmul2_rn(mr, ri2);
                        if constexpr (decltype(masked)::value) {
                            if (dj == (2 * q) * THREADS) sc.x = 0.f;
                            if (dj == (2 * q + 1) * THREADS) sc.y = 0.f;
                        }
                        ax[q] = __ffma2_rn(dx, sc, ax[q]);
                        ay[q] = __ffma2_rn(dy, sc, ay[q]);
                        az[q] = __ffma2_rn(dz, sc, az[q]);
                    }
                };
                using yes = std::true_type;
                using no = std::false_type;
                if (!special) {
#pragma unroll UNROLL
                    for (int j = 0; j < TILE; ++j) interact(tile[j], no{}, 0);
                } else {
                    const int dj0 = (int)(j0 - (ib_g0 + tid));
                    if (jn == TILE) {
                        // diagonal tile: full trip count, unrolled like the fast loop (these tiles cluster in a
                        // few CTAs, so their speed decides the kernel's tail at small N)
#pragma unroll UNROLL
                        for (int j = 0; j < TILE; ++j) interact(tile[j], yes{}, dj0 + j);
                    } else {
                        // ragged last tile: stop at n_total
#pragma unroll 1
                        for (int j = 0; j < jn; ++j) interact(tile[j], yes{}, dj0 + j);
                    }
                }
                if constexpr (SS) {
                    // fp64 sums live in shared memory (thread-private columns): frees 6*R registers for
                    // the scheduler; touched once per tile
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        ssum[(0 * R + 2 * q) * THREADS + tid] += (double)ax[q].x;
                        ssum[(0 * R + 2 * q + 1) * THREADS + tid] += (double)ax[q].y;
                        ssum[(1 * R + 2 * q) * THREADS + tid] += (double)ay[q].x;
                        ssum[(1 * R + 2 * q + 1) * THREADS + tid] += (double)ay[q].y;
                        ssum[(2 * R + 2 * q) * THREADS + tid] += (double)az[q].x;
                        ssum[(2 * R + 2 * q + 1) * THREADS + tid] += (double)az[q].y;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        sx[2 * q] += (double)ax[q].x; sx[2 * q + 1] += (double)ax[q].y;
                        sy[2 * q] += (double)ay[q].x; sy[2 * q + 1] += (double)ay[q].y;
                        sz[2 * q] += (double)az[q].x; sz[2 * q + 1] += (double)az[q].y;
                    }
                }
            } else {
                float ax[R], ay[R], az[R];
#pragma unroll
                for (int r = 0; r < R; ++r) ax[r] = ay[r] = az[r] = 0.f;
                const float e2 = p.eps2_f;
                if (!special) {
#pragma unroll UNROLL
                    for (int j = 0; j < TILE; ++j) {
                        const float4 b = tile[j];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float dx = b.x - xi[r], dy = b.y - yi[r], dz = b.z - zi[r];
                            const float d2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, e2)));
                            const float ri = rsqrt_approx(d2);
                            const float sc = (b.w * ri) * (ri * ri);
                            ax[r] = fmaf(dx, sc, ax[r]);
                            ay[r] = fmaf(dy, sc, ay[r]);
                            az[r] = fmaf(dz, sc, az[r]);
                        }
                    }
                } else {
                    const long long ibase = ib_g0 + tid;
                    for (int j = 0; j < jn; ++j) {
                        const float4 b = tile[j];
                        const long long dj = (j0 + j) - ibase;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float dx = b.x - xi[r], dy = b.y - yi[r], dz = b.z - zi[r];
                            const float d2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, e2)));
                            const float ri = rsqrt_approx(d2);
                            float sc = (b.w * ri) * (ri * ri);
                            if (dj == (long long)r * THREADS) sc = 0.f;
                            ax[r] = fmaf(dx, sc, ax[r]);
                            ay[r] = fmaf(dy, sc, ay[r]);
                            az[r] = fmaf(dz, sc, az[r]);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    sx[r] += (double)ax[r]; sy[r] += (double)ay[r]; sz[r] += (double)az[r];
                }
            }
        } else {
            // fp64: accumulate straight into the fp64 sums
            const double e2 = p.eps2_d;
            if (!special) {
#pragma unroll UNROLL
                for (int j = 0; j < TILE; ++j) {
                    const double4 b = tile[j];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double dx = b.x - xi[r], dy = b.y - yi[r], dz = b.z - zi[r];
                        const double d2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, e2)));
                        const double sc = mass_over_r3(b.w, d2);
                        sx[r] = fma(dx, sc, sx[r]);
                        sy[r] = fma(dy, sc, sy[r]);
                        sz[r] = fma(dz, sc, sz[r]);
                    }
                }
            } else {
                const long long ibase = ib_g0 + tid;
                for (int j = 0; j < jn; ++j) {
                    const double4 b = tile[j];
                    const long long dj = (j0 + j) - ibase;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double dx = b.x - xi[r], dy = b.y - yi[r], dz = b.z - zi[r];
                        const double d2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, e2)));
                        double sc = mass_over_r3(b.w, d2);
                        if (dj == (long long)r * THREADS) sc = 0.0;
                        sx[r] = fma(dx, sc, sx[r]);
                        sy[r] = fma(dy, sc, sy[r]);
                        sz[r] = fma(dz, sc, sz[r]);
                    }
                }
            }
        }
        // consumer release: this warp is done reading ring slot s
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]);

        ++seg_tiles;
        if (++jt == njt || k == ntiles - 1) {
            finish_segment();
            jt = 0; ++ib;
            seg_tiles = 0; ++seg_index;
            new_block = true;
        }
    }
    if (p.clk && blockIdx.x == 0 && tid == 0) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        p.clk[0] = clock64() - clk0;
        p.clk[1] = ns1 - ns0;
    }
}

template <typename REAL, int THREADS, int R, int TILE, int STAGES, int SS>
constexpr size_t sweep_smem_bytes() {
    return (size_t)STAGES * TILE * sizeof(typename Vec4<REAL>::type) + 2 * STAGES * sizeof(uint64_t) +
           (SS ? (size_t)3 * R * THREADS * sizeof(double) : 0);
}

// ------------------------------------------------------------------------------------------------
// Cross-GPU step barrier of the peer-store exchange.  Runs on the sweep's stream right after it, so all
// of this GPU's peer stores are ordered before the flag release.  Lane q tells peer q "rank `rank` has
// finished step `step`" (release, system scope) and then waits until peer q has said the same here.
// One barrier per step covers both hazards: new positions have landed (RAW) and nobody still reads the
// buffer that the next step will overwrite (WAR).  A clock-based timeout turns a dead peer into an
// error flag instead of a hang.
// ------------------------------------------------------------------------------------------------
struct BarrierParams {
    unsigned long long* my_flags;                 // [world] written by the peers
    unsigned long long* peer_flags[kMaxPeers + 1];   // [world] each rank's flag array (entry `rank` unused)
    int rank, world;
    unsigned long long step;
    unsigned long long timeout_ns;
    int* error;                                   // set to 1 on timeout
};

__global__ void exchange_barrier_kernel(const BarrierParams b) {
    const int q = threadIdx.x;
    if (q >= b.world || q == b.rank) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(b.peer_flags[q] + b.rank), "l"(b.step) : "memory");
    unsigned long long t0, t1, seen;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(b.my_flags + q) : "memory");
        if (seen >= b.step) return;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < b.timeout_ns);
    *b.error = 1;
}

// Tail of a symmetric multi-shard step when the host wants the state complete (stage2, last step of steps(k)):
// every peer has finished its integrate kernel, i.e. all r' have landed here.
__global__ void peer_wait_kernel(const PeerSync s) { peer_wait(s); }

}  // namespace gravb200
