// gravitation_b200 — symmetric (Newton's third law) fp32 sweep for sm_100a.
//
// The ordered sweep (nbody_kernels.cuh) evaluates every ordered pair: 12 packed FP32 operations + 2 MUFU
// per pair of interactions, and it is bound by the operand fetch of the three accumulate FFMA2
// (DESIGN.md section 3.1).  Here every UNORDERED pair {i, j} is evaluated once and applied to both bodies:
//     d = r_j - r_i ;  w = |d|^-3 ;  a_i += m_j w d ;  a_j -= m_i w d
// = 16 packed operations per pair of unordered pairs, i.e. 8 instead of 12 per ordered interaction — the
// same trick the reference's CPU kernels use (py1.py:57-64, _lib1_/lib.c:75-124, _lib4_/lib.c:192-333).
//
// The j side needs a sum over the i-bodies of different lanes.  Instead of a shuffle-reduction per j-body,
// the j-bodies travel: each lane holds ONE j-body of a 32-body chunk together with its packed partial
// acceleration, interacts it with its R register-resident i-bodies, and passes both on to the next lane
// (10 SHFL per step); after 32 steps every j-body is back home with its complete partial sum over the
// warp's 32*R i-bodies (north_star (c): warp shuffles deliver the j-tile).  Per warp and step: 256
// unordered pairs for 64 packed FP32 instructions + 8 MUFU + 10 SHFL (R = 8).
//
// Work decomposition: body-blocks of IBLK rows.  Block row I evaluates its own diagonal block ordered
// (self pair masked by index, as in the ordered kernel) and, symmetrically, the column blocks I+1 .. I+H
// (cyclically, H = half of the other blocks), so every block row carries the same amount of work whatever
// rank owns it and every unordered pair of blocks is visited exactly once.  The flat list of (row, tile)
// items is cut stream-K style into equal ranges, one per CTA — at whole tiles (SPLIT = 0, the large universes) or
// at 32-body chunks of equal cost (SPLIT = 1 twins: mid-sized universes and small shards, where a CTA holds only
// a few tiles; sym_cta_range / sym_locate_weighted).  Partial accelerations of both sides are
// added into a global fp64 accumulator with RED.ADD.F64; `integrate_kernel` (O(N)) turns the accumulator
// into a, v', r'.  The fp64 sums of fp32 tile partials are exact unless the partials of one body span
// more than 2^29 in magnitude, so results are reproducible up to that rounding, not by construction.
//
// Synchronisation: there is no CTA-wide barrier in the sweep loop.  j-tiles arrive through the TMA/mbarrier
// ring; the j-side combine of a tile is deferred into the next tile behind an mbarrier (combine_pending);
// every wait is warp-uniform (mbar_wait_warp) and the producer is the whole first warp, because a warp that
// shuffles must never split (profiles/r01_sym_divergence.md).  -DSYM_DEBUG builds count diverged warps and
// the cycles spent waiting (scripts/dbg/sym_dbg2.py).
#pragma once

#include "nbody_kernels.cuh"

namespace gravb200 {

#ifdef SYM_DEBUG   // dev builds: count warps that reach a point with fewer than 32 converged lanes
#define SYM_DIVCHK(i)                                                                                      \
    do {                                                                                                   \
        const unsigned am_ = __activemask();                                                               \
        if (p.clk && am_ != 0xffffffffu && (int)(threadIdx.x & 31) == __ffs(am_) - 1) atomicAdd(p.clk + 2036 + (i), 1ull); \
    } while (0)
#else
#define SYM_DIVCHK(i) do {} while (0)
#endif

// Several shards: the shares of the flat list follow the MEASURED speed of each GPU.  Every shard's last sweep
// publishes (items, nanoseconds) to all shards; the next sweep cuts the list in proportion to items / ns — computed
// on the device by every CTA's first thread from the same numbers with the same arithmetic, so all shards agree
// on the boundaries without a host round trip.  The GPUs of one box differ by ~0.5 % under this load, and a step
// takes as long as its slowest sweep (profiles/r02_bench_n8*.json: rank 0's sweep 37.77 / 37.94 ms on two boxes).
struct SymBalance {
    const unsigned long long* stats;                  // local [3][kMaxPeers + 1]: items, ns per shard (published speeds) and, [2][rank], the share of the last sweep; nullptr: equal shares
    unsigned long long* peer_stats[kMaxPeers + 1];    // that array on every shard (entries [rank] are ours to write)
    unsigned long long* tstart;                       // local: earliest CTA start of this launch
};

struct SymParams {
    const float4* pos_front;      // [n_pad] {x,y,z,m} of all bodies (fp32 kernel)
    const double4* pos_front_d;   // same, fp64 kernel
    double* acc64;                // [n_pad][4] fp64 accumulators, zero on entry
    const long long* row_start;   // [n_iblocks + 1] flat tile offset of every local block row
    long long n_total, row0, n_local;
    int n_iblocks;                // local block rows
    int n_gblocks;                // Bt: global body-blocks = ceil(n_total / IBLK)
    int gblock0;                  // global index of local block row 0 (row0 / IBLK, row0 % IBLK == 0)
    float eps2_f;
    double eps2_d;
    unsigned long long* clk;
    // this launch's share of the flat (row, tile) list: [item_lo, item_hi).  One shard: everything.  Several
    // shards: the GLOBAL list (n_iblocks = n_gblocks, gblock0 = 0, row0 = 0, n_local = n_total) is cut into one
    // equal range per shard — stream-K across GPUs, whatever the row ownership of the integrate step is.
    long long item_lo, item_hi;
    // chunk-granular CTA ranges (SPLIT twins) cut by COST: a chunk of a diagonal tile (evaluated ordered) is cheaper
    // than a chunk of a symmetric tile.  row_cost[n_iblocks + 1] = cost prefix per block row (a row's diagonal tiles
    // come first), [cost_lo, cost_hi) = this launch's share in cost units; nullptr: every chunk counts the same.
    const long long* row_cost;
    long long cost_lo, cost_hi;
    int w_sym, w_diag;
    PeerSync sync;                // several shards: hand-over with the peers' integrate kernels (nbody_kernels.cuh)
    SymBalance bal;               // several shards, long sweeps: speed-proportional shares instead of [item_lo, item_hi)
};

// Shard `me`'s share [lo, hi) of `total` flat items from the published (items, ns) of every shard's previous sweep
// (stats[q], stats[kMaxPeers + 1 + q]); any missing entry keeps the equal shares [eq_lo, eq_hi).  Every shard runs
// the same additions in the same order, so hi of shard r and lo of shard r + 1 are the same number: the shares
// tile [0, total) without gaps or overlaps (tests/native/sym_schedule_check.cu).  Two passes, no array.
__host__ __device__ inline void sym_share_bounds(long long total, int P, int me, const unsigned long long* stats,
                                                 long long eq_lo, long long eq_hi, long long& lo, long long& hi) {
    auto weight = [&](int q) -> double {
        const unsigned long long items = stats[q], ns = stats[(kMaxPeers + 1) + q];
        return (items == 0 || ns == 0) ? 0.0 : (double)items / (double)ns;
    };
    double sum = 0.0, c0 = 0.0;
    bool ok = true;
    for (int q = 0; q < P; ++q) {
        const double wq = weight(q);
        if (wq == 0.0) ok = false;
        if (q == me) c0 = sum;
        sum += wq;
    }
    lo = eq_lo; hi = eq_hi;
    if (!ok) return;
    const double c1 = c0 + weight(me);
    lo = me == 0 ? 0 : (long long)((double)total * (c0 / sum));
    hi = me == P - 1 ? total : (long long)((double)total * (c1 / sum));
    if (lo > total) lo = total;
    if (hi > total) hi = total;
    if (hi < lo) hi = lo;
}

// this launch's share of the flat list; all threads call it (CTA barrier inside when balancing)
__device__ __forceinline__ void sym_share(const SymParams& p, long long& share_lo, long long& share_hi) {
    share_lo = p.item_lo; share_hi = p.item_hi;
    if (p.bal.stats == nullptr) return;
    __shared__ long long s_share[2];
    if (threadIdx.x == 0) {
        long long lo, hi;
        sym_share_bounds(p.row_start[p.n_iblocks], p.sync.world, p.sync.rank, p.bal.stats, p.item_lo, p.item_hi, lo, hi);
        s_share[0] = lo; s_share[1] = hi;
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMin(p.bal.tstart, now);
    }
    __syncthreads();
    share_lo = s_share[0]; share_hi = s_share[1];
}

// end of a sweep CTA: peer_signal plus, from the last CTA, this sweep's (items, ns) for everybody's next shares
__device__ __forceinline__ void sym_finish(const SymParams& p, long long share_items) {
    const PeerSync& s = p.sync;
    if (s.done_ctr == nullptr) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int old = atomicAdd(s.done_ctr, 1u);
        if (old == gridDim.x - 1) {
            *s.done_ctr = 0;
            if (p.bal.stats) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                const unsigned long long t0 = atomicExch(p.bal.tstart, ~0ull);   // earliest start of this launch; reset for the next
                const unsigned long long ns = now > t0 ? now - t0 : 0ull;
                // published speed = 10^6 items per `best` ns, best = the FASTEST sweep seen so far, relaxed by 0.05 % per
                // step: a hiccup inside one sweep (they only ever add time: +1 % on single steps, r02k_bench_n2_*.json)
                // must not move work away from a healthy GPU, a GPU that really slows down is followed within ~10 steps
                unsigned long long best = 0;
                if (share_items > 0 && ns > 0) {
                    best = ns * 1000000ull / (unsigned long long)share_items;
                    const unsigned long long prev = p.bal.stats[(kMaxPeers + 1) + s.rank];
                    if (prev) best = min(best, prev + (prev >> 11) + 1);
                }
                for (int q = 0; q < s.world; ++q) {
                    p.bal.peer_stats[q][s.rank] = best ? 1000000ull : 0ull;
                    p.bal.peer_stats[q][(kMaxPeers + 1) + s.rank] = best;
                }
                p.bal.peer_stats[s.rank][2 * (kMaxPeers + 1) + s.rank] = (unsigned long long)share_items;   // local: the share this sweep had
            }
            __threadfence_system();
            for (int q = 0; q < s.world; ++q)
                if (q != s.rank)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(s.signal_flags[q] + s.rank), "l"(s.signal_value) : "memory");
        }
    }
}

// geometry helpers shared by host and device ---------------------------------------------------------
__host__ __device__ inline int sym_tiles_in_block(long long n_total, int iblk, int tile, int K) {
    long long cnt = n_total - (long long)K * iblk;
    if (cnt > iblk) cnt = iblk;
    return (int)((cnt + tile - 1) / tile);
}
// column blocks of block row Ig, the diagonal block included
__host__ __device__ inline int sym_ncols(int Bt, int Ig) {
    const int H = (Bt - 1) / 2;
    const int extra = ((Bt & 1) == 0 && Ig < Bt / 2) ? 1 : 0;
    return 1 + H + extra;
}
__host__ __device__ inline long long sym_row_tiles(long long n_total, int iblk, int tile, int Bt, int Ig) {
    const int nc = sym_ncols(Bt, Ig);
    const int D = iblk / tile;
    long long t = (long long)nc * D;
    const int dist_last = ((Bt - 1) - Ig + Bt) % Bt;   // where the (possibly short) last block sits in this row
    if (dist_last < nc) t -= D - sym_tiles_in_block(n_total, iblk, tile, Bt - 1);
    return t;
}

// position in the flat item list
struct SymWalker {
    int I;   // local block row
    int c;   // column block within the row (0 = diagonal)
    int t;   // tile within the column block
};

template <int IBLK, int TILE>
__host__ __device__ __forceinline__ void sym_advance(SymWalker& w, const SymParams& p) {   // host: tests/native/sym_schedule_check.cu
    const int Ig = p.gblock0 + w.I;
    const int K = (Ig + w.c) % p.n_gblocks;
    if (++w.t == sym_tiles_in_block(p.n_total, IBLK, TILE, K)) {
        w.t = 0;
        if (++w.c == sym_ncols(p.n_gblocks, Ig)) { w.c = 0; ++w.I; }
    }
}

// A CTA's range of the flat list.  SPLIT = 0: whole tiles, [lo, hi).  SPLIT = 1: the share is cut at CHUNK
// granularity (32 j-bodies, the unit of the ring) — tiles [lo, hi), of which the first one starts at chunk
// c_first and the last one ends before chunk c_last; two CTAs may then work on disjoint chunk ranges of the same
// tile.  Mid-sized universes have only a few tiles per CTA (10.26 at N = 2^16 on 148 CTAs: the slowest CTA
// carried 11), with chunks the imbalance is 1 / CHUNKS of that.  (tests/native/sym_schedule_check.cu replays it.)
struct SymRange {
    long long lo, hi;
    int c_first, c_last;
};
template <int CHUNKS, int SPLIT>
__host__ __device__ __forceinline__ bool sym_cta_range(long long share_lo, long long total, long long b, long long S, SymRange& rg) {
    if (SPLIT) {
        const long long tc = total * CHUNKS;
        const long long lo_c = sk_lo(tc, b, S), hi_c = sk_lo(tc, b + 1, S);
        if (lo_c >= hi_c) return false;
        rg.lo = share_lo + lo_c / CHUNKS;
        rg.c_first = (int)(lo_c % CHUNKS);
        rg.hi = share_lo + (hi_c + CHUNKS - 1) / CHUNKS;
        const int rem = (int)(hi_c % CHUNKS);
        rg.c_last = rem ? rem : CHUNKS;
        return true;
    }
    rg.lo = share_lo + sk_lo(total, b, S);
    rg.hi = share_lo + sk_lo(total, b + 1, S);
    rg.c_first = 0; rg.c_last = CHUNKS;
    return rg.lo < rg.hi;
}

// Cost of the flat list up to the start of tile `t` of block row `row` (tiles counted within the row; the row's
// diagonal tiles come first), on top of row_cost[row].  Host (set-up of cost_lo / cost_hi) and device.
__host__ __device__ inline long long sym_cost_in_row(long long n_total, int iblk, int tile, int Ig, long long tiles_into_row, int w_sym, int w_diag) {
    const long long nd = sym_tiles_in_block(n_total, iblk, tile, Ig);
    const int ch = tile / 32;
    return tiles_into_row < nd ? tiles_into_row * ch * w_diag : nd * ch * w_diag + (tiles_into_row - nd) * ch * w_sym;
}

// Cost-weighted chunk-granular range of CTA `b` of `S` and the walker position of its first tile.  The share
// [cost_lo, cost_hi) is cut into S equal cost intervals; a boundary position P maps to the first chunk whose cost
// interval starts at or after P — the same function for the end of CTA b and the start of CTA b + 1, so the ranges
// tile the share's chunks exactly once (tests/native/sym_schedule_check.cu).  The two searches run interleaved in
// one loop: their loads overlap, the prologue pays one dependent chain as before.
template <int IBLK, int TILE>
__host__ __device__ inline bool sym_locate_weighted(const SymParams& p, long long b, long long S, SymRange& rg, SymWalker& w0) {
    constexpr int CH = TILE / 32;
    const long long span = p.cost_hi - p.cost_lo;
    const long long P0 = p.cost_lo + sk_lo(span, b, S), P1 = p.cost_lo + sk_lo(span, b + 1, S);
    if (P0 >= P1) return false;
    int a0 = 0, b0 = p.n_iblocks, a1 = 0, b1 = p.n_iblocks;   // row_cost[a] <= P < row_cost[b]
    while (b0 - a0 > 1 || b1 - a1 > 1) {
        const int m0 = (a0 + b0) >> 1, m1 = (a1 + b1) >> 1;
        const long long v0 = p.row_cost[m0], v1 = p.row_cost[m1];
        if (b0 - a0 > 1) { if (v0 <= P0) a0 = m0; else b0 = m0; }
        if (b1 - a1 > 1) { if (v1 <= P1) a1 = m1; else b1 = m1; }
    }
    const long long rc0 = p.row_cost[a0], rc1 = p.row_cost[a1], rs0 = p.row_start[a0], rs1 = p.row_start[a1], rs0n = p.row_start[a0 + 1];
    auto chunk_in_row = [&](int row, long long r) -> long long {   // r = cost into the row; first chunk starting at or after it
        const long long nd = (long long)sym_tiles_in_block(p.n_total, IBLK, TILE, p.gblock0 + row) * CH;
        return r <= nd * p.w_diag ? (r + p.w_diag - 1) / p.w_diag : nd + (r - nd * p.w_diag + p.w_sym - 1) / p.w_sym;
    };
    const long long q0 = chunk_in_row(a0, P0 - rc0), q1 = chunk_in_row(a1, P1 - rc1);
    const long long t0 = rs0 + q0 / CH, t1 = rs1 + q1 / CH;
    const int c0 = (int)(q0 % CH), c1 = (int)(q1 % CH);
    if (t0 == t1 && c0 == c1) return false;
    rg.lo = t0; rg.c_first = c0;
    rg.hi = c1 ? t1 + 1 : t1; rg.c_last = c1 ? c1 : CH;
    // walker of tile t0: row a0, or the start of the next row when the boundary fell behind the row's last chunk
    long long rem = q0 / CH;
    int a = a0;
    if (rs0 + rem == rs0n) { ++a; rem = 0; }
    w0.I = a; w0.c = 0;
    const int Ig = p.gblock0 + a;
    for (;;) {
        const int tb = sym_tiles_in_block(p.n_total, IBLK, TILE, (Ig + w0.c) % p.n_gblocks);
        if (rem < tb) break;
        rem -= tb; ++w0.c;
    }
    w0.t = (int)rem;
    return true;
}

// ------------------------------------------------------------------------------------------------
// THREADS threads, R i-bodies per thread (even), TILE j-bodies per TMA stage (multiple of 32), STAGES.
// Dynamic shared memory: tile ring | mbarriers (full, empty, jbar) | fp64 i-sums [3][R][THREADS] | j-partials [2][NWARPS][3][TILE].
// ------------------------------------------------------------------------------------------------
template <int THREADS, int R, int TILE, int STAGES, int UNROLL, int SPLIT = 0>
__global__ void __launch_bounds__(THREADS, 1) sym_sweep_kernel(const SymParams p) {
    constexpr int IBLK = THREADS * R;
    constexpr int NWARPS = THREADS / 32;
    constexpr int P = R / 2;
    constexpr int CHUNKS = TILE / 32;
    static_assert(R % 2 == 0 && TILE % 32 == 0 && IBLK % TILE == 0, "geometry");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* tiles = reinterpret_cast<float4*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * TILE * sizeof(float4));
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* jbar = empty_bar + STAGES;                                           // j-partials of a tile complete
    double* ssum = reinterpret_cast<double*>(jbar + 1);                            // [3][R][THREADS]
    float* jpart = reinterpret_cast<float*>(ssum + (size_t)3 * R * THREADS);       // [2][NWARPS][3][TILE]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NWARPS); }
        mbar_init(jbar, NWARPS);
        mbar_fence_init();
    }
    __syncthreads();
    // several shards: the peers' integrate kernels of the previous step have stored their r' into this GPU's
    // position buffer and zeroed their rows of this GPU's accumulator — only then may this sweep read / add
    if (p.sync.wait_flags) peer_wait(p.sync);
    long long share_lo, share_hi;
    sym_share(p, share_lo, share_hi);
    const long long total = share_hi - share_lo;
    const long long S = gridDim.x;
    SymRange rg;
    SymWalker w0;
    bool located = false;   // range AND first-tile walker already known (cost-weighted cut of the SPLIT twin)
    if constexpr (SPLIT != 0) {
        if (p.row_cost) {
            if (!sym_locate_weighted<IBLK, TILE>(p, blockIdx.x, S, rg, w0)) { sym_finish(p, total); return; }
            located = true;
        }
    }
    if (!located && !sym_cta_range<CHUNKS, SPLIT>(share_lo, total, blockIdx.x, S, rg)) { sym_finish(p, total); return; }   // CTA-uniform; an idle CTA still counts as done
    const long long lo = rg.lo;
    const int ntiles = (int)(rg.hi - rg.lo);

    unsigned long long clk0 = 0, ns0 = 0;
    if (p.clk && tid == 0) {   // every CTA: start/end time stamps (debug: distribution of CTA lifetimes)
        clk0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    }

    // locate flat item `lo`: block row by binary search over row_start, then walk the column blocks
    if (!located) {
        int a = 0, b = p.n_iblocks;   // row_start[a] <= lo < row_start[b]
        while (b - a > 1) {
            const int m = (a + b) >> 1;
            if (p.row_start[m] <= lo) a = m; else b = m;
        }
        w0.I = a; w0.c = 0;
        long long rem = lo - p.row_start[a];
        const int Ig = p.gblock0 + a;
        for (;;) {
            const int tb = sym_tiles_in_block(p.n_total, IBLK, TILE, (Ig + w0.c) % p.n_gblocks);
            if (rem < tb) break;
            rem -= tb; ++w0.c;
        }
        w0.t = (int)rem;
    }

    // producer: flat order, its own walker
    SymWalker pw = w0;
    int p_slot = 0;
    auto issue_next = [&]() {
        const int K = (p.gblock0 + pw.I + pw.c) % p.n_gblocks;
        const long long j0 = (long long)K * IBLK + (long long)pw.t * TILE;
        long long cnt = p.n_total - j0;
        if (cnt > TILE) cnt = TILE;
        const uint32_t bytes = (uint32_t)(cnt * sizeof(float4));
        mbar_expect_tx(&full_bar[p_slot], bytes);
        tma_bulk_g2s(tiles + (size_t)p_slot * TILE, p.pos_front + j0, bytes, &full_bar[p_slot]);
        sym_advance<IBLK, TILE>(pw, p);
        if (++p_slot == STAGES) p_slot = 0;
    };
    griddep_launch();   // the integrate kernel may be launched now; it waits for this grid to complete before it reads
    griddep_wait();     // the previous integrate kernel is complete: positions are final, the accumulator is zero
    if (tid == 0) {
        const int pre = ntiles < (STAGES - 1) ? ntiles : (STAGES - 1);
        for (int k = 0; k < pre; ++k) issue_next();
    }

    float xi[R], yi[R], zi[R];   // NEGATED positions of this thread's i-bodies (dx = xj + (-xi))
    float2 mi[P];
    SymWalker w = w0;
    bool new_row = true;
    int c_slot = 0, e_slot = 0;
    uint32_t c_parity = 0, e_parity = 0;
    int jbuf = 0;
    const int src_lane = (lane + 1) & 31;

    // flush this thread's i-side sums of the finished row segment into the global accumulator
    auto flush_row = [&]() {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long il = (long long)w.I * IBLK + r * THREADS + tid;
            if (il < p.n_local) {
                double* dst = p.acc64 + (p.row0 + il) * 4;
                atomicAdd(dst + 0, ssum[(0 * R + r) * THREADS + tid]);
                atomicAdd(dst + 1, ssum[(1 * R + r) * THREADS + tid]);
                atomicAdd(dst + 2, ssum[(2 * R + r) * THREADS + tid]);
            }
        }
    };

    // j side of a finished symmetric tile: the warps' partials are combined in a fixed order and added to the
    // global accumulator (one RED.ADD.F64 per body and component).  The combine is DEFERRED into the next
    // tile: every warp arrives on `jbar` when its partials are written and goes on; the wait comes a whole
    // ring round (32 steps) later, when all warps have long arrived, so no warp ever idles at a CTA barrier.
    // jpart is double buffered: a buffer is rewritten two tiles later, after the wait that follows its combine.
    bool jpend = false;
    int jpend_lo = 0, jpend_n = 0, jpend_buf = 0;   // j-bodies [jpend_lo, jpend_n) of the tile (SPLIT: the chunks this CTA evaluated)
    long long jpend_j0 = 0;
    uint32_t j_parity = 0;
    auto combine_pending = [&]() {
#ifdef SYM_DEBUG
        const long long tw0 = clock64();
#endif
        mbar_wait_warp(jbar, j_parity);
#ifdef SYM_DEBUG
        if (p.clk && lane == 0) atomicAdd(p.clk + 2047, (unsigned long long)(clock64() - tw0));
#endif
        j_parity ^= 1;
        const float* jb = jpart + (size_t)jpend_buf * NWARPS * 3 * TILE;
        for (int j = (SPLIT ? jpend_lo : 0) + tid; j < jpend_n; j += THREADS) {
            double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
            for (int wv = 0; wv < NWARPS; ++wv) {
                sx += (double)jb[(wv * 3 + 0) * TILE + j];
                sy += (double)jb[(wv * 3 + 1) * TILE + j];
                sz += (double)jb[(wv * 3 + 2) * TILE + j];
            }
            double* dst = p.acc64 + (jpend_j0 + j) * 4;
            atomicAdd(dst + 0, sx);
            atomicAdd(dst + 1, sy);
            atomicAdd(dst + 2, sz);
        }
        jpend = false;
    };

    for (int k = 0; k < ntiles; ++k) {
        SYM_DIVCHK(0);
        if (new_row) {
            new_row = false;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long il = (long long)w.I * IBLK + r * THREADS + tid;
                // padding rows sit far away with zero mass: d2 overflows to inf, rsqrt gives 0, nothing is added
                float4 b = make_float4(1.0e30f, 1.0e30f, 1.0e30f, 0.f);
                if (il < p.n_local) b = p.pos_front[p.row0 + il];
                xi[r] = -b.x; yi[r] = -b.y; zi[r] = -b.z;
                if (r & 1) mi[r >> 1].y = b.w; else mi[r >> 1].x = b.w;
                ssum[(0 * R + r) * THREADS + tid] = 0.0;
                ssum[(1 * R + r) * THREADS + tid] = 0.0;
                ssum[(2 * R + r) * THREADS + tid] = 0.0;
            }
        }
        if (warp == 0) {   // producer: the whole warp waits for the free slot, one lane issues the bulk copy
            const int kk = k + STAGES - 1;
            if (kk < ntiles) {
                if (kk >= STAGES) {
                    mbar_wait_warp(&empty_bar[e_slot], e_parity);
                    if (++e_slot == STAGES) { e_slot = 0; e_parity ^= 1; }
                }
                if (lane == 0) issue_next();
                __syncwarp();
            }
        }
        SYM_DIVCHK(1);
        const int s = c_slot;
        mbar_wait_warp(&full_bar[s], c_parity);
        SYM_DIVCHK(2);
        if (++c_slot == STAGES) { c_slot = 0; c_parity ^= 1; }
        const float4* __restrict__ tile = tiles + (size_t)s * TILE;

        const int Ig = p.gblock0 + w.I;
        const int K = (Ig + w.c) % p.n_gblocks;
        const long long j0 = (long long)K * IBLK + (long long)w.t * TILE;
        long long cntl = p.n_total - j0;
        const int jn = cntl > TILE ? TILE : (int)cntl;
        const float e2 = p.eps2_f;
        // SPLIT: the chunks [cb, ce) of this tile are this CTA's (all of them except in its first and last tile)
        const int cb = (SPLIT && k == 0) ? rg.c_first : 0;
        const int ce = (SPLIT && k == ntiles - 1) ? rg.c_last : CHUNKS;

        float2 ax[P], ay[P], az[P];
#pragma unroll
        for (int q = 0; q < P; ++q) ax[q] = ay[q] = az[q] = make_float2(0.f, 0.f);

        if (w.c == 0) {
            // diagonal block: ordered evaluation, j broadcast from shared memory, self pair masked by index
            const int dj0 = (int)(j0 - ((long long)Ig * IBLK + tid));
            auto ordered = [&](const float4 b, const int dj) {
#pragma unroll
                for (int q = 0; q < P; ++q) {
                    const float2 dx = __fadd2_rn(make_float2(b.x, b.x), make_float2(xi[2 * q], xi[2 * q + 1]));
                    const float2 dy = __fadd2_rn(make_float2(b.y, b.y), make_float2(yi[2 * q], yi[2 * q + 1]));
                    const float2 dz = __fadd2_rn(make_float2(b.z, b.z), make_float2(zi[2 * q], zi[2 * q + 1]));
                    float2 d2 = __ffma2_rn(dx, dx, make_float2(e2, e2));
                    d2 = __ffma2_rn(dy, dy, d2);
                    d2 = __ffma2_rn(dz, dz, d2);
                    const float2 ri = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
                    const float2 ri2 = __fmul2_rn(ri, ri);
                    const float2 mr = __fmul2_rn(make_float2(b.w, b.w), ri);
                    float2 sc = __fmul2_rn(mr, ri2);
                    if (dj == (2 * q) * THREADS) sc.x = 0.f;
                    if (dj == (2 * q + 1) * THREADS) sc.y = 0.f;
                    ax[q] = __ffma2_rn(dx, sc, ax[q]);
                    ay[q] = __ffma2_rn(dy, sc, ay[q]);
                    az[q] = __ffma2_rn(dz, sc, az[q]);
                }
            };
            const int jbeg = cb * 32, jend = (SPLIT && ce * 32 < jn) ? ce * 32 : jn;
            if (jbeg == 0 && jend == TILE) {
#pragma unroll 2
                for (int j = 0; j < TILE; ++j) ordered(tile[j], dj0 + j);
            } else if (SPLIT) {   // partial diagonal tiles are the rule with chunk-granular ranges, not the ragged exception
#pragma unroll 2
                for (int j = jbeg; j < jend; ++j) ordered(tile[j], dj0 + j);
            } else {
#pragma unroll 1
                for (int j = jbeg; j < jend; ++j) ordered(tile[j], dj0 + j);
            }
            SYM_DIVCHK(3);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (jpend) combine_pending();
            SYM_DIVCHK(4);
        } else {
            // symmetric tile: ring over 32-body chunks
            float* jp = jpart + ((size_t)jbuf * NWARPS + warp) * 3 * TILE;
#pragma unroll 1
            for (int c = cb; c < ce; ++c) {
                const int jl = c * 32 + lane;
                SYM_DIVCHK(10);
#ifdef SYM_DEBUG
                {   // log the first diverged chunk entries: CTA, warp, tile index in the CTA, chunk, active mask
                    const unsigned am = __activemask();
                    if (p.clk && am != 0xffffffffu && lane == __ffs(am) - 1) {
                        const unsigned long long idx = atomicAdd(p.clk + 2035, 1ull);
                        if (idx < 60) p.clk[1900 + idx] = ((unsigned long long)blockIdx.x << 52) | ((unsigned long long)warp << 48) |
                                                           ((unsigned long long)k << 40) | ((unsigned long long)c << 32) | am;
                    }
                }
#endif
                // padding j-bodies sit at the opposite far corner from padding i-bodies, so d2 is never 0
                float4 bj = make_float4(-1.0e30f, -1.0e30f, -1.0e30f, 0.f);
                if (jl < jn) bj = tile[jl];
                if (c == ce - 1) {   // last read of this ring slot
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                }
                float2 jx = make_float2(0.f, 0.f), jy = jx, jz = jx;
#pragma unroll UNROLL
                for (int st = 0; st < 32; ++st) {
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        const float2 dx = __fadd2_rn(make_float2(bj.x, bj.x), make_float2(xi[2 * q], xi[2 * q + 1]));
                        const float2 dy = __fadd2_rn(make_float2(bj.y, bj.y), make_float2(yi[2 * q], yi[2 * q + 1]));
                        const float2 dz = __fadd2_rn(make_float2(bj.z, bj.z), make_float2(zi[2 * q], zi[2 * q + 1]));
                        float2 d2 = __ffma2_rn(dx, dx, make_float2(e2, e2));
                        d2 = __ffma2_rn(dy, dy, d2);
                        d2 = __ffma2_rn(dz, dz, d2);
                        const float2 ri = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
                        const float2 ri2 = __fmul2_rn(ri, ri);
                        const float2 ri3 = __fmul2_rn(ri2, ri);
                        const float2 si = __fmul2_rn(make_float2(bj.w, bj.w), ri3);
                        const float2 sj = __fmul2_rn(mi[q], ri3);
                        ax[q] = __ffma2_rn(dx, si, ax[q]);
                        ay[q] = __ffma2_rn(dy, si, ay[q]);
                        az[q] = __ffma2_rn(dz, si, az[q]);
                        jx = __ffma2_rn(dx, sj, jx);
                        jy = __ffma2_rn(dy, sj, jy);
                        jz = __ffma2_rn(dz, sj, jz);
                    }
                    bj.x = __shfl_sync(0xffffffffu, bj.x, src_lane);
                    bj.y = __shfl_sync(0xffffffffu, bj.y, src_lane);
                    bj.z = __shfl_sync(0xffffffffu, bj.z, src_lane);
                    bj.w = __shfl_sync(0xffffffffu, bj.w, src_lane);
                    jx.x = __shfl_sync(0xffffffffu, jx.x, src_lane); jx.y = __shfl_sync(0xffffffffu, jx.y, src_lane);
                    jy.x = __shfl_sync(0xffffffffu, jy.x, src_lane); jy.y = __shfl_sync(0xffffffffu, jy.y, src_lane);
                    jz.x = __shfl_sync(0xffffffffu, jz.x, src_lane); jz.y = __shfl_sync(0xffffffffu, jz.y, src_lane);
                }
                SYM_DIVCHK(7);
                // after 32 rotations every lane holds its own j-body again, with the sum over this warp's i-bodies
                if (c == cb && jpend) combine_pending();   // previous tile's j side (its buffer is the other one)
                jp[0 * TILE + jl] = -(jx.x + jx.y);
                jp[1 * TILE + jl] = -(jy.x + jy.y);
                jp[2 * TILE + jl] = -(jz.x + jz.y);
            }
        }
        SYM_DIVCHK(5);
        // i side: fp32 tile sums into this thread's fp64 sums
#pragma unroll
        for (int q = 0; q < P; ++q) {
            ssum[(0 * R + 2 * q) * THREADS + tid] += (double)ax[q].x;
            ssum[(0 * R + 2 * q + 1) * THREADS + tid] += (double)ax[q].y;
            ssum[(1 * R + 2 * q) * THREADS + tid] += (double)ay[q].x;
            ssum[(1 * R + 2 * q + 1) * THREADS + tid] += (double)ay[q].y;
            ssum[(2 * R + 2 * q) * THREADS + tid] += (double)az[q].x;
            ssum[(2 * R + 2 * q + 1) * THREADS + tid] += (double)az[q].y;
        }
        if (w.c != 0) {
            // j side: this warp's partials are in shared memory; the combine happens in the next tile
            __syncwarp();
            if (lane == 0) mbar_arrive(jbar);
            jpend = true; jpend_lo = cb * 32; jpend_n = (SPLIT && ce * 32 < jn) ? ce * 32 : jn; jpend_j0 = j0; jpend_buf = jbuf;
            jbuf ^= 1;
        }

        SYM_DIVCHK(6);
        const int rowI = w.I;
        sym_advance<IBLK, TILE>(w, p);
        if (w.I != rowI || k == ntiles - 1) {
            const SymWalker keep = w;
            w.I = rowI;
            flush_row();
            w = keep;
            new_row = true;
        }
    }
    if (jpend) combine_pending();
    if (p.clk && tid == 0) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        if (blockIdx.x == 0) { p.clk[0] = clock64() - clk0; p.clk[1] = ns1 - ns0; }
        if (blockIdx.x < 1016) { p.clk[2 + 2 * blockIdx.x] = ns0; p.clk[3 + 2 * blockIdx.x] = ns1; }
    }
    sym_finish(p, total);   // several shards: "my sweep is done" once every CTA has got here
}

template <int THREADS, int R, int TILE, int STAGES>
constexpr size_t sym_smem_bytes() {
    return (size_t)STAGES * TILE * sizeof(float4) + (2 * STAGES + 1) * sizeof(uint64_t) + (size_t)3 * R * THREADS * sizeof(double) +
           (size_t)2 * (THREADS / 32) * 3 * TILE * sizeof(float);
}

// ------------------------------------------------------------------------------------------------
// fp64 twin.  No packing (there are no packed FP64 instructions); R i-bodies per thread, sums stay in fp64
// registers for the whole block row.  Per unordered pair: 3 DADD + 3 DFMA + 6 (MUFU.RSQ64H seed refined to
// |d|^-3, see mass_over_r3) + 2 (both masses) + 6 DFMA = 20 FP64-pipe operations, i.e. 10 per ordered
// interaction instead of the ordered sweep's 16.  14 32-bit SHFL move the j-body and its partials.
// Dynamic shared memory: tile ring | mbarriers (full, empty, jbar) | j-partials [2][NWARPS][3][TILE] doubles
// (double buffered for the deferred combine).
// Padding bodies sit at +-1e150: d2 stays finite, the refined |d|^-3 underflows to 0 (inf would give
// inf * 0 = NaN in the refinement).
// ------------------------------------------------------------------------------------------------
template <int THREADS, int R, int TILE, int STAGES, int MINB, int UNROLL, int SPLIT = 0>
__global__ void __launch_bounds__(THREADS, MINB) sym_sweep_kernel_f64(const SymParams p) {
    constexpr int IBLK = THREADS * R;
    constexpr int NWARPS = THREADS / 32;
    constexpr int CHUNKS = TILE / 32;
    static_assert(TILE % 32 == 0 && IBLK % TILE == 0, "geometry");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double4* tiles = reinterpret_cast<double4*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * TILE * sizeof(double4));
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* jbar = empty_bar + STAGES;                             // j-partials of a tile complete
    double* jpart = reinterpret_cast<double*>(jbar + 1);             // [2][NWARPS][3][TILE]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NWARPS); }
        mbar_init(jbar, NWARPS);
        mbar_fence_init();
    }
    __syncthreads();
    if (p.sync.wait_flags) peer_wait(p.sync);   // as in the fp32 kernel
    long long share_lo, share_hi;
    sym_share(p, share_lo, share_hi);
    const long long total = share_hi - share_lo;
    const long long S = gridDim.x;
    SymRange rg;
    SymWalker w0;
    bool located = false;   // as in the fp32 kernel
    if constexpr (SPLIT != 0) {
        if (p.row_cost) {
            if (!sym_locate_weighted<IBLK, TILE>(p, blockIdx.x, S, rg, w0)) { sym_finish(p, total); return; }
            located = true;
        }
    }
    if (!located && !sym_cta_range<CHUNKS, SPLIT>(share_lo, total, blockIdx.x, S, rg)) { sym_finish(p, total); return; }
    const long long lo = rg.lo;
    const int ntiles = (int)(rg.hi - rg.lo);

    unsigned long long clk0 = 0, ns0 = 0;
    if (p.clk && blockIdx.x == 0 && tid == 0) {
        clk0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    }

    if (!located) {
        int a = 0, b = p.n_iblocks;
        while (b - a > 1) {
            const int m = (a + b) >> 1;
            if (p.row_start[m] <= lo) a = m; else b = m;
        }
        w0.I = a; w0.c = 0;
        long long rem = lo - p.row_start[a];
        const int Ig = p.gblock0 + a;
        for (;;) {
            const int tb = sym_tiles_in_block(p.n_total, IBLK, TILE, (Ig + w0.c) % p.n_gblocks);
            if (rem < tb) break;
            rem -= tb; ++w0.c;
        }
        w0.t = (int)rem;
    }

    SymWalker pw = w0;
    int p_slot = 0;
    auto issue_next = [&]() {
        const int K = (p.gblock0 + pw.I + pw.c) % p.n_gblocks;
        const long long j0 = (long long)K * IBLK + (long long)pw.t * TILE;
        long long cnt = p.n_total - j0;
        if (cnt > TILE) cnt = TILE;
        const uint32_t bytes = (uint32_t)(cnt * sizeof(double4));
        mbar_expect_tx(&full_bar[p_slot], bytes);
        tma_bulk_g2s(tiles + (size_t)p_slot * TILE, p.pos_front_d + j0, bytes, &full_bar[p_slot]);
        sym_advance<IBLK, TILE>(pw, p);
        if (++p_slot == STAGES) p_slot = 0;
    };
    griddep_launch();   // the integrate kernel may be launched now; it waits for this grid to complete before it reads
    griddep_wait();     // the previous integrate kernel is complete: positions are final, the accumulator is zero
    if (tid == 0) {
        const int pre = ntiles < (STAGES - 1) ? ntiles : (STAGES - 1);
        for (int k = 0; k < pre; ++k) issue_next();
    }

    double xi[R], yi[R], zi[R], mi[R];   // NEGATED positions, masses
    double sx[R], sy[R], sz[R];
    SymWalker w = w0;
    bool new_row = true;
    int c_slot = 0, e_slot = 0;
    uint32_t c_parity = 0, e_parity = 0;
    const int src_lane = (lane + 1) & 31;
    const double e2 = p.eps2_d;

    // deferred j-side combine, exactly as in the fp32 kernel: arrive on `jbar` when this warp's partials of a
    // tile are written, combine that tile inside the next one after the first ring round; no CTA barrier
    bool jpend = false;
    int jpend_lo = 0, jpend_n = 0, jpend_buf = 0, jbuf = 0;
    long long jpend_j0 = 0;
    uint32_t j_parity = 0;
    auto combine_pending = [&]() {
        mbar_wait_warp(jbar, j_parity);
        j_parity ^= 1;
        const double* jb = jpart + (size_t)jpend_buf * NWARPS * 3 * TILE;
        for (int j = (SPLIT ? jpend_lo : 0) + tid; j < jpend_n; j += THREADS) {
            double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll
            for (int wv = 0; wv < NWARPS; ++wv) {
                ax += jb[(wv * 3 + 0) * TILE + j];
                ay += jb[(wv * 3 + 1) * TILE + j];
                az += jb[(wv * 3 + 2) * TILE + j];
            }
            double* dst = p.acc64 + (jpend_j0 + j) * 4;
            atomicAdd(dst + 0, ax);
            atomicAdd(dst + 1, ay);
            atomicAdd(dst + 2, az);
        }
        jpend = false;
    };

    for (int k = 0; k < ntiles; ++k) {
        if (new_row) {
            new_row = false;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long il = (long long)w.I * IBLK + r * THREADS + tid;
                double4 b = make_double4(1.0e150, 1.0e150, 1.0e150, 0.0);
                if (il < p.n_local) b = p.pos_front_d[p.row0 + il];
                xi[r] = -b.x; yi[r] = -b.y; zi[r] = -b.z; mi[r] = b.w;
                sx[r] = 0.0; sy[r] = 0.0; sz[r] = 0.0;
            }
        }
        if (warp == 0) {   // producer: the whole warp waits for the free slot, one lane issues the bulk copy
            const int kk = k + STAGES - 1;
            if (kk < ntiles) {
                if (kk >= STAGES) {
                    mbar_wait_warp(&empty_bar[e_slot], e_parity);
                    if (++e_slot == STAGES) { e_slot = 0; e_parity ^= 1; }
                }
                if (lane == 0) issue_next();
                __syncwarp();
            }
        }
        const int s = c_slot;
        mbar_wait_warp(&full_bar[s], c_parity);
        if (++c_slot == STAGES) { c_slot = 0; c_parity ^= 1; }
        const double4* __restrict__ tile = tiles + (size_t)s * TILE;

        const int Ig = p.gblock0 + w.I;
        const int K = (Ig + w.c) % p.n_gblocks;
        const long long j0 = (long long)K * IBLK + (long long)w.t * TILE;
        long long cntl = p.n_total - j0;
        const int jn = cntl > TILE ? TILE : (int)cntl;
        const int cb = (SPLIT && k == 0) ? rg.c_first : 0;             // SPLIT: this CTA's chunks of the tile, as in the fp32 kernel
        const int ce = (SPLIT && k == ntiles - 1) ? rg.c_last : CHUNKS;

        if (w.c == 0) {
            // diagonal block: ordered, j broadcast from shared memory, self pair masked by index
            const int dj0 = (int)(j0 - ((long long)Ig * IBLK + tid));
            const int jbeg = cb * 32, jend = (SPLIT && ce * 32 < jn) ? ce * 32 : jn;
#pragma unroll 2
            for (int j = jbeg; j < jend; ++j) {
                const double4 b = tile[j];
                const int dj = dj0 + j;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double dx = b.x + xi[r], dy = b.y + yi[r], dz = b.z + zi[r];
                    const double d2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, e2)));
                    double sc = mass_over_r3(b.w, d2);
                    if (dj == r * THREADS) sc = 0.0;
                    sx[r] = fma(dx, sc, sx[r]);
                    sy[r] = fma(dy, sc, sy[r]);
                    sz[r] = fma(dz, sc, sz[r]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (jpend) combine_pending();
        } else {
            double* jp = jpart + ((size_t)jbuf * NWARPS + warp) * 3 * TILE;
#pragma unroll 1
            for (int c = cb; c < ce; ++c) {
                const int jl = c * 32 + lane;
                double4 bj = make_double4(-1.0e150, -1.0e150, -1.0e150, 0.0);
                if (jl < jn) bj = tile[jl];
                if (c == ce - 1) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                }
                double jx = 0.0, jy = 0.0, jz = 0.0;
#pragma unroll UNROLL
                for (int st = 0; st < 32; ++st) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double dx = bj.x + xi[r], dy = bj.y + yi[r], dz = bj.z + zi[r];
                        const double d2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, e2)));
                        const double w3 = mass_over_r3(1.0, d2);   // |d|^-3 (the multiply by 1.0 folds away)
                        const double si = bj.w * w3, sj = mi[r] * w3;
                        sx[r] = fma(dx, si, sx[r]);
                        sy[r] = fma(dy, si, sy[r]);
                        sz[r] = fma(dz, si, sz[r]);
                        jx = fma(dx, sj, jx);
                        jy = fma(dy, sj, jy);
                        jz = fma(dz, sj, jz);
                    }
                    bj.x = __shfl_sync(0xffffffffu, bj.x, src_lane);
                    bj.y = __shfl_sync(0xffffffffu, bj.y, src_lane);
                    bj.z = __shfl_sync(0xffffffffu, bj.z, src_lane);
                    bj.w = __shfl_sync(0xffffffffu, bj.w, src_lane);
                    jx = __shfl_sync(0xffffffffu, jx, src_lane);
                    jy = __shfl_sync(0xffffffffu, jy, src_lane);
                    jz = __shfl_sync(0xffffffffu, jz, src_lane);
                }
                if (c == cb && jpend) combine_pending();   // previous tile's j side (its buffer is the other one)
                jp[0 * TILE + jl] = -jx;
                jp[1 * TILE + jl] = -jy;
                jp[2 * TILE + jl] = -jz;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(jbar);
            jpend = true; jpend_lo = cb * 32; jpend_n = (SPLIT && ce * 32 < jn) ? ce * 32 : jn; jpend_j0 = j0; jpend_buf = jbuf;
            jbuf ^= 1;
        }

        const int rowI = w.I;
        sym_advance<IBLK, TILE>(w, p);
        if (w.I != rowI || k == ntiles - 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long il = (long long)rowI * IBLK + r * THREADS + tid;
                if (il < p.n_local) {
                    double* dst = p.acc64 + (p.row0 + il) * 4;
                    atomicAdd(dst + 0, sx[r]);
                    atomicAdd(dst + 1, sy[r]);
                    atomicAdd(dst + 2, sz[r]);
                }
            }
            new_row = true;
        }
    }
    if (jpend) combine_pending();
    if (p.clk && blockIdx.x == 0 && tid == 0) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        p.clk[0] = clock64() - clk0;
        p.clk[1] = ns1 - ns0;
    }
    sym_finish(p, total);
}

template <int THREADS, int TILE, int STAGES>
constexpr size_t sym64_smem_bytes() {
    return (size_t)STAGES * TILE * sizeof(double4) + (2 * STAGES + 1) * sizeof(uint64_t) + (size_t)2 * (THREADS / 32) * 3 * TILE * sizeof(double);
}

// ------------------------------------------------------------------------------------------------
// O(N) second half of a symmetric step: accumulator -> a, v', r' for this shard's rows, accumulator
// cleared for the next step.  Same separately rounded stage-2 arithmetic as the fused epilogue
// (finalize_body, np2.py:110-115), including the peer stores of the fused position exchange.
// ------------------------------------------------------------------------------------------------
struct IntegrateParams {
    SweepParams sp;      // pos/vel/acc buffers, G, T, row0, n_local, peers
    double* acc64;       // [n_pad][4] this shard's accumulator
    // several shards: every shard accumulated partial sums for ALL bodies; the owner of a row adds the
    // partials of all shards in rank order (own memory + NVLink peer loads, all in flight at once) and zeroes
    // them for the next step — a reduce-scatter fused into the integrate kernel.  n_src == 0: single shard.
    int n_src;
    double* acc_src[kMaxPeers + 1];
    PeerSync sync;       // several shards: wait for every shard's sweep, then tell the peers "my integrate is done"
};

template <typename REAL>
__global__ void sym_integrate_kernel(const IntegrateParams q) {
    using V4 = typename Vec4<REAL>::type;
    const long long il = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    griddep_wait();     // this shard's sweep is complete (all its atomic adds are visible)
    griddep_launch();   // the next sweep may start its prologue; it waits for this grid before it reads positions or adds
    if (q.sync.wait_flags) peer_wait(q.sync);   // every shard's sweep of this step is complete
    if (q.sp.clk && q.sync.wait_flags && blockIdx.x == 0 && threadIdx.x == 0) {   // when the wait ended (gravb200_timings: wait vs work of this kernel)
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        q.sp.clk[2048] = t;
    }
    if (il < q.sp.n_local) {
        double sx, sy, sz;
        if (q.n_src == 0) {
            double4* src = reinterpret_cast<double4*>(q.acc64) + (q.sp.row0 + il);
            const double4 s = *src;
            *src = make_double4(0.0, 0.0, 0.0, 0.0);
            sx = s.x; sy = s.y; sz = s.z;
        } else {
            sx = sy = sz = 0.0;
            constexpr int B = 8;   // sources in flight per batch (one NVLink round trip per batch, not per source)
            for (int r0 = 0; r0 < q.n_src; r0 += B) {
                double4 part[B];
#pragma unroll
                for (int u = 0; u < B; ++u)
                    if (r0 + u < q.n_src) part[u] = ld_cg_d4(reinterpret_cast<const double4*>(q.acc_src[r0 + u]) + (q.sp.row0 + il));
#pragma unroll
                for (int u = 0; u < B; ++u)
                    if (r0 + u < q.n_src) {
                        sx += part[u].x; sy += part[u].y; sz += part[u].z;   // rank order
                        double2* z = reinterpret_cast<double2*>(reinterpret_cast<double4*>(q.acc_src[r0 + u]) + (q.sp.row0 + il));
                        z[0] = make_double2(0.0, 0.0); z[1] = make_double2(0.0, 0.0);   // ready for the next sweep (peer store)
                    }
            }
        }
        const V4 ri = reinterpret_cast<const V4*>(q.sp.pos_front)[q.sp.row0 + il];
        finalize_body(q.sp, il, sx, sy, sz, ri, REAL(0));
    }
    if (q.sp.clk && q.sync.wait_flags && threadIdx.x == 0) {   // latest CTA end
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(q.sp.clk + 2049, t);
    }
    peer_signal(q.sync);
}

}  // namespace gravb200
