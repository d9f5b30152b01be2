// gravitation_b200 — persistent multi-step kernel for universes whose positions fit in one SM's shared memory.
//
// Below ~10^4 bodies one step of the large-N kernels is bound by what surrounds the arithmetic: a launch per
// step, the TMA-ring prologue, the stream-K fix-up of split i-blocks (nbody_kernels.cuh) — 25 us per step at
// N = 4096 for 6 us of arithmetic (profiles/r01_small_n.md); the reference's pc2 pays 0.29 ms per step at
// N = 256 for its six copies.  Here ONE cooperative launch runs k whole steps
// (stage 1 + stage 2 of _base_.py:136-145) with a grid-wide barrier between them:
//
//   * every CTA owns `rpc` consecutive rows (i-bodies) and keeps their velocities in registers for all k steps;
//   * per step the CTA pulls ALL positions into shared memory with TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx), one copy and one mbarrier per j-slice, so a warp starts as soon as its own slice has landed;
//   * warp (g, s) evaluates row group g (32 rows) against j-slice s, ordered, pc2.py:59-91 form.  A lane holds R
//     consecutive rows (R / 2 packed f32x2 pairs) and serves one of R interleaved sub-slices: with R = 4 the four
//     quarter-warps read four ADJACENT bodies with one LDS.128 (R = 2 needs twice as many shared-memory reads per
//     interaction and is bound by them: 36 instead of 29 cycles per pair, profiles/r02_small_n.md), so 32 rows
//     fill a warp with the same 12 packed FP32 + 2 MUFU per pair of interactions as the large-N ordered sweep.
//     The self pair is masked by index only inside the 32-body stretch of the slice that holds the warp's own rows;
//   * fp32 partial sums per <= 512 terms, fp64 across them, sub-slices combined by a shuffle butterfly, slices combined
//     through shared memory in slice order (bit-reproducible), then a = G * sum, v' = v + a*T, r' = r + v'*T with
//     separately rounded operations (np2.py:110-115) into the back buffers;
//   * grid barrier: one atomic counter, release/acquire at gpu scope, then a proxy fence because the next
//     step's positions were written by ordinary stores and are read by the TMA engine (async proxy).
//
// k = 1 without the barrier is stage 1 + fused stage 2 of a single step() (the split-stage semantics of the
// boundary: back buffers are written, gravb200_stage2 commits), so steps(k) and k calls of step() run the same
// arithmetic in the same order and agree bit for bit.
#pragma once

#include "nbody_kernels.cuh"

namespace gravb200 {

struct SmallParams {
    void* pos[2];                  // [n] {x,y,z,m} of all bodies, double buffered
    void* vel[2];                  // [n] {vx,vy,vz,0}
    void* acc;                     // [n] {ax,ay,az,0}
    unsigned long long* gbar;      // grid barrier: monotonic arrival counter
    unsigned long long gbar_base;  // its value when this launch starts
    int* error;                    // set to 1 if the grid barrier times out
    long long n;                   // bodies
    int rpc;                       // rows per CTA (even)
    int ng;                        // row groups of 32 per CTA; divides the number of warps
    int slice;                     // j-slots per slice (multiple of R * UNROLL); slices * slice >= n
    int front;                     // buffer the first step reads
    int k;                         // steps in this launch
    int integrate;                 // 0: accelerations only (k == 1)
    double G, T;
    float eps2_f;
    double eps2_d;
    unsigned long long* clk;       // optional [2]: CTA 0 writes {SM cycles, ns} of its lifetime
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// padding: rows beyond n sit far away with zero mass, padding j-slots at the opposite far corner, so d2 is never 0
// (fp32: d2 overflows to +inf and rsqrt gives 0; fp64: d2 stays finite and |d|^-3 underflows to 0)
__device__ __forceinline__ float4 ld_cg_v4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ double4 ld_cg_v4(const double4* p) { return ld_cg_d4(p); }
template <typename REAL> __device__ __forceinline__ REAL small_far();
template <> __device__ __forceinline__ float small_far<float>() { return 1.0e30f; }
template <> __device__ __forceinline__ double small_far<double>() { return 1.0e150; }

#ifdef SMALL_DEBUG   // dev builds: time stamps (ns) of the middle CTA's last step, scripts/dbg/small_dbg.py
#define SMALL_STAMP(i)                                                                         \
    do {                                                                                       \
        if (p.clk && blockIdx.x == gridDim.x / 2 && step == p.k - 1) {                         \
            unsigned long long t_;                                                             \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
            p.clk[64 + (i)] = t_;                                                              \
        }                                                                                      \
    } while (0)
#else
#define SMALL_STAMP(i) do {} while (0)
#endif

template <typename REAL>
constexpr size_t small_smem_bytes(int n_slots, int nsl, int ng) {
    return 128 + (size_t)n_slots * 4 * sizeof(REAL) + (size_t)nsl * 3 * 32 * ng * sizeof(double);
}

template <typename REAL, int THREADS, int UNROLL, int R>
__global__ void __launch_bounds__(THREADS, 1) small_steps_kernel(const SmallParams p) {
    using V4 = typename Vec4<REAL>::type;
    constexpr int NWARPS = THREADS / 32;
    constexpr bool F32 = sizeof(REAL) == 4;
    constexpr int FLUSH = 512;   // fp32: terms per lane between two fp64 flushes
    constexpr int LPS = 32 / R;  // lanes per j-sub-slice; a lane holds R consecutive rows, the warp 32 rows
    constexpr int P = R / 2;
    static_assert(NWARPS <= 16, "one mbarrier per slice in the first 128 bytes of shared memory");
    static_assert(R == 2 || R == 4 || R == 8, "rows per lane");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);            // [nsl]
    V4* tile = reinterpret_cast<V4*>(smem_raw + 128);                  // [nsl * slice]
    const int nsl = NWARPS / p.ng;
    const int n_slots = nsl * p.slice;
    const int rows_cta = 32 * p.ng;
    double* red = reinterpret_cast<double*>(tile + n_slots);           // [nsl][3][rows_cta]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = (int)p.n;
    const int row_lo = blockIdx.x * p.rpc;                             // first row of this CTA
    const int g = warp % p.ng, sl = warp / p.ng;                       // row group, j-slice of this warp

    if (tid == 0) {
        for (int s = 0; s < nsl; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    {
        V4 pad;
        pad.x = pad.y = pad.z = -small_far<REAL>(); pad.w = 0;
        for (int j = n + tid; j < n_slots; j += THREADS) tile[j] = pad;   // never overwritten: the copies cover [0, n)
    }
    __syncthreads();

    unsigned long long clk0 = 0, ns0 = 0;
    if (p.clk && blockIdx.x == 0 && tid == 0) {
        clk0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    }

    // rows of this lane: i0 .. i0 + R - 1 (R / 2 packed pairs); j-sub-slice q of the warp: lanes q * LPS .. q * LPS + LPS - 1
    // (R = 4: the four quarter-warps read four adjacent bodies with one LDS.128).  finalize: thread t < rpc owns row row_lo + t
    const int rl0 = g * 32 + R * (lane % LPS);   // first row of this lane inside the CTA
    const int i0 = row_lo + rl0;
    const int q = lane / LPS;
    const bool owner = tid < p.rpc && row_lo + tid < n;
    V4 vown;
    vown.x = vown.y = vown.z = vown.w = 0;
    if (owner) vown = reinterpret_cast<const V4*>(p.vel[p.front])[row_lo + tid];

    // iterations of this warp: t in [0, slice / R) handles j-slot sl * slice + R t + q; only the stretch that holds
    // the warp's own 32 rows needs the self-pair mask
    const int TR = p.slice / R;
    const int sl0 = sl * p.slice;
    const int R0 = row_lo + g * 32;
    int ta = (R0 - sl0) / R, tb = (R0 + 32 - sl0 + R - 1) / R;
    if (R0 < sl0) ta = 0;
    ta = ta / UNROLL * UNROLL;
    tb = (tb + UNROLL - 1) / UNROLL * UNROLL;
    if (ta > TR) ta = TR;
    if (tb > TR) tb = TR;
    if (tb < ta || R0 + 32 <= sl0) tb = ta;
    const int t_real = min(TR, (max(0, min(p.slice, n - sl0)) + R * UNROLL - 1) / (R * UNROLL) * UNROLL);   // beyond: padding only

    int front = p.front;
    for (int step = 0; step < p.k; ++step) {
        const V4* __restrict__ posf = reinterpret_cast<const V4*>(p.pos[front]);
        if (tid == 0) SMALL_STAMP(0);
        if (warp == 0) {
            if (step > 0) {
                // grid barrier: every CTA has stored its r' of step - 1 (arrival below, after the stores)
                const unsigned long long target = p.gbar_base + (unsigned long long)gridDim.x * (unsigned long long)step;
                unsigned long long t0 = 0, t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                bool ok;
                do {
                    ok = ld_acquire_gpu_u64(p.gbar) >= target;
                    if (!ok) {
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (t1 - t0 > 20ull * 1000 * 1000 * 1000) { *p.error = 2; ok = true; }
                    }
                } while (!__all_sync(0xffffffffu, ok ? 1 : 0));
            }
            if (lane == 0) SMALL_STAMP(1);
            if (lane < nsl) {   // lane s issues the copy of slice s (one lane looping over the slices took 0.9 us at 8 slices)
                fence_proxy_async_all();   // r' came from ordinary stores; the bulk copies below read it through the async proxy
                const int j0 = lane * p.slice;
                const int cnt = min(p.slice, n - j0);
                if (cnt > 0) {
                    const uint32_t bytes = (uint32_t)((size_t)cnt * sizeof(V4));
                    mbar_expect_tx(&full[lane], bytes);
                    tma_bulk_g2s(tile + j0, posf + j0, bytes, &full[lane]);
                } else {
                    mbar_arrive(&full[lane]);   // nothing to copy: complete the phase
                }
            }
            if (lane == 0) SMALL_STAMP(2);
            __syncwarp();
        }

        // this lane's rows (negated: d = r_j + (-r_i)); they come from global memory, the slice that holds them may
        // not have landed yet (they are this CTA's own rows: written by this CTA before its last CTA barrier)
        REAL nx[R], ny[R], nz[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            nx[r] = ny[r] = nz[r] = -small_far<REAL>();
            if (rl0 + r < p.rpc && i0 + r < n) {
                if constexpr (F32) {
                    // three SCALAR loads on purpose: the pair (x of row 2c, x of row 2c + 1) feeds packed instructions
                    // and must sit in an aligned register pair; components that arrive as one 128-bit load result
                    // stay in that quad and ptxas re-assembles the pair with two MOVs before every use (4.75 extra
                    // instructions per pair of interactions)
                    const float* src = reinterpret_cast<const float*>(posf + i0 + r);
                    nx[r] = -ld_cg_f32(src); ny[r] = -ld_cg_f32(src + 1); nz[r] = -ld_cg_f32(src + 2);
                } else {
                    const V4 b = ld_cg_v4(posf + i0 + r);
                    nx[r] = -b.x; ny[r] = -b.y; nz[r] = -b.z;
                }
            }
        }
        float2 px[P], py[P], pz[P];
        if constexpr (F32) {
#pragma unroll
            for (int c = 0; c < P; ++c) {
                px[c] = make_float2(nx[2 * c], nx[2 * c + 1]);
                py[c] = make_float2(ny[2 * c], ny[2 * c + 1]);
                pz[c] = make_float2(nz[2 * c], nz[2 * c + 1]);
            }
        }
        mbar_wait_warp(&full[sl], (uint32_t)(step & 1));
        if (lane == 0) SMALL_STAMP(3 + warp);
        const V4* __restrict__ tj = tile + sl0 + q;
        const int dj0 = sl0 + q - i0;   // j - i0 at t = 0

        double sx[R], sy[R], sz[R];
#pragma unroll
        for (int r = 0; r < R; ++r) sx[r] = sy[r] = sz[r] = 0.0;
        if constexpr (F32) {
            const float e2 = p.eps2_f;
            float2 ax[P], ay[P], az[P];
#pragma unroll
            for (int c = 0; c < P; ++c) ax[c] = ay[c] = az[c] = make_float2(0.f, 0.f);
            auto interact = [&](const float4 b, auto masked, const int dj) {
#pragma unroll
                for (int c = 0; c < P; ++c) {
                    const float2 dx = __fadd2_rn(make_float2(b.x, b.x), px[c]);
                    const float2 dy = __fadd2_rn(make_float2(b.y, b.y), py[c]);
                    const float2 dz = __fadd2_rn(make_float2(b.z, b.z), pz[c]);
                    float2 d2 = __ffma2_rn(dx, dx, make_float2(e2, e2));
                    d2 = __ffma2_rn(dy, dy, d2);
                    d2 = __ffma2_rn(dz, dz, d2);
                    const float2 ri = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
                    const float2 ri2 = __fmul2_rn(ri, ri);
                    const float2 mr = __fmul2_rn(make_float2(b.w, b.w), ri);
                    float2 sc = __fmul2_rn(mr, ri2);
                    if constexpr (decltype(masked)::value) {
                        if (dj == 2 * c) sc.x = 0.f;
                        if (dj == 2 * c + 1) sc.y = 0.f;
                    }
                    ax[c] = __ffma2_rn(dx, sc, ax[c]);
                    ay[c] = __ffma2_rn(dy, sc, ay[c]);
                    az[c] = __ffma2_rn(dz, sc, az[c]);
                }
            };
            auto flush = [&]() {
#pragma unroll
                for (int c = 0; c < P; ++c) {
                    sx[2 * c] += (double)ax[c].x; sx[2 * c + 1] += (double)ax[c].y;
                    sy[2 * c] += (double)ay[c].x; sy[2 * c + 1] += (double)ay[c].y;
                    sz[2 * c] += (double)az[c].x; sz[2 * c + 1] += (double)az[c].y;
                    ax[c] = ay[c] = az[c] = make_float2(0.f, 0.f);
                }
            };
            auto run = [&](int t0, int t1, auto masked) {
                for (int tb0 = t0; tb0 < t1; tb0 += FLUSH) {
                    const int te = min(t1, tb0 + FLUSH);
#pragma unroll UNROLL
                    for (int t = tb0; t < te; ++t) interact(tj[R * t], masked, dj0 + R * t);
                    flush();
                }
            };
            run(0, min(ta, t_real), std::false_type{});
            run(ta, min(tb, t_real), std::true_type{});
            run(tb, t_real, std::false_type{});
        } else {
            const double e2 = p.eps2_d;
            auto interact = [&](const double4 b, auto masked, const int dj) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double dx = b.x + nx[r], dy = b.y + ny[r], dz = b.z + nz[r];
                    const double d2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, e2)));
                    double sc = mass_over_r3(b.w, d2);
                    if constexpr (decltype(masked)::value) {
                        if (dj == r) sc = 0.0;
                    }
                    sx[r] = fma(dx, sc, sx[r]); sy[r] = fma(dy, sc, sy[r]); sz[r] = fma(dz, sc, sz[r]);
                }
            };
            auto run = [&](int t0, int t1, auto masked) {
#pragma unroll UNROLL
                for (int t = t0; t < t1; ++t) interact(tj[R * t], masked, dj0 + R * t);
            };
            run(0, min(ta, t_real), std::false_type{});
            run(ta, min(tb, t_real), std::true_type{});
            run(tb, t_real, std::false_type{});
        }

        if (lane == 0) SMALL_STAMP(19 + warp);
        // the R j-sub-slices of the warp (fixed butterfly order), then the slices in order through shared memory
#pragma unroll
        for (int off = LPS; off < 32; off <<= 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                sx[r] += __shfl_xor_sync(0xffffffffu, sx[r], off);
                sy[r] += __shfl_xor_sync(0xffffffffu, sy[r], off);
                sz[r] += __shfl_xor_sync(0xffffffffu, sz[r], off);
            }
        }
        if (q == 0) {
            double* rs = red + (size_t)sl * 3 * rows_cta + rl0;
#pragma unroll
            for (int r = 0; r < R; ++r) { rs[r] = sx[r]; rs[rows_cta + r] = sy[r]; rs[2 * rows_cta + r] = sz[r]; }
        }
        __syncthreads();
        if (tid == 0) SMALL_STAMP(35);
        if (owner) {
            double tx = 0.0, ty = 0.0, tz = 0.0;
            for (int s = 0; s < nsl; ++s) {
                const double* rs = red + (size_t)s * 3 * rows_cta + tid;
                tx += rs[0]; ty += rs[rows_cta]; tz += rs[2 * rows_cta];
            }
            const int row = row_lo + tid;
            V4 ri = tile[row];   // this step's position and the mass
            if constexpr (F32) {
                const float ax = (float)(tx * p.G), ay = (float)(ty * p.G), az = (float)(tz * p.G);
                reinterpret_cast<float4*>(p.acc)[row] = make_float4(ax, ay, az, 0.f);
                if (p.integrate) {
                    const float T = (float)p.T;
                    vown.x = __fadd_rn(vown.x, __fmul_rn(ax, T));
                    vown.y = __fadd_rn(vown.y, __fmul_rn(ay, T));
                    vown.z = __fadd_rn(vown.z, __fmul_rn(az, T));
                    ri.x = __fadd_rn(ri.x, __fmul_rn(vown.x, T));
                    ri.y = __fadd_rn(ri.y, __fmul_rn(vown.y, T));
                    ri.z = __fadd_rn(ri.z, __fmul_rn(vown.z, T));
                }
            } else {
                const double ax = __dmul_rn(tx, p.G), ay = __dmul_rn(ty, p.G), az = __dmul_rn(tz, p.G);
                reinterpret_cast<double4*>(p.acc)[row] = make_double4(ax, ay, az, 0.0);
                if (p.integrate) {
                    const double T = p.T;
                    vown.x = __dadd_rn(vown.x, __dmul_rn(ax, T));
                    vown.y = __dadd_rn(vown.y, __dmul_rn(ay, T));
                    vown.z = __dadd_rn(vown.z, __dmul_rn(az, T));
                    ri.x = __dadd_rn(ri.x, __dmul_rn(vown.x, T));
                    ri.y = __dadd_rn(ri.y, __dmul_rn(vown.y, T));
                    ri.z = __dadd_rn(ri.z, __dmul_rn(vown.z, T));
                }
            }
            if (p.integrate) {
                reinterpret_cast<V4*>(p.vel[front ^ 1])[row] = vown;
                reinterpret_cast<V4*>(p.pos[front ^ 1])[row] = ri;
            }
        }
        // everybody is done with `tile` and `red` of this step; the CTA's r' is stored: arrive at the grid barrier
        __syncthreads();
        if (tid == 0) SMALL_STAMP(36);
        if (step + 1 < p.k && tid == 0) {
            __threadfence();
            atomicAdd(p.gbar, 1ull);
        }
        front ^= 1;
    }
    if (p.clk && blockIdx.x == 0 && tid == 0) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        p.clk[0] = clock64() - clk0;
        p.clk[1] = ns1 - ns0;
    }
}

}  // namespace gravb200
