// gravb200.cu — C-ABI shim (include/gravb200.h) around the sm_100a sweep kernel.
//
// The context owns every device allocation, the stream, the events and (multi-GPU) the NCCL
// communicator.  There is deliberately no CPU implementation in this file: every compute entry
// point needs a CUDA device and fails with GRAVB200_ECUDA / GRAVB200_ENODEV otherwise.
//
// Reference interfaces replaced (pleiszenburg/gravitation, src/gravitation/kernel/):
//   start_kernel  pc2.py:96-145      -> gravb200_ctx_create + gravb200_upload
//   step_stage1   pc2.py:147-162     -> gravb200_stage1   (no per-step H2D/D2H)
//   step_stage2   np2.py:110-115     -> fused into the sweep epilogue; gravb200_stage2 commits
//   stop_kernel   _base_.py:174-177  -> gravb200_ctx_destroy
#include "../../include/gravb200.h"
#include "nbody_kernels.cuh"
#include "nbody_sym.cuh"
#include "nbody_small.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only; the library is dlopen'ed so single-GPU use needs no NCCL

#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

using namespace gravb200;

namespace {

thread_local char g_err[1024] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(GRAVB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                 \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL, resolved at run time
// ---------------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool tried = false;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.handle) return 0;
    if (g_nccl.tried) return fail(GRAVB200_ENCCL, "NCCL library could not be loaded");
    g_nccl.tried = true;
    const char* names[] = {getenv("GRAVB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return fail(GRAVB200_ENCCL, "dlopen(libnccl.so.2) failed: %s", dlerror());
#define SYM(field, name)                                                          \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name);                        \
    if (!g_nccl.field) return fail(GRAVB200_ENCCL, "NCCL symbol %s missing", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(AllGather, "ncclAllGather");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GetErrorString, "ncclGetErrorString");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
#undef SYM
    return 0;
}
#define NC(call)                                                                           \
    do {                                                                                   \
        ncclResult_t r_ = (call);                                                          \
        if (r_ != ncclSuccess)                                                             \
            return fail(GRAVB200_ENCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Kernel variants
// ---------------------------------------------------------------------------------------------
struct Variant {
    const char* name;
    int threads, r, tile, stages, minb, pack;
    size_t smem;
    const void* fn;
};

template <typename REAL, int THREADS, int R, int TILE, int STAGES, int MINB, int PACK, int UNROLL, int SS>
Variant make_variant(const char* name) {
    Variant v;
    v.name = name;
    v.threads = THREADS; v.r = R; v.tile = TILE; v.stages = STAGES; v.minb = MINB; v.pack = PACK;
    v.smem = sweep_smem_bytes<REAL, THREADS, R, TILE, STAGES, SS>();
    v.fn = (const void*)&sweep_kernel<REAL, THREADS, R, TILE, STAGES, MINB, PACK, UNROLL, SS>;
    return v;
}

// Ordered large -> small work granularity; the automatic choice takes the first one that still
// gives every resident CTA >= 8 tiles (see pick_variant).  Entries after the "auto" prefix are
// only reachable through gravb200_set_variant (ncu A/B evidence, tuning sweeps).
#define V32(T, R, TILE, ST, MB, PK, U, SS) \
    make_variant<float, T, R, TILE, ST, MB, PK, U, SS>("f32_t" #T "_r" #R "_j" #TILE "_s" #ST "_b" #MB "_p" #PK "_u" #U "_m" #SS)
#define V64(T, R, TILE, ST, MB, U) \
    make_variant<double, T, R, TILE, ST, MB, 0, U, 0>("f64_t" #T "_r" #R "_j" #TILE "_s" #ST "_b" #MB "_u" #U)

const std::vector<Variant>& variants_f32() {
    static const std::vector<Variant> v = {
        V32(256, 8, 512, 3, 1, 1, 4, 1),    // 0  auto: large N   (IBLK 2048), fp64 sums in shared memory
        V32(256, 4, 256, 3, 2, 1, 2, 0),    // 1  auto            (IBLK 1024)
        V32(128, 4, 128, 3, 4, 1, 2, 0),    // 2  auto            (IBLK 512)
        V32(128, 2, 64, 4, 4, 1, 2, 0),     // 3  auto: tiny N    (IBLK 256)
        V32(256, 8, 512, 3, 1, 0, 2, 0),    // 4  scalar-FFMA twin of 13 (A/B evidence)
        V32(512, 4, 512, 3, 1, 1, 2, 0),    // 5
        V32(256, 8, 512, 3, 1, 1, 2, 1),    // 6  0 with shared-memory sums
        V32(384, 8, 512, 3, 1, 1, 2, 1),    // 7
        V32(512, 8, 512, 3, 1, 1, 2, 1),    // 8
        V32(512, 8, 512, 3, 1, 1, 1, 1),    // 9
        V32(384, 8, 512, 3, 1, 1, 1, 1),    // 10
        V32(512, 6, 512, 3, 1, 1, 2, 1),    // 11
        V32(512, 4, 512, 3, 1, 1, 2, 1),    // 12
        V32(256, 8, 512, 3, 1, 1, 2, 0),    // 13 register-resident sums, unroll 2 (the round-1 default)
        V32(384, 6, 512, 3, 1, 1, 2, 1),    // 14
        V32(256, 12, 512, 3, 1, 1, 2, 1),   // 15
        V32(256, 8, 512, 3, 2, 1, 2, 1),    // 16
        V32(128, 8, 512, 3, 3, 1, 2, 1),    // 17
    };
    return v;
}
constexpr int kAutoF32 = 4;

const std::vector<Variant>& variants_f64() {
    static const std::vector<Variant> v = {
        V64(256, 2, 256, 3, 2, 4),   // 0 auto: large N
        V64(128, 2, 128, 3, 4, 2),   // 1 auto
        V64(128, 1, 64, 4, 4, 2),    // 2 auto: tiny N
        V64(256, 4, 256, 3, 2, 2),   // 3
        V64(512, 2, 256, 3, 1, 2),   // 4
        V64(256, 2, 256, 3, 2, 2),   // 5
        V64(256, 2, 256, 3, 2, 1),   // 6
        V64(512, 1, 256, 3, 2, 2),   // 7
    };
    return v;
}
constexpr int kAutoF64 = 3;

// Symmetric (Newton's third law) fp32 variants, nbody_sym.cuh.  Selected with variant ids >= kSymBase.
struct SymVariant {
    const char* name;
    int threads, r, tile, stages;
    size_t smem;
    const void* fn;
    const void* fn_split;   // the same kernel with CTA ranges cut at chunk (32 j-bodies) instead of tile granularity; nullptr: not built
};
template <int THREADS, int R, int TILE, int STAGES, int UNROLL, bool WITH_SPLIT>
SymVariant make_sym(const char* name) {
    SymVariant v;
    v.name = name;
    v.threads = THREADS; v.r = R; v.tile = TILE; v.stages = STAGES;
    v.smem = sym_smem_bytes<THREADS, R, TILE, STAGES>();
    v.fn = (const void*)&sym_sweep_kernel<THREADS, R, TILE, STAGES, UNROLL, 0>;
    v.fn_split = nullptr;
    if constexpr (WITH_SPLIT) v.fn_split = (const void*)&sym_sweep_kernel<THREADS, R, TILE, STAGES, UNROLL, 1>;
    return v;
}
#define VSYM(T, R, TILE, ST, U) make_sym<T, R, TILE, ST, U, false>("f32sym_t" #T "_r" #R "_j" #TILE "_s" #ST "_u" #U)
#define VSYMS(T, R, TILE, ST, U) make_sym<T, R, TILE, ST, U, true>("f32sym_t" #T "_r" #R "_j" #TILE "_s" #ST "_u" #U)   // + split twin
constexpr int kSymBase = 100;
const std::vector<SymVariant>& variants_sym() {
    static const std::vector<SymVariant> v = {
        VSYMS(256, 12, 512, 3, 2),  // 100 auto: N >= 65536 on one GPU (IBLK 3072)
        VSYMS(256, 8, 512, 3, 4),   // 101 auto: 16384 <= N < 65536; power-of-two IBLK 2048 (shards of several GPUs)
        VSYMS(256, 8, 256, 3, 2),   // 102
        VSYM(256, 8, 512, 3, 2),    // 103
        VSYM(256, 8, 512, 3, 1),    // 104
        VSYM(256, 10, 512, 3, 2),   // 105
        VSYMS(128, 8, 256, 3, 2),   // 106 auto: 8192 <= N < 16384 (IBLK 1024)
        VSYM(256, 6, 512, 3, 2),    // 107
        VSYM(256, 12, 512, 3, 1),   // 108
        VSYM(256, 12, 512, 3, 4),   // 109
        VSYM(256, 14, 512, 3, 2),   // 110
        VSYMS(256, 12, 256, 3, 2),  // 111
        VSYM(256, 8, 256, 3, 4),    // 112
        VSYM(256, 10, 256, 3, 2),   // 113
        VSYM(256, 16, 512, 3, 1),   // 114
        VSYM(384, 12, 256, 3, 1),   // 115  three warps per scheduler (<= 168 registers)
        VSYM(384, 8, 256, 3, 2),    // 116
        VSYM(384, 10, 256, 3, 1),   // 117
        VSYM(384, 12, 128, 3, 1),   // 118
    };
    return v;
}

template <int THREADS, int R, int TILE, int STAGES, int MINB, int UNROLL, bool WITH_SPLIT>
SymVariant make_sym64(const char* name) {
    SymVariant v;
    v.name = name;
    v.threads = THREADS; v.r = R; v.tile = TILE; v.stages = STAGES;
    v.smem = sym64_smem_bytes<THREADS, TILE, STAGES>();
    v.fn = (const void*)&sym_sweep_kernel_f64<THREADS, R, TILE, STAGES, MINB, UNROLL, 0>;
    v.fn_split = nullptr;
    if constexpr (WITH_SPLIT) v.fn_split = (const void*)&sym_sweep_kernel_f64<THREADS, R, TILE, STAGES, MINB, UNROLL, 1>;
    return v;
}
#define VSYM64(T, R, TILE, ST, MB, U) make_sym64<T, R, TILE, ST, MB, U, false>("f64sym_t" #T "_r" #R "_j" #TILE "_s" #ST "_b" #MB "_u" #U)
#define VSYM64S(T, R, TILE, ST, MB, U) make_sym64<T, R, TILE, ST, MB, U, true>("f64sym_t" #T "_r" #R "_j" #TILE "_s" #ST "_b" #MB "_u" #U)
const std::vector<SymVariant>& variants_sym64() {
    static const std::vector<SymVariant> v = {
        VSYM64(256, 6, 256, 3, 1, 2),   // 100 (IBLK 1536)
        VSYM64S(256, 8, 256, 3, 1, 1),  // 101 auto: N >= 2^14 on one GPU and on shards of several GPUs (IBLK 2048)
        VSYM64S(256, 4, 128, 3, 2, 2),  // 102 auto: medium N (IBLK 1024)
        VSYM64(256, 4, 256, 3, 2, 2),   // 103
        VSYM64(256, 4, 256, 3, 2, 1),   // 104
        VSYM64(256, 2, 256, 3, 2, 2),   // 105
        VSYM64(128, 4, 128, 3, 4, 2),   // 106
        VSYM64(256, 6, 256, 3, 1, 1),   // 107
        VSYM64(256, 6, 128, 3, 1, 2),   // 108
        VSYM64S(256, 8, 128, 3, 1, 1),  // 109
        VSYM64(256, 8, 256, 3, 1, 2),   // 110
    };
    return v;
}
const std::vector<SymVariant>& variants_sym_of(int dtype) { return dtype == GRAVB200_F32 ? variants_sym() : variants_sym64(); }

// Persistent multi-step kernel for universes that fit one SM's shared memory (nbody_small.cuh), ids >= kSmallBase.
struct SmallVariant {
    const char* name;
    int threads, unroll, r;
    const void* fn[2];   // [GRAVB200_F32, GRAVB200_F64]
};
template <int THREADS, int UNROLL, int R>
SmallVariant make_small(const char* name) {
    SmallVariant v;
    v.name = name;
    v.threads = THREADS; v.unroll = UNROLL; v.r = R;
    v.fn[0] = (const void*)&small_steps_kernel<float, THREADS, UNROLL, R>;
    v.fn[1] = (const void*)&small_steps_kernel<double, THREADS, UNROLL, R>;
    return v;
}
#define VSMALL(T, U, R) make_small<T, U, R>("small_t" #T "_u" #U "_r" #R)
constexpr int kSmallBase = 200;
const std::vector<SmallVariant>& variants_small() {
    static const std::vector<SmallVariant> v = {
        VSMALL(256, 4, 4),   // 200 auto: four rows per lane, the quarter-warps read four adjacent j-bodies
        VSMALL(256, 4, 2),   // 201 two rows per lane (bound by shared-memory reads)
        VSMALL(512, 4, 4),   // 202
        VSMALL(256, 2, 4),   // 203
        VSMALL(256, 2, 8),   // 204
        VSMALL(256, 1, 8),   // 205
        VSMALL(512, 2, 4),   // 206
    };
    return v;
}
constexpr size_t kSmallMaxSmem = 227 * 1024;
// Automatic range of the persistent kernel: up to two row groups (64 rows) per CTA in fp32, one in fp64 — 9472 / 4736
// bodies on 148 SMs.  Beyond, fewer CTAs carry more rows each (N = 9600 fp32: 75 CTAs, 72.7 us per step) and the
// symmetric sweep with chunk-granular CTA ranges wins (49.0 us; fp64 N = 5000: 40.0 against 49.6; profiles/r02_mid_n.md).
constexpr int kSmallAutoGroups32 = 2, kSmallAutoGroups64 = 1;

// Rows per shard (contiguous slices of ceil(n / world) rows, the last one short).  The rows
// a shard OWNS (integrates, keeps velocities of) no longer shape the symmetric sweep: its flat (row, tile) list
// of the whole universe is cut into one equal range per shard (stream-K across GPUs, setup_sym), so no block
// alignment is needed and every N keeps the fastest variant.  (Round 1 rounded the shards up to whole blocks:
// +0.8 % rows on seven of eight GPUs at N = 2^20.)
constexpr long long kSplitGainPct = 3;   // whole-tile imbalance (slowest CTA over the average, %) from which the chunk-granular twin pays
constexpr int64_t kSymShardMinN = 32768;   // automatic choice on several shards: symmetric sweep from this N on
int64_t shard_chunk(int64_t n_total, int world, int /*dtype*/) {
    return (n_total + world - 1) / world;
}

// ---------------------------------------------------------------------------------------------
// O(N) helper kernels: layout conversion between the reference's (N,3)+(N,) host arrays
// (np2.py:63-66) and the device float4/double4 state.
// ---------------------------------------------------------------------------------------------
template <typename REAL, typename V4>
__global__ void pack_rm_kernel(const REAL* __restrict__ r3, const REAL* __restrict__ m, V4* out, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    V4 o;
    o.x = r3[3 * i]; o.y = r3[3 * i + 1]; o.z = r3[3 * i + 2];
    o.w = m ? m[i] : out[i].w;
    out[i] = o;
}
template <typename REAL, typename V4>
__global__ void pack_v_kernel(const REAL* __restrict__ v3, V4* out, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    V4 o;
    o.x = v3[3 * i]; o.y = v3[3 * i + 1]; o.z = v3[3 * i + 2]; o.w = 0;
    out[i] = o;
}
template <typename REAL, typename V4>
__global__ void unpack_kernel(const V4* __restrict__ in, REAL* out3, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    V4 v = in[i];
    out3[3 * i] = v.x; out3[3 * i + 1] = v.y; out3[3 * i + 2] = v.z;
}

}  // namespace

// layout of a context's peer-visible flag array (u64 words): barrier generations | sweep done | integrate done | sweep stats (items, ns)
constexpr int kFlagsSweep = kMaxPeers + 1, kFlagsIntegrate = 2 * (kMaxPeers + 1), kFlagsBalance = 3 * (kMaxPeers + 1), kFlagsTotal = 6 * (kMaxPeers + 1);

struct gravb200_ctx {
    int dtype = 0, device = 0, rank = 0, world = 1;
    size_t esz = 4;            // sizeof(REAL)
    int64_t n_total = 0, chunk = 0, n_pad = 0, row0 = 0, n_local = 0;
    double G = 0, T = 0, eps = 0;
    bool uploaded = false, pending = false;   // pending: stage1 issued, stage2 not yet
    bool exchanged = false;                   // the position exchange of the pending step is enqueued
    int front = 0;
    void* pos[2] = {nullptr, nullptr};
    void* vel[2] = {nullptr, nullptr};
    void* acc = nullptr;
    double* partial = nullptr;
    size_t partial_bytes = 0;
    unsigned int* counters = nullptr;
    size_t counters_n = 0;
    void* stage3 = nullptr;     // device staging for (N,3) host arrays
    void* stagem = nullptr;
    unsigned long long* clk = nullptr;   // {SM cycles, ns} of CTA 0 of the last sweep
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t tev[6] = {};   // several shards, symmetric: sweep start | sweep end | integrate end | tail wait end (gravb200_timings ms[5..7])
    bool tev_ok = false;
    bool ev_sweep = false, ev_xchg = false, ev_steps = false;
    ncclComm_t comm = nullptr;
    int sm_count = 0;
    int forced_variant = -1;
    int variant = 0, grid = 0, occ = 0;
    int64_t launches = 0;
    // peer-store exchange (multi-GPU): NVLink-mapped views of every peer's position buffers and flags
    bool peer_connected = false, peer_mode = false;
    unsigned long long* flags = nullptr;          // [kMaxPeers + 1], written by the peers
    int* xerr = nullptr;                          // device flag: barrier timed out
    void* peer_pos[2][kMaxPeers + 1] = {};
    unsigned long long* peer_flags[kMaxPeers + 1] = {};
    double* peer_acc[kMaxPeers + 1] = {};
    bool peer_is_ipc[kMaxPeers + 1] = {};
    unsigned long long epoch = 0;                 // barrier generation, advanced in lockstep on all ranks
    // symmetric sweep on several shards: hand-over flags live behind the barrier flags in the same allocation
    // ([kFlagsSweep + q]: shard q's sweep of step e is done, [kFlagsIntegrate + q]: its integrate), one arrival
    // counter per kernel, and the step generation (lockstep on all ranks)
    unsigned int* done_ctr = nullptr;             // [2]
    unsigned long long* tstart = nullptr;         // earliest CTA start of the running sweep (speed-proportional shares)
    bool sym_balance = false;                     // shares of the flat list follow the measured sweep speed of every GPU
    unsigned long long sym_epoch = 0;
    bool sym_tail_pending = false;                // the last enqueued step has not been followed by its tail wait
    // symmetric sweep (fp32): global fp64 accumulator and the flat-item offsets of the local block rows
    double* acc64 = nullptr;
    long long* row_start = nullptr;
    size_t row_start_n = 0;
    bool use_sym = false;
    bool pdl = false;              // one shard, symmetric step: programmatic dependent launch of the integrate kernel and the next sweep
    int split_mode = -1;           // CTA ranges of the symmetric sweep at chunk granularity: -1 automatic, 0 never, 1 wherever the variant has the twin
    bool sym_split = false;        // what the current set-up uses
    bool split_weighted = true;    // the twin cuts by cost (diagonal chunks are cheaper), not by chunk count
    long long sym_cost_lo = 0, sym_cost_hi = 0;   // this shard's share in cost units
    int sym_w[2] = {4, 3};
    long long sym_min_n = 4096;    // automatic choice: symmetric sweep from this N on (below, the persistent kernel takes everything it fits)
    int sym_variant = 0, sym_blocks = 0, sym_gblocks = 0;
    long long sym_total = 0, sym_lo = 0, sym_hi = 0;   // flat items of the universe, this shard's share
    // persistent small-N kernel (nbody_small.cuh): geometry, grid barrier counter and its host mirror
    bool use_small = false;
    int small_variant = 0, small_ng = 1, small_rpc = 0, small_slice = 0;
    size_t small_smem = 0;
    unsigned long long* gbar = nullptr;
    unsigned long long gbar_count = 0;
    // gravb200_steps on one GPU, launch-bound sizes: kGraphSteps steps captured once per front-buffer parity
    cudaGraphExec_t step_graph[2] = {nullptr, nullptr};
    int64_t step_graph_kernels = 0;   // kernel nodes in one of them
};

namespace {

const std::vector<Variant>& variants_of(int dtype) {
    return dtype == GRAVB200_F32 ? variants_f32() : variants_f64();
}

int occupancy_of(const Variant& v, int* occ) {
    CU(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, v.fn, v.threads, v.smem));
    if (*occ < 1) return fail(GRAVB200_ECUDA, "variant %s does not fit on an SM", v.name);
    return 0;
}

long long tiles_of(const gravb200_ctx* c, const Variant& v) {
    const long long iblk = (long long)v.threads * v.r;
    const long long nib = (c->n_local + iblk - 1) / iblk;
    const long long njt = (c->n_total + v.tile - 1) / v.tile;
    return nib * njt;
}

// CTAs to launch for `tiles` flat tiles on `slots` resident CTA slots: all slots when there is enough
// work, otherwise few enough CTAs that each still gets >= 4 tiles (fewer contributors per split
// i-block, less fixed cost per unit of work)
long long grid_for(long long tiles, long long slots) {
    long long g = slots;
    if (tiles < 4 * slots) g = std::max<long long>(1, tiles / 4);
    if (g > slots) g = slots;
    if (g > tiles) g = tiles;
    return std::max<long long>(g, 1);
}

// choose the variant and size its workspace
// symmetric sweep set-up: block-row offsets, accumulator, grid
int setup_sym(gravb200_ctx* c, int sv) {
    const SymVariant& v = variants_sym_of(c->dtype)[sv];
    const int iblk = v.threads * v.r;
    const int Bt = (int)((c->n_total + iblk - 1) / iblk);
    // one shard: its block rows are all block rows.  Several shards: every shard walks the GLOBAL list and takes
    // its equal share of it (sym_items), independent of the rows it owns.
    const int nib = Bt;
    // rs[0 .. nib]: flat tile offset of every block row; rs[nib + 1 .. 2 nib + 1]: the same prefix in COST units for
    // the cost-weighted cut of the chunk-granular twins — a chunk (32 j-bodies x one block row) of a diagonal tile is
    // evaluated ordered and takes ~3/4 (fp32) or ~4/5 (fp64) of a symmetric chunk's time; a row's diagonal tiles come first
    int w_sym = c->dtype == GRAVB200_F32 ? 4 : 5, w_diag = c->dtype == GRAVB200_F32 ? 3 : 4;
    if (const char* e = getenv("GRAVB200_SPLIT_W")) {   // "w_sym,w_diag": tuning sweeps (scripts/weighted_ab.py)
        int a = 0, b = 0;
        if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 1 && b >= 1 && a <= 64 && b <= 64) { w_sym = a; w_diag = b; }
    }
    const long long ch = v.tile / 32;
    std::vector<long long> rs(2 * ((size_t)nib + 1), 0);
    long long* rc = rs.data() + nib + 1;
    for (int i = 0; i < nib; ++i) {
        const long long row_tiles = sym_row_tiles(c->n_total, iblk, v.tile, Bt, i);
        rs[i + 1] = rs[i] + row_tiles;
        rc[i + 1] = rc[i] + sym_cost_in_row(c->n_total, iblk, v.tile, i, row_tiles, w_sym, w_diag);
    }
    if (rs.size() > c->row_start_n) {
        if (c->row_start) CU(cudaFree(c->row_start));
        c->row_start = nullptr;
        CU(cudaMalloc(&c->row_start, rs.size() * sizeof(long long)));
        c->row_start_n = rs.size();
    }
    CU(cudaMemcpyAsync(c->row_start, rs.data(), rs.size() * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));   // rs is a local
    if (!c->acc64) {
        CU(cudaMalloc(&c->acc64, (size_t)c->n_pad * 4 * sizeof(double)));
        CU(cudaMemsetAsync(c->acc64, 0, (size_t)c->n_pad * 4 * sizeof(double), c->stream));
    }
    CU(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.fn, v.threads, v.smem));
    if (occ < 1) return fail(GRAVB200_ECUDA, "symmetric variant %s does not fit on an SM", v.name);
    c->sym_variant = sv;
    c->sym_blocks = nib;
    c->sym_gblocks = Bt;
    c->sym_total = rs[nib];
    c->sym_lo = sk_lo(rs[nib], c->rank, c->world);
    c->sym_hi = sk_lo(rs[nib], c->rank + 1, c->world);
    c->occ = occ;
    // Few tiles per CTA (mid-sized universes, small shards): cut the CTA ranges at chunk granularity, so the slowest
    // CTA carries 1 / CHUNKS of a tile more than the average instead of up to a whole tile, and more CTAs than
    // tiles can work (N = 24576 fp32: 277.6 -> 200.8 us per step, profiles/r02_split_ab.md).  The twin's ring loop
    // is the same instructions in another order of ptxas' choosing and 1 - 2.6 % slower (N = 2^18: 19.19 -> 19.68 ms),
    // so it is used where whole tiles leave the slowest CTA more than kSplitGainPct % above the average; the large
    // universes keep the kernel that was profiled.
    const long long share_tiles = c->sym_hi - c->sym_lo, slots = (long long)occ * c->sm_count;
    const long long slowest = (share_tiles + slots - 1) / slots;   // tiles of the slowest CTA with whole-tile ranges
    // cost of this share and of the slowest CTA with whole-tile ranges (all of its tiles symmetric)
    auto cost_at = [&](long long t) -> long long {   // cost position of the start of flat tile t
        if (t >= rs[nib]) return rc[nib];
        int a = (int)(std::upper_bound(rs.begin(), rs.begin() + nib + 1, t) - rs.begin()) - 1;
        return rc[a] + sym_cost_in_row(c->n_total, iblk, v.tile, a, t - rs[a], w_sym, w_diag);
    };
    c->sym_cost_lo = cost_at(c->sym_lo);
    c->sym_cost_hi = cost_at(c->sym_hi);
    c->sym_w[0] = w_sym; c->sym_w[1] = w_diag;
    const long long share_cost = c->sym_cost_hi - c->sym_cost_lo;
    c->sym_split = v.fn_split != nullptr &&
                   (c->split_mode == 1 || (c->split_mode < 0 && slowest * ch * w_sym * slots * 100 > share_cost * (100 + kSplitGainPct)));
    if (c->sym_split) CU(cudaFuncSetAttribute(v.fn_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
    const long long units = c->sym_split ? share_tiles * ch : share_tiles;
    c->grid = (int)std::max<long long>(1, std::min<long long>(slots, units));
    // several shards with long sweeps (>= 16 tiles per CTA): the shares follow the measured speed of each GPU
    // (nbody_sym.cuh, SymBalance); GRAVB200_BALANCE=0 keeps the equal shares
    const char* bal_env = getenv("GRAVB200_BALANCE");
    c->sym_balance = c->world > 1 && (c->sym_hi - c->sym_lo) >= 16LL * c->grid && !(bal_env && bal_env[0] == '0');
    if (c->world > 1)   // a new variant counts tiles of another size: forget the published speeds
        CU(cudaMemsetAsync(c->flags + kFlagsBalance, 0, 3 * (kMaxPeers + 1) * sizeof(unsigned long long), c->stream));
    c->use_sym = true;
    return 0;
}

struct SmallGeometry { int grid, rpc, ng, slice, nsl; size_t smem; };
// 0: fits; 1: more than 32 * NWARPS rows per CTA needed; 2: does not fit in shared memory
int small_geometry(long long n, int dtype, int sm_count, const SmallVariant& v, SmallGeometry* g) {
    const int nw = v.threads / 32;
    int ng = 1;
    while ((n + 32LL * ng - 1) / (32LL * ng) > sm_count && ng < nw) ng *= 2;
    if ((n + 32LL * ng - 1) / (32LL * ng) > sm_count) return 1;
    g->ng = ng;
    g->grid = (int)((n + 32LL * ng - 1) / (32LL * ng));
    g->rpc = (int)((n + g->grid - 1) / g->grid);
    g->rpc += g->rpc & 1;
    g->nsl = nw / ng;
    const int q = v.r * v.unroll;
    g->slice = (int)(((n + g->nsl - 1) / g->nsl + q - 1) / q * q);
    g->smem = dtype == GRAVB200_F32 ? small_smem_bytes<float>(g->nsl * g->slice, g->nsl, ng) : small_smem_bytes<double>(g->nsl * g->slice, g->nsl, ng);
    return g->smem > kSmallMaxSmem ? 2 : 0;
}

int setup_small(gravb200_ctx* c, int sv, bool forced) {
    const SmallVariant& v = variants_small()[sv];
    SmallGeometry g;
    const int fit = small_geometry(c->n_total, c->dtype, c->sm_count, v, &g);
    if (fit == 1) return forced ? fail(GRAVB200_EINVAL, "variant %s: %lld bodies need more than %d rows per CTA", v.name, (long long)c->n_total, v.threads) : 1;
    if (fit == 2) return forced ? fail(GRAVB200_EINVAL, "variant %s: %lld bodies do not fit in shared memory (%zu bytes)", v.name, (long long)c->n_total, g.smem) : 1;
    if (!forced && g.ng > (c->dtype == GRAVB200_F32 ? kSmallAutoGroups32 : kSmallAutoGroups64)) return 1;   // beyond, the symmetric sweep wins
    const void* fn = v.fn[c->dtype == GRAVB200_F32 ? 0 : 1];
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, v.threads, g.smem));
    if (occ < 1) return forced ? fail(GRAVB200_ECUDA, "variant %s does not fit on an SM", v.name) : 1;
    if (!c->gbar) {
        CU(cudaMalloc(&c->gbar, sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->gbar, 0, sizeof(unsigned long long), c->stream));
        c->gbar_count = 0;
    }
    c->small_variant = sv; c->small_ng = g.ng; c->small_rpc = g.rpc; c->small_slice = g.slice; c->small_smem = g.smem;
    c->grid = g.grid;
    c->occ = occ;
    c->use_small = true;
    return 0;
}

int launch_small(gravb200_ctx* c, int k, int integrate) {
    const SmallVariant& v = variants_small()[c->small_variant];
    SmallParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.pos[0] = c->pos[0]; sp.pos[1] = c->pos[1];
    sp.vel[0] = c->vel[0]; sp.vel[1] = c->vel[1];
    sp.acc = c->acc;
    sp.gbar = c->gbar;
    sp.gbar_base = c->gbar_count;
    sp.error = c->xerr;
    sp.n = c->n_total;
    sp.rpc = c->small_rpc; sp.ng = c->small_ng; sp.slice = c->small_slice;
    sp.front = c->front;
    sp.k = k;
    sp.integrate = integrate;
    sp.G = c->G; sp.T = c->T;
    sp.eps2_f = (float)(c->eps * c->eps);
    sp.eps2_d = c->eps * c->eps;
    sp.clk = c->clk;
    void* args[] = {&sp};
    const void* fn = v.fn[c->dtype == GRAVB200_F32 ? 0 : 1];
    if (k > 1) {   // the grid barrier needs every CTA resident: a cooperative launch guarantees it or fails
        CU(cudaLaunchCooperativeKernel(fn, dim3(c->grid), dim3(v.threads), args, c->small_smem, c->stream));
        c->gbar_count += (unsigned long long)c->grid * (unsigned long long)(k - 1);
    } else {
        CU(cudaLaunchKernel(fn, dim3(c->grid), dim3(v.threads), args, c->small_smem, c->stream));
    }
    c->launches++;
    return 0;
}

// CUDA graph of kGraphSteps (even: the front buffer is the same before and after) full steps.  Kernel
// arguments depend only on the front-buffer parity, the variant and G/T/eps, so a graph stays valid until
// one of those changes (pick_variant, upload).
constexpr int kGraphSteps = 8;
constexpr int64_t kGraphMaxN = 32768;   // above, a step takes > 0.3 ms and launch latency is noise

void graph_invalidate(gravb200_ctx* c) {
    for (auto& g : c->step_graph) {
        if (g) cudaGraphExecDestroy(g);
        g = nullptr;
    }
}

int pick_variant(gravb200_ctx* c) {
    graph_invalidate(c);
    c->use_sym = false;
    c->use_small = false;
    if (c->forced_variant >= kSmallBase) {
        if (c->forced_variant - kSmallBase >= (int)variants_small().size()) return fail(GRAVB200_EINVAL, "variant %d out of range", c->forced_variant);
        if (c->world != 1) return fail(GRAVB200_EINVAL, "the persistent small-N kernel runs on one shard only");
        return setup_small(c, c->forced_variant - kSmallBase, true);
    }
    if (c->forced_variant < 0 && c->world == 1 && c->n_local > 0) {
        const int rc = setup_small(c, 0, false);
        if (rc <= 0) return rc;   // 0: chosen, < 0: CUDA error; 1: not eligible, go on
    }
    // symmetric sweep: forced (ids >= kSymBase) or automatic once there are enough body-blocks.  Several
    // shards need the peer-store exchange (the owner of a row reads the other shards' partial sums over
    // NVLink) and block-aligned shards.
    const int n_sym = (int)variants_sym_of(c->dtype).size();
    if (c->forced_variant >= kSymBase + n_sym) return fail(GRAVB200_EINVAL, "variant %d out of range", c->forced_variant);
    if (c->world == 1 && c->n_local > 0) {
        int sv = -1;
        if (c->forced_variant >= kSymBase) sv = c->forced_variant - kSymBase;
        else if (c->forced_variant < 0 && c->n_total >= c->sym_min_n)
            sv = c->dtype == GRAVB200_F32 ? (c->n_total >= 65536 ? 0 : (c->n_total >= 9216 ? 1 : 6))
                                          : (c->n_total >= 10240 ? 1 : 2);   // profiles/r01_sym*_variants_sweep*.txt; with chunk-granular ranges: r02_split_ab.md, r02_mid_n.md
        if (sv >= 0) return setup_sym(c, sv);
    } else if (c->world > 1 && c->peer_mode && c->acc64) {
        int sv = -1;
        if (c->forced_variant >= kSymBase) sv = c->forced_variant - kSymBase;
        else if (c->forced_variant < 0 && c->n_total >= kSymShardMinN) {
            // a shard does 1 / world of the universe's pairs: choose as one GPU would for a universe with that many
            const double n_eff = (double)c->n_total / std::sqrt((double)c->world);
            sv = c->dtype == GRAVB200_F32 ? (n_eff >= 65536 ? 0 : (n_eff >= 16384 ? 1 : 6)) : (n_eff >= 16384 ? 1 : 2);
        }
        if (sv >= 0) return setup_sym(c, sv);
    } else if (c->forced_variant >= kSymBase) {
        return fail(GRAVB200_EINVAL, "on several shards the symmetric variants need the peer-store exchange");
    }
    const auto& vs = variants_of(c->dtype);
    const int n_auto = c->dtype == GRAVB200_F32 ? kAutoF32 : kAutoF64;
    int pick = n_auto - 1, occ = 0;
    if (c->forced_variant >= 0) {
        pick = c->forced_variant;
        int rc = occupancy_of(vs[pick], &occ);
        if (rc) return rc;
    } else {
        // Cost model (relative units): a CTA gets 1/occ of its SM, tiles run at the variant's measured
        // inner-loop speed (profiles/r01_sweep*_variants.jsonl, large N), the slowest CTA runs
        // ceil(tiles / grid) tiles, and every launch pays a fixed cost plus one fix-up term per
        // contributor of a split i-block.  Small N thereby moves to finer-grained variants by itself.
        static const double kSpeed32[kAutoF32] = {1.00, 0.90, 0.87, 0.76};
        static const double kSpeed64[kAutoF64] = {1.00, 0.95, 0.85};
        double best_cost = 0;
        for (int i = 0; i < n_auto; ++i) {
            int o = 0;
            int rc = occupancy_of(vs[i], &o);
            if (rc) return rc;
            const double speed = c->dtype == GRAVB200_F32 ? kSpeed32[i] : kSpeed64[i];
            const long long tiles = tiles_of(c, vs[i]);
            if (tiles <= 0) continue;
            const long long g = grid_for(tiles, (long long)o * c->sm_count);
            const double per_tile = (double)vs[i].threads * vs[i].r * vs[i].tile * o / speed;   // SM-share units
            const double rounds = (double)((tiles + g - 1) / g);
            const double njt = (double)((c->n_total + vs[i].tile - 1) / vs[i].tile);
            const double contributors = std::max(1.0, njt / std::max(1.0, (double)tiles / (double)g));
            const double fixed = 150000.0 + 6000.0 * contributors;   // ~8 us launch/prologue + ~0.3 us per contributor
            const double cost = rounds * per_tile + fixed;
            if (i == 0 || cost < best_cost) { best_cost = cost; pick = i; }
        }
        int rc = occupancy_of(vs[pick], &occ);
        if (rc) return rc;
    }
    const Variant& v = vs[pick];
    const long long total = tiles_of(c, v);
    long long grid = grid_for(total, (long long)occ * c->sm_count);
    c->variant = pick;
    c->occ = occ;
    c->grid = (int)grid;
    const long long iblk = (long long)v.threads * v.r;
    const size_t need = (size_t)2 * (size_t)std::max<long long>(grid, 1) * iblk * 4 * sizeof(double);
    if (need > c->partial_bytes) {
        if (c->partial) CU(cudaFree(c->partial));
        c->partial = nullptr;
        CU(cudaMalloc(&c->partial, need));
        c->partial_bytes = need;
    }
    const size_t nib = (size_t)((c->n_local + iblk - 1) / iblk);
    if (nib > c->counters_n) {
        if (c->counters) CU(cudaFree(c->counters));
        c->counters = nullptr;
        CU(cudaMalloc(&c->counters, std::max<size_t>(nib, 1) * sizeof(unsigned int)));
        c->counters_n = nib;
    }
    if (c->counters_n) CU(cudaMemsetAsync(c->counters, 0, c->counters_n * sizeof(unsigned int), c->stream));
    return 0;
}

int peer_barrier(gravb200_ctx* c);

int launch_sweep(gravb200_ctx* c, int integrate) {
    if (c->grid <= 0 || (c->n_local <= 0 && !(c->use_sym && c->world > 1))) return 0;
    if (c->use_small) return launch_small(c, 1, integrate);
    if (c->use_sym) {
        const SymVariant& sv = variants_sym_of(c->dtype)[c->sym_variant];
        const bool multi = c->world > 1;
        SymParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.pos_front = (const float4*)c->pos[c->front];
        sp.pos_front_d = (const double4*)c->pos[c->front];
        sp.eps2_d = c->eps * c->eps;
        sp.acc64 = c->acc64;
        sp.row_start = c->row_start;
        sp.n_total = c->n_total;
        sp.row0 = 0;                    // the flat list spans the universe; a shard takes items, not rows
        sp.n_local = c->n_total;
        sp.n_iblocks = c->sym_blocks;
        sp.n_gblocks = c->sym_gblocks;
        sp.gblock0 = 0;
        sp.eps2_f = (float)(c->eps * c->eps);
        sp.clk = c->clk;
        sp.item_lo = c->sym_lo;
        sp.item_hi = c->sym_hi;
        // device-computed shares (speed-proportional) keep the count-based cut; so do the R = 4 fp64 variants, whose
        // diagonal chunks are no cheaper than their symmetric ones (N = 7000: 55.5 us by count, 59.7 weighted 5:4;
        // the R = 8 variant at 12 288: 136.4 -> 125.6; fp32 R = 8, 4:3: 64.9 -> 62.0 at 12 288, +-0.5 % elsewhere;
        // profiles/r02_weighted_scan*.jsonl)
        if (c->sym_split && c->split_weighted && !c->sym_balance && (c->dtype == GRAVB200_F32 || sv.r >= 8)) {
            sp.row_cost = c->row_start + c->sym_blocks + 1;
            sp.cost_lo = c->sym_cost_lo; sp.cost_hi = c->sym_cost_hi;
            sp.w_sym = c->sym_w[0]; sp.w_diag = c->sym_w[1];
        }
        IntegrateParams ip;
        memset(&ip, 0, sizeof(ip));
        ip.sp.pos_front = c->pos[c->front];
        ip.sp.pos_back = c->pos[c->front ^ 1];
        ip.sp.vel_front = c->vel[c->front];
        ip.sp.vel_back = c->vel[c->front ^ 1];
        ip.sp.acc = c->acc;
        ip.sp.n_total = c->n_total;
        ip.sp.row0 = c->row0;
        ip.sp.n_local = c->n_local;
        ip.sp.G = c->G;
        ip.sp.T = c->T;
        ip.sp.integrate = integrate;
        ip.sp.n_peers = 0;
        ip.sp.clk = c->clk;   // [2048], [2049]: end of the wait for the peers' sweeps, latest CTA end (several shards)
        ip.acc64 = c->acc64;
        ip.n_src = 0;
        if (multi) {
            // step generation e: the sweep starts once every shard's integrate of generation e - 1 is done (all r'
            // have landed here, all rows of this accumulator are zero again) and publishes "sweep e done"; the
            // integrate kernel starts once every shard has published that, adds the partial sums of all shards for
            // the rows it owns over NVLink, stores r' to every peer and publishes "integrate e done".  No barrier
            // launches, no memset: two kernels per step.
            const unsigned long long e = ++c->sym_epoch;
            PeerSync& a = sp.sync;
            a.wait_flags = c->flags + kFlagsIntegrate;
            a.wait_value = e - 1;
            a.signal_value = e;
            a.done_ctr = c->done_ctr;
            a.rank = c->rank; a.world = c->world; a.error = c->xerr;
            PeerSync& b = ip.sync;
            b.wait_flags = c->flags + kFlagsSweep;
            b.wait_value = e;
            b.signal_value = e;
            b.done_ctr = c->done_ctr + 1;
            b.rank = c->rank; b.world = c->world; b.error = c->xerr;
            for (int q = 0; q < c->world; ++q) {
                a.signal_flags[q] = c->peer_flags[q] + kFlagsSweep;
                b.signal_flags[q] = c->peer_flags[q] + kFlagsIntegrate;
                ip.acc_src[ip.n_src++] = c->peer_acc[q];
                if (q != c->rank) ip.sp.peer_back[ip.sp.n_peers++] = c->peer_pos[c->front ^ 1][q];
            }
            if (c->sym_balance) {
                sp.bal.stats = c->flags + kFlagsBalance;
                sp.bal.tstart = c->tstart;
                for (int q = 0; q < c->world; ++q) sp.bal.peer_stats[q] = c->peer_flags[q] + kFlagsBalance;
            }
        }
        void* sargs[] = {&sp};
        const bool trace = multi && c->tev[0];
        if (trace) CU(cudaEventRecord(c->tev[0], c->stream));
        // one shard: both kernels carry the programmatic-serialization attribute, so each one's launch and prologue
        // overlap its predecessor's tail (griddep_wait / griddep_launch in the kernels)
        cudaLaunchAttribute pdl_attr[1];
        pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof(lc));
        lc.gridDim = dim3(c->grid); lc.blockDim = dim3(sv.threads); lc.dynamicSmemBytes = sv.smem; lc.stream = c->stream;
        lc.attrs = pdl_attr; lc.numAttrs = (c->pdl && !multi) ? 1 : 0;
        CU(cudaLaunchKernelExC(&lc, c->sym_split ? sv.fn_split : sv.fn, sargs));
        if (trace) CU(cudaEventRecord(c->tev[1], c->stream));
        const unsigned gb = (unsigned)std::max<long long>(1, (c->n_local + 255) / 256);
        void* iargs[] = {&ip};
        lc.gridDim = dim3(gb); lc.blockDim = dim3(256); lc.dynamicSmemBytes = 0;
        CU(cudaLaunchKernelExC(&lc, c->dtype == GRAVB200_F32 ? (const void*)&sym_integrate_kernel<float> : (const void*)&sym_integrate_kernel<double>, iargs));
        if (trace) CU(cudaEventRecord(c->tev[2], c->stream));
        c->sym_tail_pending = multi;
        c->launches += 2;
        return 0;
    }
    const Variant& v = variants_of(c->dtype)[c->variant];
    SweepParams p;
    p.pos_front = c->pos[c->front];
    p.pos_back = c->pos[c->front ^ 1];
    p.vel_front = c->vel[c->front];
    p.vel_back = c->vel[c->front ^ 1];
    p.acc = c->acc;
    p.partial = c->partial;
    p.counters = c->counters;
    p.n_total = c->n_total;
    p.row0 = c->row0;
    p.n_local = c->n_local;
    const long long iblk = (long long)v.threads * v.r;
    p.n_iblocks = (int)((c->n_local + iblk - 1) / iblk);
    p.n_jtiles = (int)((c->n_total + v.tile - 1) / v.tile);
    p.G = c->G;
    p.T = c->T;
    p.eps2_f = (float)(c->eps * c->eps);
    p.eps2_d = c->eps * c->eps;
    p.integrate = integrate;
    p.clk = c->clk;
    p.n_peers = 0;
    if (c->peer_mode) {
        for (int q = 0; q < c->world; ++q)
            if (q != c->rank) p.peer_back[p.n_peers++] = c->peer_pos[c->front ^ 1][q];
    }
    void* args[] = {&p};
    CU(cudaLaunchKernel(v.fn, dim3(c->grid), dim3(v.threads), args, v.smem, c->stream));
    c->launches++;
    return 0;
}

int peer_barrier(gravb200_ctx* c) {
    BarrierParams b;
    b.my_flags = c->flags;
    for (int q = 0; q < c->world; ++q) b.peer_flags[q] = c->peer_flags[q];
    b.rank = c->rank;
    b.world = c->world;
    b.step = ++c->epoch;
    b.timeout_ns = 60ull * 1000 * 1000 * 1000;
    b.error = c->xerr;
    exchange_barrier_kernel<<<1, 32, 0, c->stream>>>(b);
    CU(cudaGetLastError());
    c->launches++;
    return 0;
}

int step_graph_for(gravb200_ctx* c, cudaGraphExec_t* out) {
    const int f = c->front;
    if (!c->step_graph[f]) {
        const int64_t launches0 = c->launches;
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        for (int s = 0; s < kGraphSteps && !rc; ++s) {
            rc = launch_sweep(c, 1);
            c->front ^= 1;
        }
        const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        c->step_graph_kernels = c->launches - launches0;
        c->launches = launches0;   // nothing ran yet
        c->front = f;
        if (rc || e != cudaSuccess) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            return rc ? rc : fail(GRAVB200_ECUDA, "stream capture of the step graph failed: %s", cudaGetErrorString(e));
        }
        const cudaError_t ei = cudaGraphInstantiate(&c->step_graph[f], g, 0);
        cudaGraphDestroy(g);
        if (ei != cudaSuccess) {
            c->step_graph[f] = nullptr;
            return fail(GRAVB200_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
        }
    }
    *out = c->step_graph[f];
    return 0;
}

int check_barrier_error(gravb200_ctx* c) {
    if (!c->peer_mode) return 0;
    int e = 0;
    CU(cudaMemcpy(&e, c->xerr, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) return fail(GRAVB200_ECUDA, "peer-store exchange: a peer GPU did not reach the step barrier within 60 s");
    return 0;
}

// tail of a symmetric multi-shard step: wait (on the device) until every shard's integrate kernel is done, i.e.
// all r' have landed in this GPU's back buffer.  Between consecutive steps of gravb200_steps the next sweep does
// this wait itself; the host needs it before it looks at the state (stage2, end of steps).
int sym_tail_wait(gravb200_ctx* c) {
    if (!c->sym_tail_pending) return 0;
    PeerSync w;
    memset(&w, 0, sizeof(w));
    w.wait_flags = c->flags + kFlagsIntegrate;
    w.wait_value = c->sym_epoch;
    w.rank = c->rank; w.world = c->world; w.error = c->xerr;
    peer_wait_kernel<<<1, 32, 0, c->stream>>>(w);
    CU(cudaGetLastError());
    c->launches++;
    c->sym_tail_pending = false;
    if (c->tev[0]) { CU(cudaEventRecord(c->tev[3], c->stream)); c->tev_ok = true; }
    return 0;
}

int exchange(gravb200_ctx* c, bool last = true) {
    if (c->world == 1) return 0;
    if (c->peer_mode) {   // the data already travelled in the sweep's epilogue / the integrate kernel
        if (c->use_sym) return last ? sym_tail_wait(c) : 0;
        return peer_barrier(c);
    }
    char* back = (char*)c->pos[c->front ^ 1];
    const size_t count = (size_t)c->chunk * 4;   // scalars per shard
    NC(g_nccl.AllGather(back + (size_t)c->rank * c->chunk * 4 * c->esz, back, count,
                        c->dtype == GRAVB200_F32 ? ncclFloat32 : ncclFloat64, c->comm, c->stream));
    return 0;
}

// An upload replaces the state a pending stage1 was computed from: the step is dropped.
int drop_pending(gravb200_ctx* c) {
    if (c->pending && c->use_sym && c->world > 1 && c->acc64) {
        // the dropped step ran sweep + integrate everywhere (the integrate kernels zeroed the accumulators again);
        // what follows may overwrite positions the peers still read, so all shards meet first.  Every rank drops
        // the step (uploads and repeated stage1 calls are collective): the generations stay in lockstep.
        int rc = sym_tail_wait(c);
        if (rc) return rc;
        rc = peer_barrier(c);
        if (rc) return rc;
    }
    c->pending = false;
    return 0;
}

template <typename REAL, typename V4>
int upload_impl(gravb200_ctx* c, const void* r, const void* v, const void* m) {
    const long long n = c->n_total;
    const int tb = 256;
    const unsigned gb = (unsigned)((n + tb - 1) / tb);
    if (r) {
        CU(cudaMemcpyAsync(c->stage3, r, (size_t)n * 3 * sizeof(REAL), cudaMemcpyHostToDevice, c->stream));
        if (m) CU(cudaMemcpyAsync(c->stagem, m, (size_t)n * sizeof(REAL), cudaMemcpyHostToDevice, c->stream));
        pack_rm_kernel<REAL, V4><<<gb, tb, 0, c->stream>>>((const REAL*)c->stage3, m ? (const REAL*)c->stagem : nullptr,
                                                          (V4*)c->pos[c->front], n);
        CU(cudaGetLastError());
        c->launches++;
    }
    if (v && c->n_local > 0) {
        // only this shard's rows are kept
        const REAL* vsrc = (const REAL*)v + (size_t)c->row0 * 3;
        CU(cudaMemcpyAsync(c->stage3, vsrc, (size_t)c->n_local * 3 * sizeof(REAL), cudaMemcpyHostToDevice, c->stream));
        const unsigned gl = (unsigned)((c->n_local + tb - 1) / tb);
        pack_v_kernel<REAL, V4><<<gl, tb, 0, c->stream>>>((const REAL*)c->stage3, (V4*)c->vel[c->front], c->n_local);
        CU(cudaGetLastError());
        c->launches++;
    }
    return 0;
}

template <typename REAL, typename V4>
int download_impl(gravb200_ctx* c, void* r, void* v, void* a) {
    const int tb = 256;
    if (r) {
        const long long n = c->n_total;
        unpack_kernel<REAL, V4><<<(unsigned)((n + tb - 1) / tb), tb, 0, c->stream>>>((const V4*)c->pos[c->front], (REAL*)c->stage3, n);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(r, c->stage3, (size_t)n * 3 * sizeof(REAL), cudaMemcpyDeviceToHost, c->stream));
        c->launches++;
    }
    const long long nl = c->n_local;
    if (v && nl > 0) {
        unpack_kernel<REAL, V4><<<(unsigned)((nl + tb - 1) / tb), tb, 0, c->stream>>>((const V4*)c->vel[c->front], (REAL*)c->stage3, nl);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(v, c->stage3, (size_t)nl * 3 * sizeof(REAL), cudaMemcpyDeviceToHost, c->stream));
        c->launches++;
    }
    if (a && nl > 0) {
        unpack_kernel<REAL, V4><<<(unsigned)((nl + tb - 1) / tb), tb, 0, c->stream>>>((const V4*)c->acc, (REAL*)c->stage3, nl);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(a, c->stage3, (size_t)nl * 3 * sizeof(REAL), cudaMemcpyDeviceToHost, c->stream));
        c->launches++;
    }
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// Own rows only (several shards, each fed by its own host): positions and velocities of rows
// [row0, row0 + n_local) come from the host, the masses stay, and the device exchange hands the new positions
// to every other shard — NVLink copies into the peers' front buffers + the flag barrier (peer mode) or an
// in-place all-gather (NCCL mode).  Collective: every shard of the universe must call it.
template <typename REAL, typename V4>
int upload_rows_impl(gravb200_ctx* c, const void* r_own, const void* v_own) {
    const long long nl = c->n_local;
    const int tb = 256;
    const unsigned gl = (unsigned)((nl + tb - 1) / tb);
    V4* front = (V4*)c->pos[c->front];
    if (nl > 0) {
        CU(cudaMemcpyAsync(c->stage3, r_own, (size_t)nl * 3 * sizeof(REAL), cudaMemcpyHostToDevice, c->stream));
        pack_rm_kernel<REAL, V4><<<gl, tb, 0, c->stream>>>((const REAL*)c->stage3, nullptr, front + c->row0, nl);
        CU(cudaGetLastError());
        c->launches++;
        if (v_own) {
            REAL* vstage = (REAL*)c->stage3;   // reused after the pack above (stream ordered)
            CU(cudaMemcpyAsync(vstage, v_own, (size_t)nl * 3 * sizeof(REAL), cudaMemcpyHostToDevice, c->stream));
            pack_v_kernel<REAL, V4><<<gl, tb, 0, c->stream>>>((const REAL*)vstage, (V4*)c->vel[c->front], nl);
            CU(cudaGetLastError());
            c->launches++;
        }
    }
    // the host buffers are the caller's again; everything below is device to device and only ENQUEUED (a host
    // thread that drives several shards calls them one after the other and must not block on a peer here)
    CU(cudaStreamSynchronize(c->stream));
    if (c->world > 1) {
        if (c->peer_mode) {
            if (nl > 0)
                for (int q = 0; q < c->world; ++q)
                    if (q != c->rank)
                        CU(cudaMemcpyAsync((V4*)c->peer_pos[c->front][q] + c->row0, front + c->row0, (size_t)nl * sizeof(V4),
                                           cudaMemcpyDefault, c->stream));
            int rc = peer_barrier(c);   // all rows have landed everywhere before any shard sweeps
            if (rc) return rc;
        } else {
            char* base = (char*)c->pos[c->front];
            NC(g_nccl.AllGather(base + (size_t)c->rank * c->chunk * sizeof(V4), base, (size_t)c->chunk * 4,
                                c->dtype == GRAVB200_F32 ? ncclFloat32 : ncclFloat64, c->comm, c->stream));
        }
    }
    return 0;
}

template <typename REAL, typename V4>
int download_rows_impl(gravb200_ctx* c, void* r_own, void* v_own, void* a_own) {
    const int tb = 256;
    const long long nl = c->n_local;
    if (nl <= 0) return 0;
    const unsigned gl = (unsigned)((nl + tb - 1) / tb);
    // three disjoint thirds of the staging buffer when it is large enough (several shards), else one after the other
    const bool apart = (size_t)c->n_total >= (size_t)3 * (size_t)nl;
    const V4* src[3] = {(const V4*)c->pos[c->front] + c->row0, (const V4*)c->vel[c->front], (const V4*)c->acc};
    void* dst[3] = {r_own, v_own, a_own};
    for (int k = 0; k < 3; ++k) {
        if (!dst[k]) continue;
        REAL* stage = (REAL*)c->stage3 + (apart ? (size_t)k * nl * 3 : 0);
        unpack_kernel<REAL, V4><<<gl, tb, 0, c->stream>>>(src[k], stage, nl);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(dst[k], stage, (size_t)nl * 3 * sizeof(REAL), cudaMemcpyDeviceToHost, c->stream));
        c->launches++;
    }
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// peak probe kernels (a measured non-tensor FP32/FP64 peak for the roofline)
// ---------------------------------------------------------------------------------------------
constexpr int kProbeIters = 4096;
constexpr int kProbeChains = 12;   // independent chains per thread (latency 4, issue every 1-2 cycles)

__global__ void probe_ffma(float* out, float a, float b, unsigned long long* clk) {
    float x[kProbeChains];
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) x[c] = (float)(threadIdx.x + c);
    unsigned long long t0 = clock64(), g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    for (int i = 0; i < kProbeIters; ++i) {
#pragma unroll
        for (int c = 0; c < kProbeChains; ++c) x[c] = fmaf(x[c], a, a);   // 2 distinct registers: not operand-fetch bound
    }
    unsigned long long t1 = clock64(), g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    float s = 0;
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = g1 - g0; }
}
__global__ void probe_ffma2(float* out, float a, float b) {
    float2 x[kProbeChains];
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) x[c] = make_float2((float)(threadIdx.x + c), (float)c);
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
    for (int i = 0; i < kProbeIters; ++i) {
#pragma unroll
        for (int c = 0; c < kProbeChains; ++c) x[c] = __ffma2_rn(x[c], a2, a2);   // 2 distinct register pairs (3 pairs would cost 3 cycles)
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) s += x[c].x + x[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void probe_dfma(double* out, double a, double b) {
    double x[kProbeChains];
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) x[c] = (double)(threadIdx.x + c);
    for (int i = 0; i < kProbeIters; ++i) {
#pragma unroll
        for (int c = 0; c < kProbeChains; ++c) x[c] = fma(x[c], a, a);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void probe_mufu(float* out) {
    float x[kProbeChains];
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) x[c] = 1.5f + (float)(threadIdx.x + c);
    for (int i = 0; i < kProbeIters; ++i) {
#pragma unroll
        for (int c = 0; c < kProbeChains; ++c) x[c] = rsqrt_approx(x[c]);
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int gravb200_abi_version(void) { return 1; }

const char* gravb200_last_error(void) { return g_err; }

int gravb200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        fail(GRAVB200_ENODEV, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return 0;
    }
    return n;
}

int gravb200_sym_variant_count(int dtype) {
    if (dtype != GRAVB200_F32 && dtype != GRAVB200_F64) return 0;
    return (int)variants_sym_of(dtype).size();
}

int gravb200_small_variant_count(void) { return (int)variants_small().size(); }

int gravb200_small_geometry(int64_t n_total, int dtype, int sm_count, int variant, int64_t* out, int n) {
    if (!out || n < 6) return fail(GRAVB200_EINVAL, "out needs room for 6 values");
    if (dtype != GRAVB200_F32 && dtype != GRAVB200_F64) return fail(GRAVB200_EINVAL, "unknown dtype %d", dtype);
    if (variant < kSmallBase || variant - kSmallBase >= (int)variants_small().size()) return fail(GRAVB200_EINVAL, "variant %d is not a persistent small-N kernel", variant);
    if (n_total < 1 || sm_count < 1) return fail(GRAVB200_EINVAL, "n_total and sm_count must be >= 1");
    SmallGeometry g;
    memset(&g, 0, sizeof(g));
    const int fit = small_geometry(n_total, dtype, sm_count, variants_small()[variant - kSmallBase], &g);
    if (fit == 1) return fail(GRAVB200_EINVAL, "%lld bodies need more rows per CTA than the variant has lanes for", (long long)n_total);
    const int64_t t[6] = {g.grid, g.rpc, g.ng, g.slice, g.nsl, (int64_t)g.smem};
    memcpy(out, t, sizeof(t));
    return fit == 2 ? 1 : 0;
}

int gravb200_variant_count(int dtype) {
    if (dtype != GRAVB200_F32 && dtype != GRAVB200_F64) return 0;
    return (int)variants_of(dtype).size();
}
const char* gravb200_variant_name(int dtype, int variant) {
    if (dtype != GRAVB200_F32 && dtype != GRAVB200_F64) return "";
    if (variant >= kSmallBase) return variant - kSmallBase < (int)variants_small().size() ? variants_small()[variant - kSmallBase].name : "";
    if (variant >= kSymBase && variant - kSymBase < (int)variants_sym_of(dtype).size())
        return variants_sym_of(dtype)[variant - kSymBase].name;
    const auto& vs = variants_of(dtype);
    if (variant < 0 || variant >= (int)vs.size()) return "";
    return vs[variant].name;
}

int gravb200_nccl_unique_id(void* id) {
    if (!id) return fail(GRAVB200_EINVAL, "id is NULL");
    int rc = nccl_load();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == GRAVB200_NCCL_ID_BYTES, "NCCL id size");
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return 0;
}

int gravb200_ctx_create(int64_t n_total, int dtype, int device, int rank, int world, const void* nccl_id,
                        gravb200_ctx** out) {
    if (!out) return fail(GRAVB200_EINVAL, "out is NULL");
    *out = nullptr;
    if (n_total < 1) return fail(GRAVB200_EINVAL, "n_total must be >= 1 (got %lld)", (long long)n_total);
    if (dtype != GRAVB200_F32 && dtype != GRAVB200_F64) return fail(GRAVB200_EINVAL, "unknown dtype %d", dtype);
    if (world < 1 || rank < 0 || rank >= world) return fail(GRAVB200_EINVAL, "bad rank/world %d/%d", rank, world);
    if (world > 1 && !nccl_id) return fail(GRAVB200_EINVAL, "world > 1 needs an NCCL unique id");
    if (world > kMaxPeers + 1) return fail(GRAVB200_EINVAL, "world %d exceeds the supported %d GPUs", world, kMaxPeers + 1);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(GRAVB200_ENODEV, "no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail(GRAVB200_EINVAL, "device %d out of range [0,%d)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(GRAVB200_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);

    gravb200_ctx* c = new (std::nothrow) gravb200_ctx();
    if (!c) return fail(GRAVB200_EINVAL, "out of host memory");
    c->dtype = dtype;
    c->esz = dtype == GRAVB200_F32 ? 4 : 8;
    c->device = device;
    c->rank = rank;
    c->world = world;
    if (const char* e = getenv("GRAVB200_PDL")) c->pdl = e[0] == '1';
    if (const char* e = getenv("GRAVB200_SPLIT_WEIGHTED")) c->split_weighted = e[0] != '0';
    if (const char* e = getenv("GRAVB200_SPLIT")) c->split_mode = e[0] == '0' ? 0 : (e[0] == '1' ? 1 : -1);   // A/B runs of unmodified callers
    c->n_total = n_total;
    c->chunk = shard_chunk(n_total, world, dtype);
    c->n_pad = c->chunk * world;
    c->row0 = std::min<int64_t>((int64_t)rank * c->chunk, n_total);
    c->n_local = std::max<int64_t>(0, std::min<int64_t>(c->chunk, n_total - c->row0));
    c->sm_count = prop.multiProcessorCount;

    auto cleanup = [&](int rc) { gravb200_ctx_destroy(c); return rc; };
#define CUX(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return cleanup(fail(GRAVB200_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)));      \
    } while (0)
    CUX(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto& ev : c->ev) CUX(cudaEventCreate(&ev));
    if (world > 1) for (auto& ev : c->tev) CUX(cudaEventCreate(&ev));
    const size_t v4 = 4 * c->esz;
    for (int b = 0; b < 2; ++b) {
        CUX(cudaMalloc(&c->pos[b], (size_t)c->n_pad * v4));
        CUX(cudaMemsetAsync(c->pos[b], 0, (size_t)c->n_pad * v4, c->stream));
        CUX(cudaMalloc(&c->vel[b], (size_t)c->chunk * v4));
        CUX(cudaMemsetAsync(c->vel[b], 0, (size_t)c->chunk * v4, c->stream));
    }
    CUX(cudaMalloc(&c->acc, (size_t)c->chunk * v4));
    CUX(cudaMemsetAsync(c->acc, 0, (size_t)c->chunk * v4, c->stream));
    CUX(cudaMalloc(&c->stage3, (size_t)n_total * 3 * c->esz));
    CUX(cudaMalloc(&c->stagem, (size_t)n_total * c->esz));
    CUX(cudaMalloc(&c->clk, 4096 * sizeof(unsigned long long)));   // [0..1] CTA 0 cycles/ns, then per-CTA start/end stamps
    CUX(cudaMalloc(&c->flags, kFlagsTotal * sizeof(unsigned long long)));
    CUX(cudaMemsetAsync(c->flags, 0, kFlagsTotal * sizeof(unsigned long long), c->stream));
    CUX(cudaMalloc(&c->done_ctr, 2 * sizeof(unsigned int)));
    CUX(cudaMemsetAsync(c->done_ctr, 0, 2 * sizeof(unsigned int), c->stream));
    CUX(cudaMalloc(&c->tstart, sizeof(unsigned long long)));
    CUX(cudaMemsetAsync(c->tstart, 0xff, sizeof(unsigned long long), c->stream));
    if (world > 1) {   // symmetric sweep accumulator: must exist before peer_export
        CUX(cudaMalloc(&c->acc64, (size_t)c->n_pad * 4 * sizeof(double)));
        CUX(cudaMemsetAsync(c->acc64, 0, (size_t)c->n_pad * 4 * sizeof(double), c->stream));
    }
    CUX(cudaMalloc(&c->xerr, sizeof(int)));
    CUX(cudaMemsetAsync(c->xerr, 0, sizeof(int), c->stream));
    CUX(cudaMemsetAsync(c->clk, 0, 4096 * sizeof(unsigned long long), c->stream));
#undef CUX
    if (world > 1) {
        int rc = nccl_load();
        if (rc) return cleanup(rc);
        ncclUniqueId u;
        memcpy(&u, nccl_id, sizeof(u));
        ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, u, rank);
        if (r != ncclSuccess)
            return cleanup(fail(GRAVB200_ENCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)));
    }
    int rc = pick_variant(c);
    if (rc) return cleanup(rc);
    cudaError_t es = cudaStreamSynchronize(c->stream);
    if (es != cudaSuccess) return cleanup(fail(GRAVB200_ECUDA, "init sync: %s", cudaGetErrorString(es)));
    *out = c;
    return 0;
}

int gravb200_ctx_destroy(gravb200_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    graph_invalidate(c);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (int b = 0; b < 2; ++b) {
        if (c->pos[b]) cudaFree(c->pos[b]);
        if (c->vel[b]) cudaFree(c->vel[b]);
    }
    if (c->acc) cudaFree(c->acc);
    if (c->partial) cudaFree(c->partial);
    if (c->counters) cudaFree(c->counters);
    if (c->stage3) cudaFree(c->stage3);
    if (c->stagem) cudaFree(c->stagem);
    if (c->clk) cudaFree(c->clk);
    for (int q = 0; q <= kMaxPeers; ++q) {
        if (!c->peer_is_ipc[q]) continue;
        for (int b = 0; b < 2; ++b)
            if (c->peer_pos[b][q]) cudaIpcCloseMemHandle(c->peer_pos[b][q]);
        if (c->peer_flags[q]) cudaIpcCloseMemHandle(c->peer_flags[q]);
        if (c->peer_acc[q]) cudaIpcCloseMemHandle(c->peer_acc[q]);
    }
    if (c->flags) cudaFree(c->flags);
    if (c->done_ctr) cudaFree(c->done_ctr);
    if (c->tstart) cudaFree(c->tstart);
    if (c->acc64) cudaFree(c->acc64);
    if (c->row_start) cudaFree(c->row_start);
    if (c->xerr) cudaFree(c->xerr);
    if (c->gbar) cudaFree(c->gbar);
    for (auto& ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->tev)
        if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int gravb200_upload(gravb200_ctx* c, const void* r, const void* v, const void* m, double G, double T, double eps) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (!r || !v || !m) return fail(GRAVB200_EINVAL, "r, v and m are required");
    if (eps < 0) return fail(GRAVB200_EINVAL, "eps must be >= 0");
    CU(cudaSetDevice(c->device));
    if (c->G != G || c->T != T || c->eps != eps) graph_invalidate(c);
    c->G = G; c->T = T; c->eps = eps;
    int rc0 = drop_pending(c);
    if (rc0) return rc0;
    int rc = c->dtype == GRAVB200_F32 ? upload_impl<float, float4>(c, r, v, m) : upload_impl<double, double4>(c, r, v, m);
    if (rc) return rc;
    // masses travel in .w of both buffers: copy front -> back once so the epilogue's w is consistent
    CU(cudaMemcpyAsync(c->pos[c->front ^ 1], c->pos[c->front], (size_t)c->n_pad * 4 * c->esz,
                       cudaMemcpyDeviceToDevice, c->stream));
    // peer-store mode: nobody may start writing r' into this GPU's back buffer before the copy above
    // is done everywhere (every rank calls upload, so the barrier generations stay in lockstep)
    // The barrier is only ENQUEUED (the first sweep is stream-ordered behind it): a host thread that
    // drives several shards uploads them one after the other and must not block here.
    CU(cudaStreamSynchronize(c->stream));
    if (c->peer_mode) { rc = peer_barrier(c); if (rc) return rc; }
    c->uploaded = true;
    return 0;
}

int gravb200_upload_positions(gravb200_ctx* c, const void* r) {
    if (!c || !r) return fail(GRAVB200_EINVAL, "ctx / r is NULL");
    if (!c->uploaded) return fail(GRAVB200_EINVAL, "gravb200_upload must come first");
    CU(cudaSetDevice(c->device));
    int rc = drop_pending(c);
    if (rc) return rc;
    rc = c->dtype == GRAVB200_F32 ? upload_impl<float, float4>(c, r, nullptr, nullptr)
                                      : upload_impl<double, double4>(c, r, nullptr, nullptr);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int gravb200_upload_rows(gravb200_ctx* c, const void* r_own, const void* v_own) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (!c->uploaded) return fail(GRAVB200_EINVAL, "gravb200_upload must come first (masses, G, T)");
    if (!r_own && c->n_local > 0) return fail(GRAVB200_EINVAL, "r_own is NULL");
    CU(cudaSetDevice(c->device));
    int rc = drop_pending(c);
    if (rc) return rc;
    return c->dtype == GRAVB200_F32 ? upload_rows_impl<float, float4>(c, r_own, v_own) : upload_rows_impl<double, double4>(c, r_own, v_own);
}

int gravb200_download_rows(gravb200_ctx* c, void* r_own, void* v_own, void* a_own) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (!c->uploaded) return fail(GRAVB200_EINVAL, "no state uploaded");
    CU(cudaSetDevice(c->device));
    return c->dtype == GRAVB200_F32 ? download_rows_impl<float, float4>(c, r_own, v_own, a_own)
                                    : download_rows_impl<double, double4>(c, r_own, v_own, a_own);
}

int gravb200_stage1(gravb200_ctx* c) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (!c->uploaded) return fail(GRAVB200_EINVAL, "no state uploaded");
    CU(cudaSetDevice(c->device));
    // a second stage1 without a stage2 in between recomputes the same step: whatever the first one left in
    // the multi-shard accumulator must not be added twice
    if (c->pending) { int rc0 = drop_pending(c); if (rc0) return rc0; }
    CU(cudaEventRecord(c->ev[0], c->stream));
    int rc = launch_sweep(c, 1);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[1], c->stream));
    c->ev_sweep = true;
    c->pending = true;
    c->exchanged = false;
    return 0;
}

int gravb200_exchange(gravb200_ctx* c) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (!c->pending) return fail(GRAVB200_EINVAL, "exchange without a preceding stage1");
    if (c->exchanged || c->world == 1) return 0;
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->ev[2], c->stream));
    int rc = exchange(c);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[3], c->stream));
    c->ev_xchg = true;
    c->exchanged = true;
    return 0;
}

int gravb200_peer_barrier(gravb200_ctx* c) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (c->world == 1 || !c->peer_mode) return 0;
    if (c->pending) return fail(GRAVB200_EINVAL, "peer barrier between stage1 and stage2");
    CU(cudaSetDevice(c->device));
    return peer_barrier(c);
}

int gravb200_group_begin(void) {
    int rc = nccl_load();
    if (rc) return rc;
    NC(g_nccl.GroupStart());
    return 0;
}
int gravb200_group_end(void) {
    int rc = nccl_load();
    if (rc) return rc;
    NC(g_nccl.GroupEnd());
    return 0;
}

int gravb200_stage2(gravb200_ctx* c) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (!c->pending) return fail(GRAVB200_EINVAL, "stage2 without a preceding stage1");
    int rc = gravb200_exchange(c);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->front ^= 1;
    c->pending = false;
    return check_barrier_error(c);
}

int gravb200_steps(gravb200_ctx* c, int k) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (!c->uploaded) return fail(GRAVB200_EINVAL, "no state uploaded");
    if (k < 0) return fail(GRAVB200_EINVAL, "k < 0");
    CU(cudaSetDevice(c->device));
    int rc0 = drop_pending(c);
    if (rc0) return rc0;
    CU(cudaEventRecord(c->ev[0], c->stream));
    int s = 0;
    if (c->use_small && c->n_local > 0) {
        // all k steps in ONE cooperative launch (grid barrier between steps), in chunks that keep a launch finite
        for (; s < k;) {
            const int kk = std::min(k - s, 1 << 16);
            int rc = launch_small(c, kk, 1);
            if (rc) return rc;
            if (kk & 1) c->front ^= 1;
            s += kk;
        }
    } else if (c->world == 1 && c->n_total <= kGraphMaxN && c->n_local > 0) {
        // launch-bound sizes: whole groups of kGraphSteps steps go out as one graph launch each
        for (; k - s >= kGraphSteps; s += kGraphSteps) {
            cudaGraphExec_t ex = nullptr;
            int rc = step_graph_for(c, &ex);
            if (rc) return rc;
            CU(cudaGraphLaunch(ex, c->stream));
            c->launches += c->step_graph_kernels;
        }
    }
    for (; s < k; ++s) {
        int rc = launch_sweep(c, 1);
        if (rc) return rc;
        rc = exchange(c, s == k - 1);   // symmetric sweep on several shards: the next sweep waits for the peers itself
        if (rc) return rc;
        c->front ^= 1;
    }
    CU(cudaEventRecord(c->ev[1], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->ev_sweep = false;
    c->ev_steps = true;
    if (c->use_small && k > 1) {
        int e = 0;
        CU(cudaMemcpy(&e, c->xerr, sizeof(int), cudaMemcpyDeviceToHost));
        if (e) return fail(GRAVB200_ECUDA, "persistent step kernel: a CTA did not reach the grid barrier within 20 s");
    }
    return check_barrier_error(c);
}

int gravb200_sync(gravb200_ctx* c) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int gravb200_download(gravb200_ctx* c, void* r, void* v, void* a) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (!c->uploaded) return fail(GRAVB200_EINVAL, "no state uploaded");
    CU(cudaSetDevice(c->device));
    return c->dtype == GRAVB200_F32 ? download_impl<float, float4>(c, r, v, a) : download_impl<double, double4>(c, r, v, a);
}

int gravb200_partition(int64_t n_total, int dtype, int world, int rank, int64_t* row0, int64_t* n_local) {
    if (n_total < 0 || world < 1 || rank < 0 || rank >= world) return fail(GRAVB200_EINVAL, "bad partition arguments");
    if (dtype != GRAVB200_F32 && dtype != GRAVB200_F64) return fail(GRAVB200_EINVAL, "dtype must be GRAVB200_F32 or GRAVB200_F64");
    const int64_t chunk = shard_chunk(n_total, world, dtype);
    const int64_t r0 = std::min<int64_t>((int64_t)rank * chunk, n_total);
    if (row0) *row0 = r0;
    if (n_local) *n_local = std::max<int64_t>(0, std::min<int64_t>(chunk, n_total - r0));
    return 0;
}

int gravb200_shard(const gravb200_ctx* c, int64_t* row0, int64_t* n_local) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (row0) *row0 = c->row0;
    if (n_local) *n_local = c->n_local;
    return 0;
}

int gravb200_timings(gravb200_ctx* c, float* ms, int n) {
    if (!c || !ms) return fail(GRAVB200_EINVAL, "ctx / ms is NULL");
    CU(cudaSetDevice(c->device));
    for (int i = 0; i < n; ++i) ms[i] = -1.f;
    if (n > 0 && c->ev_sweep) CU(cudaEventElapsedTime(&ms[0], c->ev[0], c->ev[1]));
    if (n > 1 && c->ev_xchg) CU(cudaEventElapsedTime(&ms[1], c->ev[2], c->ev[3]));
    if (n > 2 && c->ev_steps) CU(cudaEventElapsedTime(&ms[2], c->ev[0], c->ev[1]));
    if (n > 3) {
        unsigned long long h[2] = {0, 0};
        CU(cudaMemcpy(h, c->clk, sizeof(h), cudaMemcpyDeviceToHost));
        ms[3] = h[1] ? (float)((double)h[0] / (double)h[1] * 1e3) : -1.f;   // cycles/ns -> MHz
        if (n > 4) ms[4] = (float)((double)h[1] * 1e-6);   // lifetime of CTA 0 of the last sweep, ms
        if (n > 7 && c->tev_ok && c->use_sym && c->world > 1 && cudaEventQuery(c->tev[3]) == cudaSuccess)
            for (int i = 0; i < 3; ++i) CU(cudaEventElapsedTime(&ms[5 + i], c->tev[i], c->tev[i + 1]));
        if (n > 9 && c->use_sym && c->world > 1) {   // the integrate kernel's own work, after its wait for the peers' sweeps
            unsigned long long st[2] = {0, 0};
            CU(cudaMemcpy(st, c->clk + 2048, sizeof(st), cudaMemcpyDeviceToHost));
            if (st[0] && st[1] > st[0]) ms[9] = (float)((double)(st[1] - st[0]) * 1e-6);
            CU(cudaMemset(c->clk + 2048, 0, sizeof(st)));
        }
        if (n > 8 && c->use_sym && c->world > 1 && c->sym_balance) {   // this shard's share of the last sweep against the equal share
            unsigned long long items = 0;
            CU(cudaMemcpy(&items, c->flags + kFlagsBalance + 2 * (kMaxPeers + 1) + c->rank, sizeof(items), cudaMemcpyDeviceToHost));
            if (items && c->sym_total > 0) ms[8] = (float)((double)items * c->world / (double)c->sym_total);
        }
#ifdef SYM_DEBUG
        if (n > 16) {   // [2036..2046] divergence counters of SYM_DIVCHK, [2047] cycles spent in jbar waits
            unsigned long long d[12];
            CU(cudaMemcpy(d, c->clk + 2036, sizeof(d), cudaMemcpyDeviceToHost));
            ms[10] = (float)d[10]; ms[11] = (float)((double)d[11] * 1e-6);
            for (int i = 0; i < 10 && 12 + i < n; ++i) ms[12 + i] = (float)d[i];
            CU(cudaMemset(c->clk + 2036, 0, sizeof(d)));
        }
#endif
    }
    return 0;
}

int gravb200_info(const gravb200_ctx* c, int64_t* info, int n) {
    if (!c || !info) return fail(GRAVB200_EINVAL, "ctx / info is NULL");
    int64_t vals[13];
    vals[12] = c->use_sym && c->sym_split ? 1 : 0;
    if (c->use_small) {
        const SmallVariant& v = variants_small()[c->small_variant];
        const int64_t t[12] = {c->grid, v.threads, v.r, c->small_slice, 1, (int64_t)c->small_smem,
                               c->launches, c->sm_count, 1, c->occ, 0, kSmallBase + c->small_variant};
        memcpy(vals, t, sizeof(t));
    } else if (c->use_sym) {
        const SymVariant& v = variants_sym_of(c->dtype)[c->sym_variant];
        const int64_t t[12] = {c->grid, v.threads, v.r, v.tile, v.stages, (int64_t)v.smem,
                               c->launches, c->sm_count, 1, c->occ, c->peer_mode ? 1 : 0, kSymBase + c->sym_variant};
        memcpy(vals, t, sizeof(t));
    } else {
        const Variant& v = variants_of(c->dtype)[c->variant];
        const int64_t t[12] = {c->grid, v.threads, v.r, v.tile, v.stages, (int64_t)v.smem,
                               c->launches, c->sm_count, v.pack, c->occ, c->peer_mode ? 1 : 0, c->variant};
        memcpy(vals, t, sizeof(t));
    }
    for (int i = 0; i < n && i < 13; ++i) info[i] = vals[i];
    return 0;
}

int gravb200_set_variant(gravb200_ctx* c, int variant) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (variant >= kSmallBase) {
        if (variant - kSmallBase >= (int)variants_small().size()) return fail(GRAVB200_EINVAL, "variant %d out of range", variant);
    } else if (variant >= kSymBase) {
        if (variant - kSymBase >= (int)variants_sym_of(c->dtype).size()) return fail(GRAVB200_EINVAL, "variant %d out of range", variant);
    } else if (variant >= (int)variants_of(c->dtype).size()) return fail(GRAVB200_EINVAL, "variant %d out of range", variant);
    if (c->pending) return fail(GRAVB200_EINVAL, "cannot switch variant between stage1 and stage2");
    CU(cudaSetDevice(c->device));
    const int before = c->forced_variant;
    c->forced_variant = variant < 0 ? -1 : variant;
    int rc = pick_variant(c);
    if (rc) {   // e.g. a persistent variant the universe does not fit: the context keeps working with what it had
        c->forced_variant = before;
        pick_variant(c);
        return rc;
    }
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int gravb200_set_split(gravb200_ctx* c, int mode) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (mode < -1 || mode > 1) return fail(GRAVB200_EINVAL, "split mode %d: -1 (automatic), 0 or 1", mode);
    if (c->pending) return fail(GRAVB200_EINVAL, "cannot switch the split mode between stage1 and stage2");
    CU(cudaSetDevice(c->device));
    c->split_mode = mode;
    int rc = pick_variant(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

void* gravb200_device_ptr(gravb200_ctx* c, int which) {
    if (!c) return nullptr;
    switch (which) {
        case 0: return c->pos[c->front];
        case 1: return c->pos[c->front ^ 1];
        case 2: return c->vel[c->front];
        case 3: return c->acc;
        case 4: return c->clk;
        default: return nullptr;
    }
}

namespace {
struct PeerBlob {   // what one shard publishes to the others (GRAVB200_PEER_BLOB_BYTES)
    int32_t pid, device, rank, world;
    uint64_t pos[2], flags, acc64;   // raw device pointers, meaningful inside the owner's process
    cudaIpcMemHandle_t h_pos[2], h_flags, h_acc64;
};
static_assert(sizeof(PeerBlob) <= GRAVB200_PEER_BLOB_BYTES, "peer blob size");
}  // namespace

int gravb200_peer_export(gravb200_ctx* c, void* blob) {
    if (!c || !blob) return fail(GRAVB200_EINVAL, "ctx / blob is NULL");
    CU(cudaSetDevice(c->device));
    PeerBlob b;
    memset(&b, 0, sizeof(b));
    b.pid = (int32_t)getpid();
    b.device = c->device;
    b.rank = c->rank;
    b.world = c->world;
    b.pos[0] = (uint64_t)c->pos[0];
    b.pos[1] = (uint64_t)c->pos[1];
    b.flags = (uint64_t)c->flags;
    b.acc64 = (uint64_t)c->acc64;
    if (c->acc64) CU(cudaIpcGetMemHandle(&b.h_acc64, c->acc64));
    CU(cudaIpcGetMemHandle(&b.h_pos[0], c->pos[0]));
    CU(cudaIpcGetMemHandle(&b.h_pos[1], c->pos[1]));
    CU(cudaIpcGetMemHandle(&b.h_flags, c->flags));
    memset(blob, 0, GRAVB200_PEER_BLOB_BYTES);
    memcpy(blob, &b, sizeof(b));
    return 0;
}

int gravb200_peer_connect(gravb200_ctx* c, const void* blobs) {
    if (!c || !blobs) return fail(GRAVB200_EINVAL, "ctx / blobs is NULL");
    if (c->world == 1) return 0;
    if (c->pending) return fail(GRAVB200_EINVAL, "cannot connect peers between stage1 and stage2");
    CU(cudaSetDevice(c->device));
    const int32_t mypid = (int32_t)getpid();
    for (int q = 0; q < c->world; ++q) {
        PeerBlob b;
        memcpy(&b, (const char*)blobs + (size_t)q * GRAVB200_PEER_BLOB_BYTES, sizeof(b));
        if (b.rank != q || b.world != c->world) return fail(GRAVB200_EINVAL, "peer blob %d is from rank %d of %d", q, b.rank, b.world);
        if (q == c->rank) {
            c->peer_pos[0][q] = c->pos[0]; c->peer_pos[1][q] = c->pos[1]; c->peer_flags[q] = c->flags;
            c->peer_acc[q] = c->acc64;
            continue;
        }
        if (b.pid == mypid) {
            // same process: plain peer access to the other context's allocations
            if (b.device != c->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, c->device, b.device));
                if (!can) return fail(GRAVB200_ECUDA, "GPU %d cannot access GPU %d", c->device, b.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(GRAVB200_ECUDA, "cudaDeviceEnablePeerAccess(%d): %s", b.device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            c->peer_pos[0][q] = (void*)b.pos[0]; c->peer_pos[1][q] = (void*)b.pos[1];
            c->peer_flags[q] = (unsigned long long*)b.flags;
            c->peer_acc[q] = (double*)b.acc64;
            c->peer_is_ipc[q] = false;
        } else {
            void* p0 = nullptr; void* p1 = nullptr; void* pf = nullptr;
            CU(cudaIpcOpenMemHandle(&p0, b.h_pos[0], cudaIpcMemLazyEnablePeerAccess));
            CU(cudaIpcOpenMemHandle(&p1, b.h_pos[1], cudaIpcMemLazyEnablePeerAccess));
            CU(cudaIpcOpenMemHandle(&pf, b.h_flags, cudaIpcMemLazyEnablePeerAccess));
            c->peer_pos[0][q] = p0; c->peer_pos[1][q] = p1; c->peer_flags[q] = (unsigned long long*)pf;
            if (b.acc64) {
                void* pa = nullptr;
                CU(cudaIpcOpenMemHandle(&pa, b.h_acc64, cudaIpcMemLazyEnablePeerAccess));
                c->peer_acc[q] = (double*)pa;
            }
            c->peer_is_ipc[q] = true;
        }
    }
    c->peer_connected = true;
    return 0;
}

int gravb200_set_exchange_mode(gravb200_ctx* c, int mode) {
    if (!c) return fail(GRAVB200_EINVAL, "ctx is NULL");
    if (mode != GRAVB200_XCHG_NCCL && mode != GRAVB200_XCHG_PEER) return fail(GRAVB200_EINVAL, "unknown exchange mode %d", mode);
    if (c->pending) return fail(GRAVB200_EINVAL, "cannot switch the exchange between stage1 and stage2");
    if (mode == GRAVB200_XCHG_PEER && c->world > 1 && !c->peer_connected)
        return fail(GRAVB200_EINVAL, "peer-store exchange needs gravb200_peer_connect first");
    c->peer_mode = (mode == GRAVB200_XCHG_PEER) && c->world > 1;
    CU(cudaSetDevice(c->device));
    int rc = pick_variant(c);   // the symmetric sweep on several shards depends on the exchange mode
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int gravb200_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(GRAVB200_EINVAL, "out is NULL");
    *out = nullptr;
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return 0;
}
int gravb200_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return 0;
}

int gravb200_peak_probe(int device, double* out, int n) {
    if (!out || n < 5) return fail(GRAVB200_EINVAL, "out needs room for 5 doubles");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(GRAVB200_ENODEV, "no CUDA device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    void* buf = nullptr;
    unsigned long long* clk = nullptr;
    CU(cudaMalloc(&buf, (size_t)threads * blocks * sizeof(double)));
    CU(cudaMalloc(&clk, 2 * sizeof(unsigned long long)));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    const double nthreads = (double)threads * blocks;
    const double ops = nthreads * kProbeIters * kProbeChains;
    float ms = 0;
    auto best = [&](auto&& launch) -> double {
        double bestms = 1e30;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(a);
            launch();
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            if (rep > 0 && ms < bestms) bestms = ms;
        }
        return bestms;
    };
    double t = best([&] { probe_ffma<<<blocks, threads>>>((float*)buf, 1.0001f, 0.5f, clk); });
    out[0] = ops * 2 / (t * 1e-3) / 1e12;
    unsigned long long hclk[2] = {0, 0};
    CU(cudaMemcpy(hclk, clk, sizeof(hclk), cudaMemcpyDeviceToHost));
    out[4] = hclk[1] ? (double)hclk[0] / (double)hclk[1] * 1e3 : 0.0;   // cycles per ns -> MHz
    t = best([&] { probe_ffma2<<<blocks, threads>>>((float*)buf, 1.0001f, 0.5f); });
    out[1] = ops * 4 / (t * 1e-3) / 1e12;
    t = best([&] { probe_dfma<<<blocks, threads>>>((double*)buf, 1.0001, 0.5); });
    out[2] = ops * 2 / (t * 1e-3) / 1e12;
    t = best([&] { probe_mufu<<<blocks, threads>>>((float*)buf); });
    out[3] = ops / (t * 1e-3) / 1e9;
    CU(cudaGetLastError());
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(buf);
    cudaFree(clk);
    return 0;
}

}  // extern "C"
