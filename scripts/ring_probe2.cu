// Dev microbenchmark (not product): variants of the symmetric ring step (see ring_probe.cu), to find out
// what limits the inner loop of sym_sweep_kernel beyond the FMA pipe.  One launch per variant, prints SM
// cycles per ring step per warp scheduler and the projected T interactions/s.
//   MODE 0  baseline: j-body and packed partials rotate by 10 SHFL at the end of the step
//   MODE 1  j-body SHFLs issued at the TOP of the step (next body arrives while this one is in use)
//   MODE 2  j-body read from shared memory (LDS.128, rotating index, prefetched one step ahead); 6 SHFL
//   MODE 3  MODE 2 + the halves of the packed partials are summed before travelling (3 FADD + 3 SHFL)
//   MODE 4  MODE 1 with scalar FFMA for the j-side accumulation
//   MODE 5  MODE 1 with the six accumulates of a pair written zig-zag (consecutive ones share one operand);
//           ptxas reorders them anyway: no gain
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float rsq(float x){float y; asm("rsqrt.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
#define SH(v) __shfl_sync(0xffffffffu, (v), src)
template <int R, int THREADS, int MODE, int UNR>
__global__ void __launch_bounds__(THREADS, 1) ring(const float4* __restrict__ pos, float4* out, int steps, unsigned long long* cyc) {
    constexpr int P = R / 2;
    __shared__ float4 tile[THREADS];   // one 32-body chunk per warp
    float2 xi[P], yi[P], zi[P], mi[P], ax[P], ay[P], az[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        float4 a = pos[(threadIdx.x * R + 2 * q) & 4095], b = pos[(threadIdx.x * R + 2 * q + 1) & 4095];
        xi[q] = make_float2(-a.x, -b.x); yi[q] = make_float2(-a.y, -b.y); zi[q] = make_float2(-a.z, -b.z); mi[q] = make_float2(a.w, b.w);
        ax[q] = ay[q] = az[q] = make_float2(0.f, 0.f);
    }
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    float4 bj = pos[4096 + threadIdx.x];
    tile[threadIdx.x] = bj;
    __syncwarp();
    float2 jx = make_float2(0.f, 0.f), jy = jx, jz = jx;
    const int src = (lane + 1) & 31;
    float4 bn = tile[wbase + ((lane + 1) & 31)];
    unsigned long long t0 = clock64();
#pragma unroll UNR
    for (int s = 0; s < steps; ++s) {
        float4 nx;
        if (MODE == 1 || MODE == 4 || MODE == 5) { nx.x = SH(bj.x); nx.y = SH(bj.y); nx.z = SH(bj.z); nx.w = SH(bj.w); }
        if (MODE == 2 || MODE == 3) { nx = tile[wbase + ((lane + s + 2) & 31)]; }
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const float2 dx = __fadd2_rn(make_float2(bj.x, bj.x), xi[q]);
            const float2 dy = __fadd2_rn(make_float2(bj.y, bj.y), yi[q]);
            const float2 dz = __fadd2_rn(make_float2(bj.z, bj.z), zi[q]);
            float2 d2 = __fmul2_rn(dx, dx);
            d2 = __ffma2_rn(dy, dy, d2);
            d2 = __ffma2_rn(dz, dz, d2);
            const float2 ri = make_float2(rsq(d2.x), rsq(d2.y));
            const float2 ri2 = __fmul2_rn(ri, ri);
            const float2 ri3 = __fmul2_rn(ri2, ri);
            const float2 si = __fmul2_rn(make_float2(bj.w, bj.w), ri3);
            const float2 sj = __fmul2_rn(mi[q], ri3);
            if (MODE != 5) {
                ax[q] = __ffma2_rn(dx, si, ax[q]);
                ay[q] = __ffma2_rn(dy, si, ay[q]);
                az[q] = __ffma2_rn(dz, si, az[q]);
            }
            if (MODE == 5) {   // zig-zag: consecutive accumulates share one operand (dx | sj | dy | si | dz)
                ax[q] = __ffma2_rn(dx, si, ax[q]);
                jx = __ffma2_rn(dx, sj, jx);
                jy = __ffma2_rn(dy, sj, jy);
                ay[q] = __ffma2_rn(dy, si, ay[q]);
                az[q] = __ffma2_rn(dz, si, az[q]);
                jz = __ffma2_rn(dz, sj, jz);
            } else if (MODE == 4) {
                jx.x = fmaf(dx.x, sj.x, jx.x); jx.y = fmaf(dx.y, sj.y, jx.y);
                jy.x = fmaf(dy.x, sj.x, jy.x); jy.y = fmaf(dy.y, sj.y, jy.y);
                jz.x = fmaf(dz.x, sj.x, jz.x); jz.y = fmaf(dz.y, sj.y, jz.y);
            } else {
                jx = __ffma2_rn(dx, sj, jx);
                jy = __ffma2_rn(dy, sj, jy);
                jz = __ffma2_rn(dz, sj, jz);
            }
        }
        if (MODE == 0) { bj.x = SH(bj.x); bj.y = SH(bj.y); bj.z = SH(bj.z); bj.w = SH(bj.w); }
        else if (MODE == 1 || MODE == 4 || MODE == 5) bj = nx;
        else { bj = bn; bn = nx; }
        if (MODE == 3) {
            const float tx = jx.x + jx.y, ty = jy.x + jy.y, tz = jz.x + jz.y;
            jx = make_float2(SH(tx), 0.f); jy = make_float2(SH(ty), 0.f); jz = make_float2(SH(tz), 0.f);
        } else {
            jx.x = SH(jx.x); jx.y = SH(jx.y); jy.x = SH(jy.x); jy.y = SH(jy.y); jz.x = SH(jz.x); jz.y = SH(jz.y);
        }
    }
    unsigned long long t1 = clock64();
    float4 o = make_float4(jx.x + jx.y, jy.x + jy.y, jz.x + jz.y, bj.x + bn.y);
#pragma unroll
    for (int q = 0; q < P; ++q) { o.x += ax[q].x + ax[q].y; o.y += ay[q].x + ay[q].y; o.z += az[q].x + az[q].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = o;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int R, int THREADS, int MODE, int UNR>
void run(float4* pos, float4* out, unsigned long long* cyc, int sms) {
    const int steps = 8192;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, ring<R, THREADS, MODE, UNR>);
    ring<R, THREADS, MODE, UNR><<<sms, THREADS>>>(pos, out, steps, cyc);
    ring<R, THREADS, MODE, UNR><<<sms, THREADS>>>(pos, out, steps, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_step = (double)h / steps / (THREADS / 128.0);
    const double ordered = 32.0 * R * 2;
    printf("{\"probe\":\"ring2\",\"mode\":%d,\"R\":%d,\"threads\":%d,\"unroll\":%d,\"regs\":%d,\"cycles_per_step_per_smsp\":%.1f,\"fma_floor\":%d,\"T_inter_s_at_1965MHz\":%.3f,\"err\":\"%s\"}\n",
           MODE, R, THREADS, UNR, fa.numRegs, per_step, R * 16, ordered / per_step * 4 * sms * 1.965e9 / 1e12, cudaGetErrorString(e));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    float4 *pos, *out; unsigned long long* cyc;
    cudaMalloc(&pos, 8192 * 16); cudaMalloc(&out, sms * 512 * 16); cudaMalloc(&cyc, 8);
    float4* h = (float4*)malloc(8192 * 16);
    for (int i = 0; i < 8192; ++i) h[i] = make_float4(i * 1.37f, i * 0.91f + 3.f, i * 2.11f - 7.f, 1.f + (i % 7));
    cudaMemcpy(pos, h, 8192 * 16, cudaMemcpyHostToDevice);
    run<12, 256, 0, 2>(pos, out, cyc, sms);
    run<12, 256, 5, 2>(pos, out, cyc, sms);
    run<12, 256, 5, 1>(pos, out, cyc, sms);
    run<12, 256, 1, 2>(pos, out, cyc, sms);
    run<12, 256, 2, 2>(pos, out, cyc, sms);
    run<12, 256, 3, 2>(pos, out, cyc, sms);
    run<12, 256, 4, 2>(pos, out, cyc, sms);
    run<12, 256, 0, 1>(pos, out, cyc, sms);
    run<12, 256, 1, 1>(pos, out, cyc, sms);
    run<12, 256, 2, 1>(pos, out, cyc, sms);
    run<12, 256, 2, 4>(pos, out, cyc, sms);
    run<8, 256, 0, 2>(pos, out, cyc, sms);
    run<8, 256, 1, 2>(pos, out, cyc, sms);
    run<8, 256, 2, 2>(pos, out, cyc, sms);
    run<8, 384, 0, 2>(pos, out, cyc, sms);
    run<8, 384, 1, 2>(pos, out, cyc, sms);
    run<8, 384, 2, 2>(pos, out, cyc, sms);
    run<8, 384, 2, 1>(pos, out, cyc, sms);
    run<6, 384, 2, 2>(pos, out, cyc, sms);
    run<6, 512, 2, 2>(pos, out, cyc, sms);
    run<6, 512, 2, 1>(pos, out, cyc, sms);
    run<10, 256, 2, 2>(pos, out, cyc, sms);
    run<10, 384, 2, 1>(pos, out, cyc, sms);
    run<14, 256, 2, 2>(pos, out, cyc, sms);
    run<16, 256, 2, 1>(pos, out, cyc, sms);
    run<12, 384, 2, 1>(pos, out, cyc, sms);
    run<12, 384, 0, 1>(pos, out, cyc, sms);
    run<12, 384, 1, 1>(pos, out, cyc, sms);
    run<8, 512, 2, 1>(pos, out, cyc, sms);
    run<8, 512, 0, 1>(pos, out, cyc, sms);
    run<10, 384, 2, 2>(pos, out, cyc, sms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
