// Dev microbenchmark (not product): cost of a SYMMETRIC inner step (Newton's third law) in the ring form.
// Each lane holds one j-body {x,y,z,m} and its packed partial acceleration; each step every lane interacts
// its R register-resident i-bodies (packed pairs) with the j-body it currently holds, updating BOTH sides,
// then the j-body and its partials rotate to the next lane (warp shuffles).  Prints SM cycles per step per
// warp scheduler; 256 unique pairs (= 512 ordered interactions) per warp per step at R = 8.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float rsq(float x){float y; asm("rsqrt.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
template <int R, int SYM>
__global__ void __launch_bounds__(256, 1) ring(const float4* __restrict__ pos, float4* out, int steps, unsigned long long* cyc) {
    constexpr int P = R / 2;
    float2 xi[P], yi[P], zi[P], mi[P], ax[P], ay[P], az[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        float4 a = pos[threadIdx.x * R + 2 * q], b = pos[threadIdx.x * R + 2 * q + 1];
        xi[q] = make_float2(-a.x, -b.x); yi[q] = make_float2(-a.y, -b.y); zi[q] = make_float2(-a.z, -b.z); mi[q] = make_float2(a.w, b.w);
        ax[q] = ay[q] = az[q] = make_float2(0.f, 0.f);
    }
    float4 bj = pos[4096 + threadIdx.x];
    float2 jx = make_float2(0.f, 0.f), jy = jx, jz = jx;
    const int src = (threadIdx.x + 1) & 31;
    unsigned long long t0 = clock64();
#pragma unroll 2
    for (int s = 0; s < steps; ++s) {
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
            ax[q] = __ffma2_rn(dx, si, ax[q]);
            ay[q] = __ffma2_rn(dy, si, ay[q]);
            az[q] = __ffma2_rn(dz, si, az[q]);
            if (SYM) {
                const float2 sj = __fmul2_rn(mi[q], ri3);
                jx = __ffma2_rn(dx, sj, jx);
                jy = __ffma2_rn(dy, sj, jy);
                jz = __ffma2_rn(dz, sj, jz);
            }
        }
        bj.x = __shfl_sync(0xffffffffu, bj.x, src); bj.y = __shfl_sync(0xffffffffu, bj.y, src);
        bj.z = __shfl_sync(0xffffffffu, bj.z, src); bj.w = __shfl_sync(0xffffffffu, bj.w, src);
        if (SYM) {
            jx.x = __shfl_sync(0xffffffffu, jx.x, src); jx.y = __shfl_sync(0xffffffffu, jx.y, src);
            jy.x = __shfl_sync(0xffffffffu, jy.x, src); jy.y = __shfl_sync(0xffffffffu, jy.y, src);
            jz.x = __shfl_sync(0xffffffffu, jz.x, src); jz.y = __shfl_sync(0xffffffffu, jz.y, src);
        }
    }
    unsigned long long t1 = clock64();
    float4 o = make_float4(jx.x + jx.y, jy.x + jy.y, jz.x + jz.y, 0.f);
#pragma unroll
    for (int q = 0; q < P; ++q) { o.x += ax[q].x + ax[q].y; o.y += ay[q].x + ay[q].y; o.z += az[q].x + az[q].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = o;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int R, int SYM>
void run(const char* name, float4* pos, float4* out, unsigned long long* cyc, int sms) {
    const int steps = 8192;
    for (int threads : {128, 256}) {
        ring<R, SYM><<<sms, threads>>>(pos, out, steps, cyc);
        ring<R, SYM><<<sms, threads>>>(pos, out, steps, cyc);
        cudaDeviceSynchronize();
        unsigned long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double per_step = (double)h / steps / (threads / 128.0);
        const double ordered = 32.0 * R * (SYM ? 2 : 1);
        printf("{\"probe\":\"%s\",\"R\":%d,\"threads\":%d,\"cycles_per_step_per_smsp\":%.1f,\"ordered_interactions_per_cycle_per_smsp\":%.3f,\"T_inter_s_at_1965MHz\":%.3f}\n",
               name, R, threads, per_step, ordered / per_step, ordered / per_step * 4 * sms * 1.965e9 / 1e12);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float4 *pos, *out; unsigned long long* cyc;
    cudaMalloc(&pos, 8192 * 16); cudaMalloc(&out, p.multiProcessorCount * 256 * 16); cudaMalloc(&cyc, 8);
    float4* h = (float4*)malloc(8192 * 16);
    for (int i = 0; i < 8192; ++i) h[i] = make_float4(i * 1.37f, i * 0.91f + 3.f, i * 2.11f - 7.f, 1.f + (i % 7));
    cudaMemcpy(pos, h, 8192 * 16, cudaMemcpyHostToDevice);
    run<8, 1>("ring symmetric", pos, out, cyc, p.multiProcessorCount);
    run<8, 0>("ring one-sided", pos, out, cyc, p.multiProcessorCount);
    run<6, 1>("ring symmetric", pos, out, cyc, p.multiProcessorCount);
    run<10, 1>("ring symmetric", pos, out, cyc, p.multiProcessorCount);
    run<12, 1>("ring symmetric", pos, out, cyc, p.multiProcessorCount);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
