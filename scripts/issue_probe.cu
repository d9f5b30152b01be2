// Dev microbenchmark (not product): does a packed FFMA2 hold the SMSP dispatch port for one cycle or two?
// Runs loops of 12 independent FFMA2 with 0/2/4 independent MUFU.RSQ (+ an LDS) mixed in and prints
// SM cycles per loop trip per warp scheduler.  If 12 FFMA2 + 2 MUFU costs 24 cycles the MUFU issue
// hides in the FFMA2 shadow; 26 means the port is blocked.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 8192
template <int NMUFU, int NLDS, int PACKED>
__global__ void __launch_bounds__(512) mix(float* out, float a, float b, unsigned long long* cyc) {
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(a, b, a, b);
    __syncthreads();
    float2 x[12];
    float y[4];
#pragma unroll
    for (int c = 0; c < 12; ++c) x[c] = make_float2(threadIdx.x + c, c);
#pragma unroll
    for (int c = 0; c < 4; ++c) y[c] = 1.5f + threadIdx.x + c;
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b);
    float4 acc4 = make_float4(0, 0, 0, 0);
    unsigned long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            if (PACKED) x[c] = __ffma2_rn(x[c], a2, b2);
            else { x[c].x = fmaf(x[c].x, a, b); x[c].y = fmaf(x[c].y, a, b); }
        }
#pragma unroll
        for (int c = 0; c < NMUFU; ++c) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(y[c]));
#pragma unroll
        for (int c = 0; c < NLDS; ++c) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[(i + c) & 63]))); acc4.x += v.x; }
    }
    unsigned long long t1 = clock64();
    float s = acc4.x;
#pragma unroll
    for (int c = 0; c < 12; ++c) s += x[c].x + x[c].y;
#pragma unroll
    for (int c = 0; c < 4; ++c) s += y[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NMUFU, int NLDS, int PACKED>
void run(const char* name, float* out, unsigned long long* cyc, int sms) {
    for (int threads : {128, 256, 512}) {
        mix<NMUFU, NLDS, PACKED><<<sms, threads>>>(out, 1.0001f, 0.5f, cyc);
        mix<NMUFU, NLDS, PACKED><<<sms, threads>>>(out, 1.0001f, 0.5f, cyc);
        cudaDeviceSynchronize();
        unsigned long long h = 0;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double warps_per_smsp = threads / 32 / 4.0;
        printf("{\"probe\":\"%s\",\"threads\":%d,\"cycles_per_trip_per_smsp\":%.3f}\n", name, threads,
               (double)h / ITERS / warps_per_smsp);
    }
}
int main() {
    float* out; unsigned long long* cyc;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    cudaMalloc(&out, p.multiProcessorCount * 512 * 4); cudaMalloc(&cyc, 8);
    run<0, 0, 1>("12xFFMA2", out, cyc, p.multiProcessorCount);
    run<2, 0, 1>("12xFFMA2+2MUFU", out, cyc, p.multiProcessorCount);
    run<4, 0, 1>("12xFFMA2+4MUFU", out, cyc, p.multiProcessorCount);
    run<2, 1, 1>("12xFFMA2+2MUFU+1LDS", out, cyc, p.multiProcessorCount);
    run<0, 0, 0>("24xFFMA", out, cyc, p.multiProcessorCount);
    run<2, 0, 0>("24xFFMA+2MUFU", out, cyc, p.multiProcessorCount);
    run<4, 0, 0>("24xFFMA+4MUFU", out, cyc, p.multiProcessorCount);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
