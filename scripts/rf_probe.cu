// Dev microbenchmark (not product): register-file read bandwidth limits of packed FFMA2 on sm_100a.
// Each kernel runs a loop of 12 independent packed ops with a different operand pattern and prints SM
// cycles per op per warp scheduler.  2.0 = FMA-pipe bound; > 2 = operand fetch bound.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 8192
#define NC 12
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float2* in, unsigned long long* cyc) {
    float2 x[NC], y[NC], z[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { x[c] = in[threadIdx.x + c]; y[c] = in[threadIdx.x + 64 + c]; z[c] = in[threadIdx.x + 128 + c]; }
    const float2 s = in[threadIdx.x + 200];
    const float sc = in[threadIdx.x + 201].x;
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (MODE == 0) x[c] = __ffma2_rn(y[c], z[c], x[c]);            // 3 distinct pairs
            if (MODE == 1) x[c] = __ffma2_rn(y[c], y[c], x[c]);            // 2 distinct pairs
            if (MODE == 2) x[c] = __ffma2_rn(y[c], s, x[c]);               // 2 distinct pairs + 1 shared pair
            if (MODE == 3) x[c] = __fmul2_rn(y[c], z[c]);                  // 2 source pairs (independent of x)
            if (MODE == 4) x[c] = __ffma2_rn(y[c], make_float2(sc, sc), x[c]);   // pair, scalar bcast, pair
            if (MODE == 5) x[c] = __fadd2_rn(make_float2(sc, sc), y[c]);   // scalar bcast + pair
            if (MODE == 6) x[c] = __ffma2_rn(x[c], s, s);                  // 1 distinct pair + shared
            if (MODE == 7) x[c] = __ffma2_rn(y[c], z[(c + 1) % NC], x[c]); // 3 distinct pairs, other mix
        }
        if (MODE == 3 || MODE == 5) { y[0].x += 1.0f; }   // keep the loop from being hoisted
    }
    unsigned long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) acc += x[c].x + x[c].y + y[c].x;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, float2* in, unsigned long long* cyc, int sms) {
    for (int threads : {128, 256}) {
        k<MODE><<<sms, threads>>>(out, in, cyc);
        k<MODE><<<sms, threads>>>(out, in, cyc);
        cudaDeviceSynchronize();
        unsigned long long h = 0;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("{\"probe\":\"%s\",\"threads\":%d,\"cycles_per_op_per_smsp\":%.3f}\n", name, threads,
               (double)h / ITERS / NC / (threads / 128.0));
    }
}
int main() {
    float* out; float2* in; unsigned long long* cyc;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    cudaMalloc(&out, p.multiProcessorCount * 512 * 4); cudaMalloc(&cyc, 8);
    cudaMalloc(&in, 1024 * 8); cudaMemset(in, 0, 1024 * 8);
    run<0>("ffma2 x=y*z+x (3 pairs)", out, in, cyc, p.multiProcessorCount);
    run<7>("ffma2 x=y*z'+x (3 pairs)", out, in, cyc, p.multiProcessorCount);
    run<1>("ffma2 x=y*y+x (2 pairs)", out, in, cyc, p.multiProcessorCount);
    run<2>("ffma2 x=y*s+x (s shared)", out, in, cyc, p.multiProcessorCount);
    run<3>("fmul2 x=y*z", out, in, cyc, p.multiProcessorCount);
    run<4>("ffma2 x=y*bcast+x", out, in, cyc, p.multiProcessorCount);
    run<5>("fadd2 x=bcast+y", out, in, cyc, p.multiProcessorCount);
    run<6>("ffma2 x=x*s+s", out, in, cyc, p.multiProcessorCount);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
