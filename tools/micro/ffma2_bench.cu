// Microbenchmark: issue rate of packed FP32 (FFMA2 / fma.rn.f32x2, sm_100) against scalar FFMA, alone and mixed with
// integer work -- input to the decision whether a 2-pixels-per-lane compositor would save issue slots.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float s) {
    float a[8];
    unsigned long long p[8];
    int q[4];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); }
    for (int i = 0; i < 4; i++) q[i] = threadIdx.x + i;
    const unsigned long long s2 = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 3) {  // 8 scalar FFMA
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma1(a[i], s, s);
        }
        if (MODE == 1 || MODE == 4) {  // 4 FFMA2 (same flops as MODE 0)
#pragma unroll
            for (int i = 0; i < 4; i++) p[i] = fma2(p[i], s2, s2);
        }
        if (MODE == 2) {  // 8 FFMA2 (twice the flops)
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(p[i], s2, s2);
        }
        if (MODE == 3 || MODE == 4) {  // + 8 integer ops
#pragma unroll
            for (int i = 0; i < 4; i++) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(q[i]) : "r"(it)); asm volatile("shf.l.wrap.b32 %0, %0, %0, 3;" : "+r"(q[i])); }
        }
    }
    float r = 0;
    for (int i = 0; i < 8; i++) r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    for (int i = 0; i < 4; i++) r += q[i];
    if (r == 12345.678f) out[0] = r;
}

template <int MODE>
float run(float *out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, iters, 0.999f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 0.999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float *out;
    cudaMalloc(&out, 4);
    const int iters = 20000;
    const double warps = 148.0 * 8 * 8;
    const char *names[5] = {"8 FFMA", "4 FFMA2", "8 FFMA2", "8 FFMA + 8 INT", "4 FFMA2 + 8 INT"};
    float ms[5] = {run<0>(out, iters), run<1>(out, iters), run<2>(out, iters), run<3>(out, iters), run<4>(out, iters)};
    const int fp_instr[5] = {8, 4, 8, 8, 4}, all_instr[5] = {8, 4, 8, 16, 12};
    const int flops[5] = {16, 16, 32, 16, 16};
    for (int m = 0; m < 5; m++) {
        const double cyc = ms[m] * 1e-3 * 1.965e9;  // nominal boost clock
        printf("%-18s %8.3f ms  %.2f warp-instr/clk/SM (fp only %.2f)  %.1f TFLOP/s\n", names[m], ms[m], warps * iters * all_instr[m] / cyc / 148.0,
               warps * iters * fp_instr[m] / cyc / 148.0, warps * 32 * iters * flops[m] / (ms[m] * 1e-3) / 1e12);
    }
    return 0;
}
