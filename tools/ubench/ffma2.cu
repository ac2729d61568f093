// Micro-benchmark: issue rate of packed fp32 (FFMA2 / FMUL2 / FADD2) vs scalar FFMA on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }
template <int MODE>
__global__ void k(float *out, int iters, float s) {
    float a[8]; u64 p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; p[i] = pk(a[i], a[i] + 0.5f); }
    const u64 ps = pk(s, s), pc = pk(0.25f, 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], s, 0.25f);                                            // 8 scalar FFMA
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps), "l"(pc));   // 8 FFMA2
            if (MODE == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
            if (MODE == 4) a[i] = a[i] * s;
            if (MODE == 5) a[i] = a[i] + s;
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) r += a[i] + lo(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char *name, float *out) {
    const int iters = 20000, blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 100, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)blocks * threads / 32 * iters * 8;
    printf("%-8s %8.3f ms  %.2f warp-instr/clk/SM (at 1.965 GHz)  %.1f T elem-ops/s\n", name, ms,
           winst / (ms * 1e-3) / 148 / 1.965e9, winst * 32 * (MODE >= 1 && MODE <= 3 ? 2 : 1) / (ms * 1e-3) / 1e12);
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("FFMA", out); run<1>("FFMA2", out); run<2>("FMUL2", out); run<3>("FADD2", out); run<4>("FMUL", out); run<5>("FADD", out);
    return 0;
}
