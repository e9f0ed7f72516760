// ubench_ffma2.cu -- issue/throughput micro-benchmark for FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_ffma2 ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
    float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b) {
    unsigned d; asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ float ex2(float a) {
    float d; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d;
}
constexpr int IT = 4096;
// MODE 0: 8 FFMA; 1: 8 FFMA2; 2: 8 FFMA + 8 IADD; 3: 8 FFMA2 + 8 IADD; 4: 4 FFMA2 + 8 IADD
// 5: 8 FFMA + 1 MUFU; 6: 8 FFMA2 + 2 MUFU; 7: 4 FFMA2 + 4 FFMA; 8: 8 FFMA2 + 16 IADD
template <int MODE> __global__ void kern(float* out, float x, float y) {
    float a[8]; u64 p[8]; unsigned q[16];
    for (int k = 0; k < 8; k++) { a[k] = x + k + threadIdx.x; float2 t = make_float2(a[k], a[k] + 1.f); p[k] = *reinterpret_cast<u64*>(&t); }
    for (int k = 0; k < 16; k++) q[k] = threadIdx.x + k;
    float2 yy = make_float2(y, y); u64 Y = *reinterpret_cast<u64*>(&yy);
    float m0 = x, m1 = y;
    for (int it = 0; it < IT; it++) {
        if (MODE == 0 || MODE == 2 || MODE == 5) {
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = ffma1(a[k], y, x);
        }
        if (MODE == 1 || MODE == 3 || MODE == 6 || MODE == 8) {
#pragma unroll
            for (int k = 0; k < 8; k++) p[k] = ffma2(p[k], Y, Y);
        }
        if (MODE == 4 || MODE == 7) {
#pragma unroll
            for (int k = 0; k < 4; k++) p[k] = ffma2(p[k], Y, Y);
        }
        if (MODE == 7) {
#pragma unroll
            for (int k = 0; k < 4; k++) a[k] = ffma1(a[k], y, x);
        }
        if (MODE == 2 || MODE == 3 || MODE == 4) {
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = iadd(q[k], q[(k + 1) & 7]);
        }
        if (MODE == 8) {
#pragma unroll
            for (int k = 0; k < 16; k++) q[k] = iadd(q[k], q[(k + 1) & 15]);
        }
        if (MODE == 5) m0 = ex2(m0);
        if (MODE == 6) { m0 = ex2(m0); m1 = ex2(m1); }
    }
    float s = m0 + m1;
    for (int k = 0; k < 8; k++) { float2 t = *reinterpret_cast<float2*>(&p[k]); s += a[k] + t.x + t.y; }
    for (int k = 0; k < 16; k++) s += (float)q[k];
    if (s == 12345.678f) out[0] = s;
}
template <int MODE> void run(const char* name, int ninstr, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int nsm = pr.multiProcessorCount;
    for (int wps = 1; wps <= 16; wps *= 2) {   // warps per SMSP
        int threads = 128 * wps > 1024 ? 1024 : 128 * wps;
        int blocks_per_sm = (128 * wps) / threads;
        dim3 g(nsm * blocks_per_sm);
        kern<MODE><<<g, threads>>>(d, 1.0f, 0.5f);
        cudaEventRecord(e0);
        kern<MODE><<<g, threads>>>(d, 1.0f, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
        double cyc = ms * 1e-3 * clk * 1e3;  // assumes max clock
        double per = cyc / ((double)IT * ninstr * wps);
        printf("%-28s warps/SMSP=%2d  %.3f ms  cycles/warp-instr/SMSP=%.3f (at %d MHz)\n", name, wps, ms, per, clk / 1000);
    }
}
int main() {
    float* d; cudaMalloc(&d, 4);
    run<0>("8 FFMA", 8, d);
    run<1>("8 FFMA2", 8, d);
    run<2>("8 FFMA + 8 IADD", 16, d);
    run<3>("8 FFMA2 + 8 IADD", 16, d);
    run<4>("4 FFMA2 + 8 IADD", 12, d);
    run<5>("8 FFMA + 1 MUFU", 9, d);
    run<6>("8 FFMA2 + 2 MUFU", 10, d);
    run<7>("4 FFMA2 + 4 FFMA", 8, d);
    run<8>("8 FFMA2 + 16 IADD", 24, d);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
