// Micro-benchmark: issue/throughput of FFMA vs FFMA2 (fma.rn.f32x2) and mixed ALU on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
    float a[8]; unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; p[i] = (unsigned long long)__float_as_uint(a[i]) * 0x100000001ull; }
    const float m = 1.0000001f, c = 1e-9f;
    const unsigned long long m2 = (unsigned long long)__float_as_uint(m) * 0x100000001ull, c2 = (unsigned long long)__float_as_uint(c) * 0x100000001ull;
    unsigned ia[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ia[i] = threadIdx.x * 7 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], m, c);
            if (MODE == 1) p[i] = fma2(p[i], m2, c2);
            if (MODE == 2) { a[i] = fmaf(a[i], m, c); ia[i] = (ia[i] ^ 0x5bd1e995u) + (ia[i] >> 3); }   // FFMA + 2 ALU
            if (MODE == 3) { p[i] = fma2(p[i], m2, c2); ia[i] = (ia[i] ^ 0x5bd1e995u) + (ia[i] >> 3); } // FFMA2 + 2 ALU
            if (MODE == 4) { ia[i] = (ia[i] ^ 0x5bd1e995u) + (ia[i] >> 3); }
            if (MODE == 5) { a[i] = fmaxf(a[i] * m, c); }   // FMUL + FMNMX
        }
    }
    long long t1 = clock64();
    float s = 0; unsigned long long q = 0; unsigned iq = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += a[i]; q ^= p[i]; iq += ia[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(q & 0xffff) + (float)iq;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int threads, int opsPerIter) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<MODE><<<148, threads>>>(out, cyc, 1.f); cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, cyc, 1.f); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double warpInstr = (double)ITERS * 8 * opsPerIter * (threads / 32);
    printf("%-22s threads %4d  cycles %.0f  warp-instr/cycle/SM %.3f  (per SMSP %.3f)\n", name, threads, c, warpInstr / c, warpInstr / c / 4);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {128, 256, 512, 1024}) {
        run<0>("FFMA", th, 1);
        run<1>("FFMA2", th, 1);
        run<2>("FFMA+3ALU", th, 4);
        run<3>("FFMA2+3ALU", th, 4);
        run<4>("3ALU", th, 3);
        run<5>("FMUL+FMNMX", th, 2);
    }
    return 0;
}
