// pipe-concurrency micro-benchmark: do IMAD.WIDE (fma-heavy pipe), DFMA/DADD (fp64 pipe) and IADD3 (alu pipe) warps overlap?
// MODE bit0: even warps run A, bit1: odd warps run B (A/B chosen by template)
#include <cstdio>
#include <cstdint>
enum { OP_IMADW = 0, OP_DFMA = 1, OP_DADD = 2, OP_IADD3 = 3, OP_IMAD = 4, OP_NONE = 5 };
template <int OP> __device__ __forceinline__ void body(int iters, void* o)
{
    if (OP == OP_IMADW) {
        uint64_t a[8]; for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i; uint32_t b = blockIdx.x * 3 + 1;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 16; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = (uint64_t)((uint32_t)a[i]) * b + a[i];
        }
        uint64_t s = 0; for (int i = 0; i < 8; i++) s += a[i]; if (s == 0x1234567) ((uint64_t*)o)[threadIdx.x] = s;
    } else if (OP == OP_IMAD) {
        uint32_t a[8]; for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i; uint32_t b = blockIdx.x * 3 + 1;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 16; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = a[i] * b + a[(i + 1) & 7];
        }
        uint32_t s = 0; for (int i = 0; i < 8; i++) s += a[i]; if (s == 0x1234567) ((uint32_t*)o)[threadIdx.x] = s;
    } else if (OP == OP_DFMA) {
        double a[8]; for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i; double b = blockIdx.x * 3 + 1.000001;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 16; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, a[i]);
        }
        double s = 0; for (int i = 0; i < 8; i++) s += a[i]; if (s == 0.1234567) ((double*)o)[threadIdx.x] = s;
    } else if (OP == OP_DADD) {
        double a[8]; for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i; double b = blockIdx.x * 3 + 1.000001;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 16; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = a[i] + b;
        }
        double s = 0; for (int i = 0; i < 8; i++) s += a[i]; if (s == 0.1234567) ((double*)o)[threadIdx.x] = s;
    } else if (OP == OP_IADD3) {
        uint32_t a[8]; for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i; uint32_t b = blockIdx.x * 3 + 1;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 16; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        }
        uint32_t s = 0; for (int i = 0; i < 8; i++) s += a[i]; if (s == 0x1234567) ((uint32_t*)o)[threadIdx.x] = s;
    }
}
template <int A, int B> __global__ void __launch_bounds__(256) k_mix(void* o, int iters)
{
    if ((threadIdx.x >> 5) & 1) body<B>(iters, o); else body<A>(iters, o);
}
template <class F> float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
const char* nm[] = { "IMAD.WIDE", "DFMA", "DADD", "IADD3", "IMAD", "-" };
template <int A, int B> void run(void* o)
{
    int blocks = 148 * 8, iters = 2000;
    float ms = timeit([&] { k_mix<A, B><<<blocks, 256>>>(o, iters); });
    // each half of the warps issues iters*128 ops per thread
    double ops_half = (double)blocks * 128 * iters * 128;
    printf("%-10s | %-10s : %.3f ms   A %.2f Tops/s  B %.2f Tops/s\n", nm[A], nm[B], ms, A == OP_NONE ? 0 : ops_half / ms / 1e9, B == OP_NONE ? 0 : ops_half / ms / 1e9);
}
int main()
{
    void* o; cudaMalloc(&o, 1 << 20);
    run<OP_IMADW, OP_NONE>(o); run<OP_DFMA, OP_NONE>(o); run<OP_DADD, OP_NONE>(o); run<OP_IADD3, OP_NONE>(o); run<OP_IMAD, OP_NONE>(o);
    run<OP_IMADW, OP_IMADW>(o); run<OP_DFMA, OP_DFMA>(o); run<OP_IADD3, OP_IADD3>(o); run<OP_IMAD, OP_IMAD>(o);
    run<OP_IMADW, OP_DFMA>(o); run<OP_IMADW, OP_DADD>(o); run<OP_IMADW, OP_IADD3>(o); run<OP_DFMA, OP_IADD3>(o); run<OP_IMAD, OP_DFMA>(o); run<OP_IMAD, OP_IADD3>(o);
    return 0;
}
