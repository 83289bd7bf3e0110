// Second DFMA experiment: is a 254-bit Montgomery multiply on the FP64 pipe (5 x 52-bit limbs) worth building into the
// kernels?  Three measurements on the same launch shape as bbg_bench_field_mul:
//   A. the bare building block: N independent 52x52 -> 104-bit products (DFMA.RZ, DADD, DFMA.RZ) each accumulated into
//      two 64-bit integer columns -- the sustained products/s of the instruction mix with perfect ILP;
//   B. a full multiply (q on the integer pipe via mul.lo.u64, biases folded into the column initialisers), one and two
//      independent chains per thread, checked against the IMAD.WIDE multiply;
//   C. the IMAD.WIDE multiply of field.cuh in the same harness (baseline).
#include "/root/repo/aztec-2.0_b200/csrc/field.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace bbg;

#define C104 20282409603651670423947251286016.0    /* 2^104 */
#define C104_52 20282409603651674927546878656512.0 /* 2^104 + 2^52 */
#define C52 4503599627370496.0                     /* 2^52 */
#define BH 0x4670000000000000ll                    /* bits(2^104) */
#define BL 0x4330000000000000ll                    /* bits(2^52)  */
#define M52 0xFFFFFFFFFFFFFull

struct F52 {
    double l[5];
};
struct K52 {
    double p[5];
    unsigned long long pinv; // -p^-1 mod 2^52
};
__constant__ K52 c_k52;

__device__ __forceinline__ void mul52(double a, double b, long long& hi, long long& lo)
{
    double h = __fma_rz(a, b, C104);
    double t = C104_52 - h;
    double l = __fma_rz(a, b, t);
    hi = __double_as_longlong(h);
    lo = __double_as_longlong(l);
}
__device__ __forceinline__ double u52_to_double(unsigned long long v) // v < 2^52
{
    return __longlong_as_double((long long)(v | (unsigned long long)BL)) - C52;
}

// bias bookkeeping: number of lo / hi bit patterns added to column k BEFORE column k is read (see text in DESIGN)
__host__ __device__ constexpr int n_lo_ab(int k) { int c = 0; for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) if (i + j == k) c++; return c; }
__host__ __device__ constexpr int n_hi_ab(int k) { return k >= 1 ? n_lo_ab(k - 1) : 0; }
// q_i * p_j: lo -> col i+j, hi -> col i+j+1.  Counted for column k only if the step i < k (lo with j >= 1, hi always)
__host__ __device__ constexpr int n_lo_qp(int k) { int c = 0; for (int i = 0; i < 5; i++) for (int j = 1; j < 5; j++) if (i + j == k) c++; return c; }
__host__ __device__ constexpr int n_hi_qp(int k) { int c = 0; for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) if (i + j + 1 == k) c++; return c; }
__host__ __device__ constexpr long long col_init(int k)
{
    return (long long)(0ull - ((unsigned long long)(n_lo_ab(k) + n_lo_qp(k)) * (unsigned long long)BL +
                               (unsigned long long)(n_hi_ab(k) + n_hi_qp(k)) * (unsigned long long)BH));
}

// r = a * b * 2^-260 mod p, limbs normalised to < 2^52 (value < 2p for inputs < 2p... checked numerically below)
__device__ __forceinline__ F52 mul_f52(const F52& a, const F52& b)
{
    long long col[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) col[k] = col_init(k);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            long long h, l;
            mul52(a.l[i], b.l[j], h, l);
            col[i + j] += l;
            col[i + j + 1] += h;
        }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const unsigned long long low = (unsigned long long)col[i] & M52;
        const unsigned long long q = (low * c_k52.pinv) & M52; // integer pipe (IMAD.WIDE + 2 IMAD), otherwise idle
        const double qd = u52_to_double(q);
        long long h0, l0;
        mul52(qd, c_k52.p[0], h0, l0);
        col[i + 1] += h0 + ((col[i] + l0 - BL) >> 52); // column i is now divisible by 2^52: carry it up
#pragma unroll
        for (int j = 1; j < 5; ++j) {
            long long h, l;
            mul52(qd, c_k52.p[j], h, l);
            col[i + j] += l;
            col[i + j + 1] += h;
        }
    }
    F52 r;
    long long carry = 0;
#pragma unroll
    for (int k = 5; k < 10; ++k) {
        long long v = col[k] + carry;
        r.l[k - 5] = u52_to_double((unsigned long long)v & M52);
        carry = v >> 52;
    }
    return r;
}

// ---- conversions (host + device): 8 x u32 <-> 5 x 52
__host__ __device__ inline void limbs32_to_52(const uint32_t* w, uint64_t* o)
{
    unsigned __int128 acc = 0;
    int bits = 0, k = 0;
    for (int i = 0; i < 8; ++i) {
        acc |= (unsigned __int128)w[i] << bits;
        bits += 32;
        while (bits >= 52 && k < 4) {
            o[k++] = (uint64_t)acc & M52;
            acc >>= 52;
            bits -= 52;
        }
    }
    o[4] = (uint64_t)acc;
}
__host__ __device__ inline void limbs52_to_32(const uint64_t* o, uint32_t* w)
{
    unsigned __int128 acc = 0;
    int bits = 0, k = 0;
    for (int i = 0; i < 5; ++i) {
        acc |= (unsigned __int128)o[i] << bits;
        bits += 52;
        while (bits >= 32 && k < 8) {
            w[k++] = (uint32_t)acc;
            acc >>= 32;
            bits -= 32;
        }
    }
}

__global__ void k_check(const fr_t* a, const fr_t* b, fr_t* out_imad, uint64_t* out52, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t x = fe_load<FrParams>(a + i), y = fe_load<FrParams>(b + i);
    fe_store(out_imad + i, fe_mul(x, y));
    uint64_t xa[5], ya[5];
    limbs32_to_52(x.l, xa);
    limbs32_to_52(y.l, ya);
    F52 fx, fy;
    for (int k = 0; k < 5; ++k) {
        fx.l[k] = u52_to_double(xa[k]);
        fy.l[k] = u52_to_double(ya[k]);
    }
    F52 r = mul_f52(fx, fy);
    for (int k = 0; k < 5; ++k) out52[i * 5 + k] = (uint64_t)(long long)r.l[k];
}

// A: bare product mix.  NPROD independent products per iteration, accumulated into 2 x NACC integer columns.
template <int NPROD> __global__ void __launch_bounds__(256) k_products(long long* o, int iters)
{
    double a[5], b[5];
    for (int i = 0; i < 5; i++) {
        a[i] = (double)(threadIdx.x * 7 + i + 1);
        b[i] = (double)(blockIdx.x + i * 3 + 5);
    }
    long long col[10];
    for (int i = 0; i < 10; i++) col[i] = i;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int pidx = 0; pidx < NPROD; ++pidx) {
            long long h, l;
            mul52(a[pidx % 5], b[(pidx / 5) % 5], h, l);
            col[(pidx % 5 + (pidx / 5) % 5)] += l;
            col[(pidx % 5 + (pidx / 5) % 5) + 1] += h;
        }
        // keep the operands changing without touching the fp64 pipe
        a[0] = __longlong_as_double((__double_as_longlong(a[0]) & ~0xFll) | (col[3] & 0xF));
    }
    long long s = 0;
    for (int i = 0; i < 10; i++) s += col[i];
    if (s == 0x123456789) o[threadIdx.x] = s;
}

// B / C: multiply chains.  CHAINS independent (x, y) pairs per thread.
template <int CHAINS> __global__ void __launch_bounds__(256) k_chain_dfma(double* o, int iters)
{
    F52 x[CHAINS], y[CHAINS];
    for (int c = 0; c < CHAINS; c++)
        for (int i = 0; i < 5; i++) {
            x[c].l[i] = (double)(threadIdx.x * 7 + i + 1 + c);
            y[c].l[i] = (double)(blockIdx.x + i * 3 + 5 + 11 * c);
        }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = mul_f52(x[c], y[c]);
#pragma unroll
        for (int c = 0; c < CHAINS; c++) y[c] = mul_f52(y[c], x[c]);
    }
    double s = 0;
    for (int c = 0; c < CHAINS; c++) s += x[c].l[0] + y[c].l[1];
    if (s == 0.5) o[threadIdx.x] = s;
}
template <int CHAINS> __global__ void __launch_bounds__(256) k_chain_imad(fr_t* o, int iters)
{
    fr_t x[CHAINS], y[CHAINS];
    for (int c = 0; c < CHAINS; c++) {
        for (int i = 0; i < 8; i++) {
            x[c].l[i] = threadIdx.x * 7 + i + 1 + c;
            y[c].l[i] = blockIdx.x + i * 3 + 5 + 11 * c;
        }
        x[c].l[7] &= 0x0fffffff;
        y[c].l[7] &= 0x0fffffff;
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = fe_mul(x[c], y[c]);
#pragma unroll
        for (int c = 0; c < CHAINS; c++) y[c] = fe_mul(y[c], x[c]);
    }
    uint32_t s = 0;
    for (int c = 0; c < CHAINS; c++) s += x[c].l[0] ^ y[c].l[1];
    if (s == 0x12345678) fe_store(o + threadIdx.x, x[0]);
}

typedef unsigned __int128 u128;
int main()
{
    uint32_t pw[8];
    for (int i = 0; i < 8; i++) pw[i] = FrParams::P(i);
    uint64_t p52[5];
    limbs32_to_52(pw, p52);
    uint64_t p0 = p52[0], inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - p0 * inv;
    K52 K;
    for (int i = 0; i < 5; i++) K.p[i] = (double)p52[i];
    K.pinv = (0 - inv) & M52;
    cudaMemcpyToSymbol(c_k52, &K, sizeof(K));

    const int n = 1 << 16;
    std::vector<uint32_t> ha(n * 8), hb(n * 8);
    srand(1);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 8; k++) {
            ha[i * 8 + k] = ((uint32_t)rand() << 16) ^ rand();
            hb[i * 8 + k] = ((uint32_t)rand() << 16) ^ rand();
        }
    for (int i = 0; i < n; i++) {
        ha[i * 8 + 7] %= 0x60000000;
        hb[i * 8 + 7] %= 0x60000000;
    } // < 2p
    fr_t *da, *db, *dout;
    uint64_t* d52;
    cudaMalloc(&da, n * 32);
    cudaMalloc(&db, n * 32);
    cudaMalloc(&dout, n * 32);
    cudaMalloc(&d52, n * 40);
    cudaMemcpy(da, ha.data(), n * 32, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), n * 32, cudaMemcpyHostToDevice);
    k_check<<<n / 128, 128>>>(da, db, dout, d52, n);
    std::vector<uint32_t> ho(n * 8);
    std::vector<uint64_t> h52(n * 5);
    cudaMemcpy(ho.data(), dout, n * 32, cudaMemcpyDeviceToHost);
    cudaMemcpy(h52.data(), d52, n * 40, cudaMemcpyDeviceToHost);
    int bad = 0, over = 0, unnorm = 0;
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 5; k++)
            if (h52[i * 5 + k] >> 52) unnorm++;
        uint32_t w[8];
        limbs52_to_32(&h52[i * 5], w);
        uint64_t src[4];
        for (int k = 0; k < 4; k++) src[k] = (uint64_t)w[2 * k] | ((uint64_t)w[2 * k + 1] << 32);
        uint64_t t[5];
        t[4] = src[3] >> 60;
        for (int k = 3; k > 0; k--) t[k] = (src[k] << 4) | (src[k - 1] >> 60);
        t[0] = src[0] << 4;
        uint64_t P4[5] = { (uint64_t)pw[0] | ((uint64_t)pw[1] << 32), (uint64_t)pw[2] | ((uint64_t)pw[3] << 32),
                           (uint64_t)pw[4] | ((uint64_t)pw[5] << 32), (uint64_t)pw[6] | ((uint64_t)pw[7] << 32), 0 };
        auto geq = [&](uint64_t* x) {
            for (int k = 4; k >= 0; k--) {
                if (x[k] != P4[k]) return x[k] > P4[k];
            }
            return true;
        };
        auto sub = [&](uint64_t* x) {
            u128 br = 0;
            for (int k = 0; k < 5; k++) {
                u128 d = (u128)x[k] - P4[k] - br;
                x[k] = (uint64_t)d;
                br = (d >> 64) & 1;
            }
        };
        int guard = 0;
        while (geq(t) && guard++ < 100) sub(t);
        uint64_t r32[5];
        for (int k = 0; k < 4; k++) r32[k] = (uint64_t)ho[i * 8 + 2 * k] | ((uint64_t)ho[i * 8 + 2 * k + 1] << 32);
        r32[4] = 0;
        guard = 0;
        while (geq(r32) && guard++ < 100) sub(r32);
        bool eq = true;
        for (int k = 0; k < 5; k++) eq &= (t[k] == r32[k]);
        if (!eq) {
            if (bad < 3) printf("mismatch at %d\n", i);
            bad++;
        }
        if (h52[i * 5 + 4] >> 47) over++;
    }
    printf("check: %d mismatches of %d, %d unnormalised limbs, %d results with top limb >= 2^47\n", bad, n, unnorm, over);

    auto timeit = [&](auto f) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        f();
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    };
    const int blocks = 148 * 8, iters = 1000;
    float ms;
    const double thr = (double)blocks * 256;
    ms = timeit([&] { k_products<25><<<blocks, 256>>>((long long*)dout, iters); });
    printf("A products x25 : %.3f ms  %.2f T products/s  (= %.1f G mul/s at 50 products per multiply)\n", ms, thr * iters * 25 / ms / 1e9,
           thr * iters * 25 / 50 / ms / 1e6);
    ms = timeit([&] { k_products<50><<<blocks, 256>>>((long long*)dout, iters); });
    printf("A products x50 : %.3f ms  %.2f T products/s  (= %.1f G mul/s)\n", ms, thr * iters * 50 / ms / 1e9, thr * iters * 50 / 50 / ms / 1e6);
    ms = timeit([&] { k_chain_dfma<1><<<blocks, 256>>>((double*)dout, iters); });
    printf("B DFMA 1 chain : %.3f ms %.1f Gmul/s\n", ms, thr * iters * 2 / ms / 1e6);
    ms = timeit([&] { k_chain_dfma<2><<<blocks, 256>>>((double*)dout, iters); });
    printf("B DFMA 2 chains: %.3f ms %.1f Gmul/s\n", ms, thr * iters * 4 / ms / 1e6);
    ms = timeit([&] { k_chain_dfma<2><<<blocks, 128>>>((double*)dout, iters); });
    printf("B DFMA 2 chains, 128 thr: %.3f ms %.1f Gmul/s\n", ms, thr / 2 * iters * 4 / ms / 1e6);
    ms = timeit([&] { k_chain_imad<1><<<blocks, 256>>>(dout, iters); });
    printf("C IMAD 1 chain : %.3f ms %.1f Gmul/s\n", ms, thr * iters * 2 / ms / 1e6);
    ms = timeit([&] { k_chain_imad<2><<<blocks, 256>>>(dout, iters); });
    printf("C IMAD 2 chains: %.3f ms %.1f Gmul/s\n", ms, thr * iters * 4 / ms / 1e6);
    return 0;
}
