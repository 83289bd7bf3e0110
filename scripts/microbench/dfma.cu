// Feasibility experiment: 254-bit Montgomery multiply on the FP64 pipe (5 x 52-bit limbs held in doubles),
// checked against the IMAD.WIDE multiply, timed alone and co-resident with IMAD warps.
#include "/root/repo/aztec-2.0_b200/csrc/field.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace bbg;

// ---- 52-bit limb representation: v = sum l[i] * 2^(52 i), l[i] integer-valued doubles in [0, 2^52)
struct F52 { double l[5]; };

__device__ __forceinline__ double u52_to_double(uint64_t v)  // v < 2^52, exact
{
    return __longlong_as_double((long long)(v | 0x4330000000000000ull)) - 4503599627370496.0; // 2^52
}
// split a*b (a, b < 2^52, integers) into hi*2^52 + lo, returned as int64 bit patterns WITH biases:
//   hi_bits = bits(2^104 + hi*2^52)  -> mantissa == hi     lo_bits = bits(2^52 + lo) -> mantissa == lo
#define C104 20282409603651670423947251286016.0            /* 2^104 */
#define C104_52 20282409603651674927546878656512.0         /* 2^104 + 2^52 */
__device__ __forceinline__ void mul52(double a, double b, long long& hi_bits, long long& lo_bits)
{
    double hi = __fma_rz(a, b, C104);
    double lo = __fma_rz(a, b, C104_52 - hi);
    hi_bits = __double_as_longlong(hi);
    lo_bits = __double_as_longlong(lo);
}
#define BIAS_HI 0x4670000000000000ll  /* bits(2^104) */
#define BIAS_LO 0x4330000000000000ll  /* bits(2^52)  */

template <class F> struct P52 {
    // p in 5 x 52-bit limbs and pinv52 = -p^-1 mod 2^52, filled on the host
    double p[5];
    double pinv;
};
__constant__ P52<FrParams> c_fr52;

// Montgomery product with R' = 2^260: r = a*b*2^-260 mod p (coarse, < 2p for inputs < 2p... checked numerically)
__device__ __forceinline__ F52 mul_f52(const F52& a, const F52& b, const P52<FrParams>& K)
{
    // column accumulators as int64 (sum of biased bit patterns; biases removed at the end of each column)
    long long col[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) col[i] = 0;
    int nhi[11], nlo[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) { nhi[i] = 0; nlo[i] = 0; }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            long long h, l;
            mul52(a.l[i], b.l[j], h, l);
            col[i + j] += l; nlo[i + j]++;
            col[i + j + 1] += h; nhi[i + j + 1]++;
        }
        // column i is complete up to carries from lower columns (already folded in): reduce it
        long long v = col[i] - (long long)nhi[i] * BIAS_HI - (long long)nlo[i] * BIAS_LO; // true integer value of column i
        uint64_t low = (uint64_t)v & 0xFFFFFFFFFFFFFull;
        double q;
        {
            long long h, l;
            mul52(u52_to_double(low), K.pinv, h, l);
            q = u52_to_double((uint64_t)(l - BIAS_LO));
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            long long h, l;
            mul52(q, K.p[j], h, l);
            col[i + j] += l; nlo[i + j]++;
            col[i + j + 1] += h; nhi[i + j + 1]++;
        }
        // now column i is divisible by 2^52: carry it into column i+1
        long long v2 = col[i] - (long long)nhi[i] * BIAS_HI - (long long)nlo[i] * BIAS_LO;
        col[i + 1] += (v2 >> 52);
    }
    F52 r;
    long long carry = 0;
#pragma unroll
    for (int k = 5; k < 10; ++k) {
        long long v = col[k] - (long long)nhi[k] * BIAS_HI - (long long)nlo[k] * BIAS_LO + carry;
        r.l[k - 5] = u52_to_double((uint64_t)v & 0xFFFFFFFFFFFFFull);
        carry = v >> 52;
    }
    // top carry must be zero for in-range inputs (value < 2p < 2^255)
    return r;
}

// conversions (host + device): 8 x u32 <-> 5 x 52
__host__ __device__ inline void limbs32_to_52(const uint32_t* w, uint64_t* o)
{
    unsigned __int128 acc = 0; int bits = 0, k = 0;
    for (int i = 0; i < 8; ++i) {
        acc |= (unsigned __int128)w[i] << bits; bits += 32;
        while (bits >= 52 && k < 4) { o[k++] = (uint64_t)acc & 0xFFFFFFFFFFFFFull; acc >>= 52; bits -= 52; }
    }
    o[4] = (uint64_t)acc;
}
__host__ __device__ inline void limbs52_to_32(const uint64_t* o, uint32_t* w)
{
    unsigned __int128 acc = 0; int bits = 0, k = 0;
    for (int i = 0; i < 5; ++i) {
        acc |= (unsigned __int128)o[i] << bits; bits += 52;
        while (bits >= 32 && k < 8) { w[k++] = (uint32_t)acc; acc >>= 32; bits -= 32; }
    }
}

__global__ void k_check(const fr_t* a, const fr_t* b, fr_t* out_imad, uint64_t* out52, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t x = fe_load<FrParams>(a + i), y = fe_load<FrParams>(b + i);
    fe_store(out_imad + i, fe_mul(x, y));
    uint64_t xa[5], ya[5];
    limbs32_to_52(x.l, xa); limbs32_to_52(y.l, ya);
    F52 fx, fy;
    for (int k = 0; k < 5; ++k) { fx.l[k] = u52_to_double(xa[k]); fy.l[k] = u52_to_double(ya[k]); }
    F52 r = mul_f52(fx, fy, c_fr52);
    for (int k = 0; k < 5; ++k) out52[i * 5 + k] = (uint64_t)(long long)r.l[k];
}

// MODE 0: all warps IMAD; 1: all warps DFMA; 2: even warps IMAD, odd warps DFMA
template <int MODE> __global__ void __launch_bounds__(256) k_chain(fr_t* o, int iters)
{
    const bool dfma = MODE == 1 || (MODE == 2 && ((threadIdx.x >> 5) & 1));
    if (!dfma) {
        fr_t x, y;
        for (int i = 0; i < 8; i++) { x.l[i] = threadIdx.x * 7 + i + 1; y.l[i] = blockIdx.x + i * 3 + 5; }
        x.l[7] &= 0x0fffffff; y.l[7] &= 0x0fffffff;
#pragma unroll 1
        for (int it = 0; it < iters; it++) { x = fe_mul(x, y); y = fe_mul(y, x); }
        if (x.l[0] == 0x12345678 && y.l[1] == 77) fe_store(o + threadIdx.x, x);
    } else {
        F52 x, y;
        for (int i = 0; i < 5; i++) { x.l[i] = (double)(threadIdx.x * 7 + i + 1); y.l[i] = (double)(blockIdx.x + i * 3 + 5); }
#pragma unroll 1
        for (int it = 0; it < iters; it++) { x = mul_f52(x, y, c_fr52); y = mul_f52(y, x, c_fr52); }
        if (x.l[0] == 0.5 && y.l[1] == 77.25) o[threadIdx.x].l[0] = (uint32_t)x.l[2];
    }
}

typedef unsigned __int128 u128;
struct Big { uint64_t w[5]; }; // 320 bits
int main()
{
    // constants
    uint32_t pw[8]; for (int i = 0; i < 8; i++) pw[i] = FrParams::P(i);
    uint64_t p52[5]; limbs32_to_52(pw, p52);
    // pinv52 = -p^-1 mod 2^52 (Newton)
    uint64_t p0 = p52[0], inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - p0 * inv;
    uint64_t pinv = (0 - inv) & 0xFFFFFFFFFFFFFull;
    P52<FrParams> K; for (int i = 0; i < 5; i++) K.p[i] = (double)p52[i]; K.pinv = (double)pinv;
    cudaMemcpyToSymbol(c_fr52, &K, sizeof(K));

    const int n = 1 << 16;
    std::vector<uint32_t> ha(n * 8), hb(n * 8);
    srand(1);
    for (int i = 0; i < n; i++) for (int k = 0; k < 8; k++) { ha[i*8+k] = ((uint32_t)rand() << 16) ^ rand(); hb[i*8+k] = ((uint32_t)rand() << 16) ^ rand(); }
    for (int i = 0; i < n; i++) { ha[i*8+7] %= 0x60000000; hb[i*8+7] %= 0x60000000; } // < 2p
    fr_t *da, *db, *dout; uint64_t* d52;
    cudaMalloc(&da, n * 32); cudaMalloc(&db, n * 32); cudaMalloc(&dout, n * 32); cudaMalloc(&d52, n * 40);
    cudaMemcpy(da, ha.data(), n * 32, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), n * 32, cudaMemcpyHostToDevice);
    k_check<<<n / 128, 128>>>(da, db, dout, d52, n);
    std::vector<uint32_t> ho(n * 8); std::vector<uint64_t> h52(n * 5);
    cudaMemcpy(ho.data(), dout, n * 32, cudaMemcpyDeviceToHost); cudaMemcpy(h52.data(), d52, n * 40, cudaMemcpyDeviceToHost);
    // check: r52 * 2^260 == a*b (mod p)  <=>  r52 * 2^4 == r32 * ... : r32 = a b 2^-256, r52 = a b 2^-260  => r32 == 16 * r52 mod p
    // do it with python-free big ints: compare (r52 * 16) mod p with r32 mod p using 320-bit arithmetic
    int bad = 0, over = 0;
    for (int i = 0; i < n; i++) {
        uint32_t w[8]; limbs52_to_32(&h52[i * 5], w);
        // value52 as 5x64
        u128 acc = 0; uint64_t v[5] = {0,0,0,0,0};
        // times 16
        uint64_t src[4]; for (int k = 0; k < 4; k++) src[k] = (uint64_t)w[2*k] | ((uint64_t)w[2*k+1] << 32);
        uint64_t t[5]; t[4] = src[3] >> 60; for (int k = 3; k > 0; k--) t[k] = (src[k] << 4) | (src[k-1] >> 60); t[0] = src[0] << 4;
        // reduce mod p by repeated subtraction (t < 32p)
        uint64_t P4[5] = { (uint64_t)pw[0] | ((uint64_t)pw[1] << 32), (uint64_t)pw[2] | ((uint64_t)pw[3] << 32), (uint64_t)pw[4] | ((uint64_t)pw[5] << 32), (uint64_t)pw[6] | ((uint64_t)pw[7] << 32), 0 };
        auto geq = [&](uint64_t* x) { for (int k = 4; k >= 0; k--) { if (x[k] != P4[k]) return x[k] > P4[k]; } return true; };
        auto sub = [&](uint64_t* x) { u128 br = 0; for (int k = 0; k < 5; k++) { u128 d = (u128)x[k] - P4[k] - br; x[k] = (uint64_t)d; br = (d >> 64) & 1; } };
        int guard = 0; while (geq(t) && guard++ < 100) sub(t);
        uint64_t r32[5]; for (int k = 0; k < 4; k++) r32[k] = (uint64_t)ho[i*8+2*k] | ((uint64_t)ho[i*8+2*k+1] << 32); r32[4] = 0;
        guard = 0; while (geq(r32) && guard++ < 100) sub(r32);
        bool eq = true; for (int k = 0; k < 5; k++) eq &= (t[k] == r32[k]);
        if (!eq) { if (bad < 3) printf("mismatch at %d\n", i); bad++; }
        if (h52[i*5+4] >> 47) over++;
        (void)acc; (void)v;
    }
    printf("check: %d mismatches of %d, %d results with top limb >= 2^47\n", bad, n, over);

    auto timeit = [&](auto f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; };
    int blocks = 148 * 8, iters = 1000;
    float ms;
    ms = timeit([&] { k_chain<0><<<blocks, 256>>>(dout, iters); }); printf("IMAD only : %.3f ms %.1f Gmul/s\n", ms, (double)blocks * 256 * iters * 2 / ms / 1e6);
    ms = timeit([&] { k_chain<1><<<blocks, 256>>>(dout, iters); }); printf("DFMA only : %.3f ms %.1f Gmul/s\n", ms, (double)blocks * 256 * iters * 2 / ms / 1e6);
    ms = timeit([&] { k_chain<2><<<blocks, 256>>>(dout, iters); }); printf("mixed 1:1 : %.3f ms %.1f Gmul/s\n", ms, (double)blocks * 256 * iters * 2 / ms / 1e6);
    return 0;
}
