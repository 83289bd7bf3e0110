// field.cuh -- BN254 fq / fr arithmetic held entirely in registers (8 x 32-bit limbs).
//
// Semantics follow barretenberg's field<Params> exactly (bb/ecc/fields/field_impl.hpp:34-198,
// field_impl_generic.hpp:171-272, 392-499): Montgomery form with R = 2^256, values kept only
// coarsely reduced in [0, 2p); mul/sqr have no final subtraction; add subtracts 2p when the sum
// reaches it; sub adds 2p on borrow; reduce_once is one conditional subtraction of p; is_zero
// accepts 0 or p.  Memory layout = the reference's 4 x u64 little-endian limbs = our 8 x u32.
#pragma once
#include <cstdint>

#include "mont_asm.inc"

namespace bbg {

struct FqParams {
    // p = 0x30644e72e131a029 b85045b68181585d 97816a916871ca8d 3c208c16d87cfd47 (bb/ecc/curves/bn254/fq.hpp:11-14)
    static __host__ __device__ constexpr uint32_t P(int i)
    {
        constexpr uint32_t t[8] = { 0xd87cfd47, 0x3c208c16, 0x6871ca8d, 0x97816a91, 0x8181585d, 0xb85045b6, 0xe131a029, 0x30644e72 };
        return t[i];
    }
    static __host__ __device__ constexpr uint32_t P2(int i)
    {
        constexpr uint32_t t[8] = { 0xb0f9fa8e, 0x7841182d, 0xd0e3951a, 0x2f02d522, 0x0302b0bb, 0x70a08b6d, 0xc2634053, 0x60c89ce5 };
        return t[i];
    }
    // R^2 mod p (fq.hpp:16-19)
    static __host__ __device__ constexpr uint32_t R2(int i)
    {
        constexpr uint32_t t[8] = { 0x538afa89, 0xf32cfc5b, 0xd44501fb, 0xb5e71911, 0x0a417ff6, 0x47ab1eff, 0xcab8351f, 0x06d89f71 };
        return t[i];
    }
    // R mod p
    static __host__ __device__ constexpr uint32_t ONE(int i)
    {
        constexpr uint32_t t[8] = { 0xc58f0d9d, 0xd35d438d, 0xf5c70b3d, 0x0a78eb28, 0x7879462c, 0x666ea36f, 0x9a07df2f, 0x0e0a77c1 };
        return t[i];
    }
    static constexpr uint32_t NINV = 0xe4866389; // -p^-1 mod 2^32 (low half of fq.hpp:41 r_inv)
    static constexpr bool IS_FQ = true;
};
struct FrParams {
    // r = 0x30644e72e131a029 b85045b68181585d 2833e84879b97091 43e1f593f0000001 (bb/ecc/curves/bn254/fr.hpp:12-15)
    static __host__ __device__ constexpr uint32_t P(int i)
    {
        constexpr uint32_t t[8] = { 0xf0000001, 0x43e1f593, 0x79b97091, 0x2833e848, 0x8181585d, 0xb85045b6, 0xe131a029, 0x30644e72 };
        return t[i];
    }
    static __host__ __device__ constexpr uint32_t P2(int i)
    {
        constexpr uint32_t t[8] = { 0xe0000002, 0x87c3eb27, 0xf372e122, 0x5067d090, 0x0302b0ba, 0x70a08b6d, 0xc2634053, 0x60c89ce5 };
        return t[i];
    }
    static __host__ __device__ constexpr uint32_t R2(int i)
    {
        constexpr uint32_t t[8] = { 0xae216da7, 0x1bb8e645, 0xe35c59e3, 0x53fe3ab1, 0x53bb8085, 0x8c49833d, 0x7f4e44a5, 0x0216d0b1 };
        return t[i];
    }
    static __host__ __device__ constexpr uint32_t ONE(int i)
    {
        constexpr uint32_t t[8] = { 0x4ffffffb, 0xac96341c, 0x9f60cd29, 0x36fc7695, 0x7879462e, 0x666ea36f, 0x9a07df2f, 0x0e0a77c1 };
        return t[i];
    }
    static constexpr uint32_t NINV = 0xefffffff; // low half of fr.hpp:42 r_inv
    static constexpr bool IS_FQ = false;
};

template <class F> struct alignas(16) Fe {
    uint32_t l[8];
};
using fq_t = Fe<FqParams>;
using fr_t = Fe<FrParams>;

template <class F> __host__ __device__ __forceinline__ Fe<F> fe_zero()
{
    Fe<F> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = 0;
    return r;
}
template <class F> __host__ __device__ __forceinline__ Fe<F> fe_one()
{
    Fe<F> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = F::ONE(i);
    return r;
}

// ---- 128-bit vector memory access (an fe is two 16-byte halves)
template <class F> __device__ __forceinline__ Fe<F> fe_load(const void* p)
{
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fe<F> r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
template <class F> __device__ __forceinline__ Fe<F> fe_load_nc(const void* p)
{
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fe<F> r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
template <class F> __device__ __forceinline__ void fe_store(void* p, const Fe<F>& v)
{
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

#ifdef __CUDACC__
// ---- carry-chain helpers (each chain lives in ONE asm statement so the CC flag never escapes)
__device__ __forceinline__ void add8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8])
{
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}
// r = a - b, returns borrow mask (0xffffffff when a < b)
__device__ __forceinline__ uint32_t sub8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8])
{
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(borrow)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return borrow;
}
// r = a - K (compile-time constant limbs), returns borrow mask
template <class F, bool TWICE> __device__ __forceinline__ uint32_t sub8_mod(uint32_t (&r)[8], const uint32_t (&a)[8])
{
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(borrow)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "n"(TWICE ? F::P2(0) : F::P(0)), "n"(TWICE ? F::P2(1) : F::P(1)), "n"(TWICE ? F::P2(2) : F::P(2)),
          "n"(TWICE ? F::P2(3) : F::P(3)), "n"(TWICE ? F::P2(4) : F::P(4)), "n"(TWICE ? F::P2(5) : F::P(5)),
          "n"(TWICE ? F::P2(6) : F::P(6)), "n"(TWICE ? F::P2(7) : F::P(7)));
    return borrow;
}
// r = a + (mask & 2p)
template <class F> __device__ __forceinline__ void add8_masked_2p(uint32_t (&r)[8], const uint32_t (&a)[8], uint32_t mask)
{
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = F::P2(i) & mask;
    add8(r, a, m);
}

// ---- field ops
template <class F> __device__ __forceinline__ Fe<F> fe_add(const Fe<F>& a, const Fe<F>& b)
{
    Fe<F> s, t;
    add8(s.l, a.l, b.l);                       // < 4p < 2^256
    uint32_t borrow = sub8_mod<F, true>(t.l, s.l); // s - 2p
#pragma unroll
    for (int i = 0; i < 8; ++i) s.l[i] = borrow ? s.l[i] : t.l[i];
    return s;
}
template <class F> __device__ __forceinline__ Fe<F> fe_sub(const Fe<F>& a, const Fe<F>& b)
{
    Fe<F> d, r;
    uint32_t borrow = sub8(d.l, a.l, b.l);
    add8_masked_2p<F>(r.l, d.l, borrow);
    return r;
}
// a - b + 2p with NO reduction: for a, b in [0, 2p) the result lies in (0, 4p) < 2^256.  Only legal as the left operand of
// a Montgomery multiply whose other operand is canonical (< p): (4p * p) / 2^256 + p < 1.76 p, back inside [0, 2p).
// Saves the borrow extraction and the eight masking instructions of fe_sub (used by the NTT butterflies, whose
// twiddle tables are stored canonical).
template <class F> __device__ __forceinline__ Fe<F> fe_sub_lazy(const Fe<F>& a, const Fe<F>& b)
{
    Fe<F> d, r;
    asm("sub.cc.u32 %0, %8, %16;\n\t"
        "subc.cc.u32 %1, %9, %17;\n\t"
        "subc.cc.u32 %2, %10, %18;\n\t"
        "subc.cc.u32 %3, %11, %19;\n\t"
        "subc.cc.u32 %4, %12, %20;\n\t"
        "subc.cc.u32 %5, %13, %21;\n\t"
        "subc.cc.u32 %6, %14, %22;\n\t"
        "subc.u32 %7, %15, %23;"
        : "=r"(d.l[0]), "=r"(d.l[1]), "=r"(d.l[2]), "=r"(d.l[3]), "=r"(d.l[4]), "=r"(d.l[5]), "=r"(d.l[6]), "=r"(d.l[7])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
        : "r"(d.l[0]), "r"(d.l[1]), "r"(d.l[2]), "r"(d.l[3]), "r"(d.l[4]), "r"(d.l[5]), "r"(d.l[6]), "r"(d.l[7]),
          "n"(F::P2(0)), "n"(F::P2(1)), "n"(F::P2(2)), "n"(F::P2(3)), "n"(F::P2(4)), "n"(F::P2(5)), "n"(F::P2(6)), "n"(F::P2(7)));
    return r;
}
template <class F> __device__ __forceinline__ Fe<F> fe_dbl(const Fe<F>& a) { return fe_add(a, a); }
// -a = 2p - a (field_impl.hpp:148-157); maps 0 -> 2p which is still a legal coarse zero? No: 2p is out of
// range, so zero is mapped to zero explicitly.
template <class F> __device__ __forceinline__ Fe<F> fe_neg(const Fe<F>& a)
{
    Fe<F> p2, r;
#pragma unroll
    for (int i = 0; i < 8; ++i) p2.l[i] = F::P2(i);
    sub8(r.l, p2.l, a.l);
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) any |= a.l[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = any ? r.l[i] : 0u;
    return r;
}
template <class F> __device__ __forceinline__ Fe<F> fe_reduce_once(const Fe<F>& a)
{
    Fe<F> t, r;
    uint32_t borrow = sub8_mod<F, false>(t.l, a.l);
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = borrow ? a.l[i] : t.l[i];
    return r;
}
template <class F> __device__ __forceinline__ bool fe_is_zero(const Fe<F>& a)
{
    uint32_t z = 0, e = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        z |= a.l[i];
        e |= a.l[i] ^ F::P(i);
    }
    return z == 0 || e == 0;
}
template <class F> __device__ __forceinline__ bool fe_eq(const Fe<F>& a, const Fe<F>& b)
{
    return fe_is_zero(fe_sub(a, b));
}

#ifdef BBG_PORTABLE_MUL
// plain C++ CIOS (64-bit accumulators); kept as an on-device cross-check of the PTX path
template <class F> __device__ __forceinline__ Fe<F> fe_mul(const Fe<F>& a, const Fe<F>& b)
{
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c += (uint64_t)a.l[j] * b.l[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * F::NINV;
        c = (uint64_t)m * F::P(0) + t[0];
        c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            c += (uint64_t)m * F::P(j) + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[7] = (uint32_t)c;
        t[8] = t[9] + (uint32_t)(c >> 32);
    }
    Fe<F> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = t[i];
    return r;
}
template <class F> __device__ __forceinline__ Fe<F> fe_sqr(const Fe<F>& a) { return fe_mul(a, a); }
#else
template <class F> __device__ __forceinline__ Fe<F> fe_mul(const Fe<F>& a, const Fe<F>& b)
{
    Fe<F> r;
    if constexpr (F::IS_FQ) {
        fq_mul_ptx(r.l, a.l, b.l);
    } else {
        fr_mul_ptx(r.l, a.l, b.l);
    }
    return r;
}
template <class F> __device__ __forceinline__ Fe<F> fe_sqr(const Fe<F>& a)
{
    Fe<F> r;
    if constexpr (F::IS_FQ) {
        fq_sqr_ptx(r.l, a.l);
    } else {
        fr_sqr_ptx(r.l, a.l);
    }
    return r;
}
#endif

// from_montgomery_form: * 1 then reduce_once => canonical (field_impl.hpp:246-250)
template <class F> __device__ __forceinline__ Fe<F> fe_from_mont(const Fe<F>& a)
{
    Fe<F> one = fe_zero<F>();
    one.l[0] = 1;
    return fe_reduce_once(fe_mul(a, one));
}
// to_montgomery_form: * R^2, reduce_once (field_impl.hpp:234-244); input must be < 2p
template <class F> __device__ __forceinline__ Fe<F> fe_to_mont(const Fe<F>& a)
{
    Fe<F> r2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r2.l[i] = F::R2(i);
    return fe_reduce_once(fe_mul(a, r2));
}
// a^e, e a 64-bit exponent (square-and-multiply, MSB first)
template <class F> __device__ __forceinline__ Fe<F> fe_pow(const Fe<F>& a, uint64_t e)
{
    Fe<F> acc = fe_one<F>();
    if (e == 0) {
        return acc;
    }
    int msb = 63 - __clzll((long long)e);
    for (int i = msb; i >= 0; --i) {
        acc = fe_sqr(acc);
        if ((e >> i) & 1) {
            acc = fe_mul(acc, a);
        }
    }
    return acc;
}
#endif // __CUDACC__

} // namespace bbg
