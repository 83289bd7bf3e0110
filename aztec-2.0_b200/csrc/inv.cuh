// inv.cuh -- modular inversion on the ALU pipe (shifts, adds, selects; no multiplies in the main loop).
//
// barretenberg inverts with a^(p-2) (bb/ecc/fields/field_impl.hpp:323-329): ~254 squarings + ~125 multiplies,
// i.e. ~200 k cycles of the IMAD.WIDE pipe that bounds every MSM kernel here.  The batched-affine bucket
// accumulation (msm.cu, k_msm_pair_pass) needs one inversion per batch of additions, so this header gives it
// a different algorithm with the same result: Kaliski's "almost Montgomery inverse" (a binary extended
// Euclid that only shifts, adds and subtracts 256-bit integers) followed by ONE Montgomery multiplication by
// a tabulated power of two.  The loop issues only IADD3 / SHF / LOP3 / SEL, which run on the ALU pipe next to
// other warps' IMAD.WIDE traffic, and it is branch-free per iteration, so 32 lanes invert 32 different
// values in lock step (the trip count is the warp maximum, 254..508).
//
// Formulation (u stays odd; the roles of (u, r) and (v, s) are swapped instead of handling "u > v" apart):
//     u = p, v = a, r = 0, s = 1, k = 0, sigma = +1        invariants: a s = sigma v 2^k, a r = -sigma u 2^k (mod p)
//     while v != 0:
//         if v odd:
//             if v < u: swap(u, v); swap(r, s); sigma = -sigma
//             v = v - u; s = s + r
//         v >>= 1; r <<= 1; k += 1
//     => u = 1 and a^-1 2^k = -sigma r (mod p),  r < 2p,  254 <= k <= 508 for a 254-bit p.
// Montgomery fix-up: for a = d R (Montgomery form of d) the wanted d^-1 R equals montmul(a^-1 2^k, C_k) with
// C_k = R^3 2^-k mod p; the 255 constants C_254..C_508 are built on the host (inv_fix_table) from R^2 = C_256.
//
// The arithmetic below is plain C++ on 8 x u32 limbs so that the same source is unit-tested on the host
// (tests/test_host_inverse.py builds it with g++); nvcc turns the carry chains into IADD3.X.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define BBG_HD __host__ __device__ __forceinline__
#else
#define BBG_HD inline
#endif

namespace bbg {
namespace inv {

// r = a - b, returns 1 on borrow
BBG_HD uint32_t sub256(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8])
{
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t t = (uint64_t)a[i] - b[i] - borrow;
        r[i] = (uint32_t)t;
        borrow = (t >> 32) & 1;
    }
    return (uint32_t)borrow;
}
BBG_HD void add256(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8])
{
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t t = (uint64_t)a[i] + b[i] + carry;
        r[i] = (uint32_t)t;
        carry = t >> 32;
    }
}

// One lane's state.  step() is branch-free; a lane whose v is already 0 is left untouched.
struct Kaliski {
    uint32_t u[8], v[8], r[8], s[8];
    uint32_t k;
    uint32_t neg; // 1 when sigma == -1

    BBG_HD void init(const uint32_t (&a)[8], const uint32_t (&p)[8])
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            u[i] = p[i];
            v[i] = a[i];
            r[i] = 0;
            s[i] = 0;
        }
        s[0] = 1;
        k = 0;
        neg = 0;
    }
    BBG_HD bool alive() const
    {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= v[i];
        return o != 0;
    }
    BBG_HD void step()
    {
        const uint32_t live = alive() ? 1u : 0u;
        const uint32_t vo = v[0] & 1u; // 0 for a finished lane (v == 0)
        uint32_t t[8];
        const uint32_t b = sub256(t, v, u);   // t = v - u, b = (v < u)
        const uint32_t swap = vo & b;
        // |v - u|: conditional negate
        const uint32_t m = 0u - b;
        {
            uint64_t c = b;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                c += (uint64_t)(t[i] ^ m);
                t[i] = (uint32_t)c;
                c >>= 32;
            }
        }
        uint32_t sum[8];
        add256(sum, r, s);
        uint32_t nv[8], nr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            nv[i] = vo ? t[i] : v[i];      // v - u (or u - v after the swap)
            nr[i] = swap ? s[i] : r[i];
            u[i] = swap ? v[i] : u[i];     // min(u, v) when v is odd
            s[i] = vo ? sum[i] : s[i];     // s + r is symmetric under the swap
        }
#pragma unroll
        for (int i = 0; i < 7; ++i) v[i] = (nv[i] >> 1) | (nv[i + 1] << 31);
        v[7] = nv[7] >> 1;
        // r <<= live
#pragma unroll
        for (int i = 7; i > 0; --i) r[i] = live ? ((nr[i] << 1) | (nr[i - 1] >> 31)) : nr[i];
        r[0] = live ? (nr[0] << 1) : nr[0];
        neg ^= swap;
        k += live;
    }
    // a^-1 2^k mod p in [0, p)
    BBG_HD void finish(uint32_t (&out)[8], const uint32_t (&p)[8]) const
    {
        uint32_t t[8], rr[8];
        const uint32_t b = sub256(t, r, p); // r < 2p: one conditional subtraction
#pragma unroll
        for (int i = 0; i < 8; ++i) rr[i] = b ? r[i] : t[i];
        uint32_t pm[8];
        sub256(pm, p, rr); // p - r  (r == 0 cannot happen for invertible a)
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = neg ? rr[i] : pm[i];
    }
};

static constexpr uint32_t K_MIN = 254, K_MAX = 508, FIX_ENTRIES = K_MAX - K_MIN + 1;

// Host: table[k - K_MIN] = R^3 2^-k mod p (canonical), from r2 = R^2 mod p = C_256.
inline void inv_fix_table(const uint32_t (&p)[8], const uint32_t (&r2)[8], uint32_t* table /* FIX_ENTRIES * 8 */)
{
    auto put = [&](uint32_t k, const uint32_t (&x)[8]) {
        for (int i = 0; i < 8; ++i) table[(size_t)(k - K_MIN) * 8 + i] = x[i];
    };
    uint32_t cur[8];
    for (int i = 0; i < 8; ++i) cur[i] = r2[i];
    put(256, cur);
    for (uint32_t k = 257; k <= K_MAX; ++k) { // halve mod p
        uint32_t t[8];
        uint32_t top = 0;
        if (cur[0] & 1u) {
            uint64_t c = 0;
            for (int i = 0; i < 8; ++i) {
                c += (uint64_t)cur[i] + p[i];
                t[i] = (uint32_t)c;
                c >>= 32;
            }
            top = (uint32_t)c;
        } else {
            for (int i = 0; i < 8; ++i) t[i] = cur[i];
        }
        for (int i = 0; i < 7; ++i) cur[i] = (t[i] >> 1) | (t[i + 1] << 31);
        cur[7] = (t[7] >> 1) | (top << 31);
        put(k, cur);
    }
    for (int i = 0; i < 8; ++i) cur[i] = r2[i];
    for (uint32_t k = 255; k >= K_MIN; --k) { // double mod p
        uint32_t t[8], d[8];
        for (int i = 7; i > 0; --i) t[i] = (cur[i] << 1) | (cur[i - 1] >> 31);
        t[0] = cur[0] << 1; // cur < p < 2^254: no overflow
        const uint32_t b = sub256(d, t, p);
        for (int i = 0; i < 8; ++i) cur[i] = b ? t[i] : d[i];
        put(k, cur);
    }
}

} // namespace inv
} // namespace bbg
