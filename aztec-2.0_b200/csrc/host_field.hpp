// host_field.hpp -- a few scalar fr operations on the HOST, used only to derive launch parameters
// (roots of unity, 1/n, coset generator powers).  Montgomery form, R = 2^256, canonical outputs.
// Constants: bb/ecc/curves/bn254/fr.hpp:12-20,42.
#pragma once
#include <cstdint>
#include <cstring>

namespace bbg {
namespace hf {

typedef unsigned __int128 u128;

struct Fr {
    uint64_t d[4];
};

static const uint64_t MOD[4] = { 0x43E1F593F0000001ULL, 0x2833E84879B97091ULL, 0xB85045B68181585DULL, 0x30644E72E131A029ULL };
static const uint64_t R2[4] = { 0x1BB8E645AE216DA7ULL, 0x53FE3AB1E35C59E3ULL, 0x8C49833D53BB8085ULL, 0x0216D0B17F4E44A5ULL };
static const uint64_t NINV = 0xc2e1f593efffffffULL;

inline bool geq_mod(const uint64_t* a)
{
    for (int i = 3; i >= 0; --i) {
        if (a[i] != MOD[i]) return a[i] > MOD[i];
    }
    return true;
}
inline void sub_mod(uint64_t* a)
{
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a[i] - MOD[i] - borrow;
        a[i] = (uint64_t)t;
        borrow = (uint64_t)(t >> 64) & 1;
    }
}
// canonical Montgomery product
inline Fr mul(const Fr& a, const Fr& b)
{
    uint64_t t[6] = { 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a.d[i] * b.d[j] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * NINV;
        c = (u128)m * MOD[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * MOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr r;
    memcpy(r.d, t, 32);
    while (geq_mod(r.d)) sub_mod(r.d);
    return r;
}
inline Fr sqr(const Fr& a) { return mul(a, a); }
inline Fr reduce(const Fr& a)
{
    Fr r = a;
    while (geq_mod(r.d)) sub_mod(r.d);
    return r;
}
inline Fr from_u64(uint64_t v)
{
    Fr a = { { v, 0, 0, 0 } }, r2;
    memcpy(r2.d, R2, 32);
    return mul(a, r2);
}
inline Fr one() { return from_u64(1); }
inline Fr zero()
{
    Fr r = { { 0, 0, 0, 0 } };
    return r;
}
// canonical a + b and a - b (inputs any representative below 2^256 - r)
inline Fr add(const Fr& a, const Fr& b)
{
    Fr x = reduce(a), y = reduce(b), r;
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)x.d[i] + y.d[i];
        r.d[i] = (uint64_t)c;
        c >>= 64;
    }
    while (geq_mod(r.d)) sub_mod(r.d); // x + y < 2r < 2^255: no carry out
    return r;
}
inline Fr sub(const Fr& a, const Fr& b)
{
    Fr x = reduce(a), y = reduce(b), r;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)x.d[i] - y.d[i] - borrow;
        r.d[i] = (uint64_t)t;
        borrow = (uint64_t)(t >> 64) & 1;
    }
    if (borrow) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)r.d[i] + MOD[i];
            r.d[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
inline Fr pow(const Fr& a, const uint64_t e[4])
{
    Fr acc = one();
    for (int i = 255; i >= 0; --i) {
        acc = sqr(acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
    }
    return acc;
}
inline Fr invert(const Fr& a)
{
    uint64_t e[4] = { MOD[0] - 2, MOD[1], MOD[2], MOD[3] };
    return pow(a, e);
}
inline Fr load(const void* p)
{
    Fr r;
    memcpy(r.d, p, 32);
    return r;
}

} // namespace hf
} // namespace bbg
