// g1.cuh -- BN254 G1 (y^2 = x^3 + 3) group law for the MSM kernels.
//
// Storage formats are barretenberg's (bb/ecc/groups/affine_element.hpp:7-72, element.hpp:28-126):
//   affine   64 B {x, y}     Montgomery fq, point at infinity <=> bit 63 of x.data[3]  (bit 31 of limb 7)
//   Jacobian 96 B {x, y, z}  same infinity flag on x
// Internally the kernels accumulate in extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2; infinity <=> ZZ == 0) because the mixed addition costs 8M + 2S there against
// 7M + 4S for the reference's Jacobian madd-2007-bl (element_impl.hpp:243-330), with no field
// inversions.  The group element computed is identical; only the projective representative differs,
// which is why results are compared through affine_element::to_buffer().
//
// All formulas handle the exceptional cases the reference handles (element_impl.hpp:263-271,
// 397-405): P + P falls through to doubling, P + (-P) gives infinity, infinity is an identity.
#pragma once
#include "field.cuh"

namespace bbg {

using fq = Fe<FqParams>;

struct alignas(16) affine_t {
    fq x, y;
};
struct alignas(16) jac_t {
    fq x, y, z;
};
struct alignas(16) xyzz_t {
    fq x, y, zz, zzz;
};

static constexpr uint32_t INF_BIT = 0x80000000u; // bit 255 of x

__device__ __forceinline__ bool affine_is_inf(const affine_t& p) { return (p.x.l[7] & INF_BIT) != 0; }
__device__ __forceinline__ bool xyzz_is_inf(const xyzz_t& p) { return fe_is_zero(p.zz); }

__device__ __forceinline__ xyzz_t xyzz_infinity()
{
    xyzz_t r;
    r.x = fe_zero<FqParams>();
    r.y = fe_zero<FqParams>();
    r.zz = fe_zero<FqParams>();
    r.zzz = fe_zero<FqParams>();
    return r;
}
__device__ __forceinline__ xyzz_t xyzz_from_affine(const affine_t& p)
{
    xyzz_t r;
    if (affine_is_inf(p)) {
        return xyzz_infinity();
    }
    r.x = p.x;
    r.y = p.y;
    r.zz = fe_one<FqParams>();
    r.zzz = fe_one<FqParams>();
    return r;
}
__device__ __forceinline__ affine_t affine_load(const affine_t* p)
{
    affine_t r;
    r.x = fe_load_nc<FqParams>(&p->x);
    r.y = fe_load_nc<FqParams>(&p->y);
    return r;
}
__device__ __forceinline__ affine_t affine_neg(const affine_t& p)
{
    affine_t r;
    r.x = p.x;
    r.y = fe_neg(p.y);
    return r;
}
__device__ __forceinline__ xyzz_t xyzz_load(const xyzz_t* p)
{
    xyzz_t r;
    r.x = fe_load<FqParams>(&p->x);
    r.y = fe_load<FqParams>(&p->y);
    r.zz = fe_load<FqParams>(&p->zz);
    r.zzz = fe_load<FqParams>(&p->zzz);
    return r;
}
__device__ __forceinline__ void xyzz_store(xyzz_t* p, const xyzz_t& v)
{
    fe_store(&p->x, v.x);
    fe_store(&p->y, v.y);
    fe_store(&p->zz, v.zz);
    fe_store(&p->zzz, v.zzz);
}

// 2 * (x, y) for an affine point, result in XYZZ ("mdbl-2008-s-1", a = 0): 2M + 5S... here 3M + 4S
__device__ __forceinline__ xyzz_t xyzz_dbl_affine(const affine_t& p)
{
    xyzz_t r;
    fq u = fe_dbl(p.y);           // U = 2 y
    fq v = fe_sqr(u);             // V = U^2
    fq w = fe_mul(u, v);          // W = U V
    fq s = fe_mul(p.x, v);        // S = x V
    fq xx = fe_sqr(p.x);
    fq m = fe_add(fe_dbl(xx), xx); // M = 3 x^2
    r.x = fe_sub(fe_sqr(m), fe_dbl(s));
    r.y = fe_sub(fe_mul(m, fe_sub(s, r.x)), fe_mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}
// 2 * P in XYZZ ("dbl-2008-s-1", a = 0)
__device__ __forceinline__ xyzz_t xyzz_dbl(const xyzz_t& p)
{
    if (xyzz_is_inf(p)) {
        return p;
    }
    xyzz_t r;
    fq u = fe_dbl(p.y);
    fq v = fe_sqr(u);
    fq w = fe_mul(u, v);
    fq s = fe_mul(p.x, v);
    fq xx = fe_sqr(p.x);
    fq m = fe_add(fe_dbl(xx), xx);
    r.x = fe_sub(fe_sqr(m), fe_dbl(s));
    r.y = fe_sub(fe_mul(m, fe_sub(s, r.x)), fe_mul(w, p.y));
    r.zz = fe_mul(v, p.zz);
    r.zzz = fe_mul(w, p.zzz);
    return r;
}

// acc += (x2, y2)  ("madd-2008-s": 8M + 2S).  b must not be the point at infinity; acc may be.
__device__ __forceinline__ void xyzz_madd(xyzz_t& acc, const affine_t& b)
{
    if (xyzz_is_inf(acc)) {
        acc.x = b.x;
        acc.y = b.y;
        acc.zz = fe_one<FqParams>();
        acc.zzz = fe_one<FqParams>();
        return;
    }
    fq u2 = fe_mul(b.x, acc.zz);
    fq s2 = fe_mul(b.y, acc.zzz);
    fq p = fe_sub(u2, acc.x);
    fq r = fe_sub(s2, acc.y);
    if (__builtin_expect(fe_is_zero(p), 0)) {
        if (fe_is_zero(r)) {
            acc = xyzz_dbl_affine(b); // P + P (reference: element_impl.hpp:263-266)
        } else {
            acc = xyzz_infinity();    // P + (-P) (:267-270)
        }
        return;
    }
    fq pp = fe_sqr(p);
    fq ppp = fe_mul(p, pp);
    fq q = fe_mul(acc.x, pp);
    fq x3 = fe_sub(fe_sub(fe_sqr(r), ppp), fe_dbl(q));
    fq y3 = fe_sub(fe_mul(r, fe_sub(q, x3)), fe_mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fe_mul(acc.zz, pp);
    acc.zzz = fe_mul(acc.zzz, ppp);
}

// a += b, both XYZZ ("add-2008-s": 12M + 2S), all exceptional cases handled.
// Everything in this header is force-inlined: a __noinline__ helper called from a divergent branch was
// observed (compute-sanitizer, sm_100a, CUDA 12.9) to clobber a uniform register that the other
// lanes of the warp still needed, so the kernels keep ONE inlined add site per loop instead of calls.
__device__ __forceinline__ void xyzz_add(xyzz_t& a, const xyzz_t& b)
{
    if (xyzz_is_inf(b)) {
        return;
    }
    if (xyzz_is_inf(a)) {
        a = b;
        return;
    }
    fq u1 = fe_mul(a.x, b.zz);
    fq u2 = fe_mul(b.x, a.zz);
    fq s1 = fe_mul(a.y, b.zzz);
    fq s2 = fe_mul(b.y, a.zzz);
    fq p = fe_sub(u2, u1);
    fq r = fe_sub(s2, s1);
    if (__builtin_expect(fe_is_zero(p), 0)) {
        if (fe_is_zero(r)) {
            a = xyzz_dbl(a);
        } else {
            a = xyzz_infinity();
        }
        return;
    }
    fq pp = fe_sqr(p);
    fq ppp = fe_mul(p, pp);
    fq q = fe_mul(u1, pp);
    fq x3 = fe_sub(fe_sub(fe_sqr(r), ppp), fe_dbl(q));
    fq y3 = fe_sub(fe_mul(r, fe_sub(q, x3)), fe_mul(s1, ppp));
    a.x = x3;
    a.y = y3;
    a.zz = fe_mul(fe_mul(a.zz, b.zz), pp);
    a.zzz = fe_mul(fe_mul(a.zzz, b.zzz), ppp);
}

// lane-to-lane move of a whole point (warp-shuffle g1 reductions)
__device__ __forceinline__ xyzz_t xyzz_shfl_down(const xyzz_t& p, int delta)
{
    xyzz_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, p.x.l[i], delta);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, p.y.l[i], delta);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, p.zz.l[i], delta);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, p.zzz.l[i], delta);
    }
    return r;
}
// c ? a : b, limb-wise (keeps both operands in registers; no local-memory indexing)
__device__ __forceinline__ xyzz_t xyzz_select(bool c, const xyzz_t& a, const xyzz_t& b)
{
    xyzz_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.x.l[i] = c ? a.x.l[i] : b.x.l[i];
        r.y.l[i] = c ? a.y.l[i] : b.y.l[i];
        r.zz.l[i] = c ? a.zz.l[i] : b.zz.l[i];
        r.zzz.l[i] = c ? a.zzz.l[i] : b.zzz.l[i];
    }
    return r;
}

// a^(p-2) (field_impl.hpp:323-329 invert = pow(modulus - 2)); plain square-and-multiply, loop kept rolled
template <class F> __device__ __forceinline__ Fe<F> fe_inv(const Fe<F>& a)
{
    Fe<F> acc = fe_one<F>();
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = F::P(i);
    e[0] -= 2; // both moduli end in ...47 / ...01: no borrow
#pragma unroll 1
    for (int i = 253; i >= 0; --i) {
        acc = fe_sqr(acc);
        uint32_t limb = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) limb = (i >> 5) == k ? e[k] : limb;
        if ((limb >> (i & 31)) & 1) {
            acc = fe_mul(acc, a);
        }
    }
    return acc;
}

// XYZZ -> affine (x = X / ZZ, y = Y / ZZZ) with one field inversion; infinity keeps barretenberg's flag
__device__ __forceinline__ affine_t xyzz_to_affine(const xyzz_t& p)
{
    affine_t r;
    if (xyzz_is_inf(p)) {
        r.x = fe_zero<FqParams>();
        r.y = fe_zero<FqParams>();
        r.x.l[7] |= INF_BIT;
        return r;
    }
    fq inv = fe_inv(fe_mul(p.zz, p.zzz));      // 1 / (ZZ * ZZZ)
    r.x = fe_reduce_once(fe_mul(p.x, fe_mul(inv, p.zzz)));
    r.y = fe_reduce_once(fe_mul(p.y, fe_mul(inv, p.zz)));
    return r;
}

// XYZZ -> the reference's 96-byte Jacobian element: with Z := ZZZ we have Z^2 = ZZ^3, Z^3 = ZZZ^3,
// hence X_j = X * ZZ^2, Y_j = Y * ZZZ^2.  Infinity is encoded exactly like g1::element::self_set_infinity
// applied to g1::one (element_impl.hpp:497-516): x = 1 with bit 255 set, y = 2, z = 1 (Montgomery).
__device__ __forceinline__ jac_t xyzz_to_jacobian(const xyzz_t& p)
{
    jac_t r;
    if (xyzz_is_inf(p)) {
        r.x = fe_one<FqParams>();
        r.x.l[7] |= INF_BIT;
        r.y = fe_dbl(fe_one<FqParams>());
        r.z = fe_one<FqParams>();
        return r;
    }
    r.x = fe_mul(p.x, fe_sqr(p.zz));
    r.y = fe_mul(p.y, fe_sqr(p.zzz));
    r.z = p.zzz;
    return r;
}
// the reference's Jacobian element -> XYZZ (ZZ = z^2, ZZZ = z^3)
__device__ __forceinline__ xyzz_t xyzz_from_jacobian(const jac_t& p)
{
    if (p.x.l[7] & INF_BIT) {
        return xyzz_infinity();
    }
    xyzz_t r;
    r.x = p.x;
    r.y = p.y;
    r.zz = fe_sqr(p.z);
    r.zzz = fe_mul(r.zz, p.z);
    return r;
}

} // namespace bbg
