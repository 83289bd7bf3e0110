/* oracle/bn254_oracle.c -- TEST INFRASTRUCTURE ONLY (see bn254_oracle.h).
 *
 * CPU restatement, in plain C, of the algorithms on barretenberg's PLONK-prover hot path.
 * "bb/" below = /root/reference/barretenberg/src/aztec/.  Every function cites the reference
 * lines it follows.  Written for clarity, not speed: it is the checker for the CUDA path.
 */
#include "bn254_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------
 * Field parameters (bb/ecc/curves/bn254/fq.hpp:9-59, fr.hpp:10-60)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t p[4];      /* modulus */
    uint64_t r2[4];     /* R^2 mod p */
    uint64_t ninv;      /* -p^-1 mod 2^64 ("r_inv") */
} field_params;

static const field_params FQ = {
    { 0x3C208C16D87CFD47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL },
    { 0xF32CFC5B538AFA89ULL, 0xB5E71911D44501FBULL, 0x47AB1EFF0A417FF6ULL, 0x06D89F71CAB8351FULL },
    0x87d20782e4866389ULL
};
static const field_params FR = {
    { 0x43E1F593F0000001ULL, 0x2833E84879B97091ULL, 0xB85045B68181585DULL, 0x30644E72E131A029ULL },
    { 0x1BB8E645AE216DA7ULL, 0x53FE3AB1E35C59E3ULL, 0x8C49833D53BB8085ULL, 0x0216D0B17F4E44A5ULL },
    0xc2e1f593efffffffULL
};
/* fr: a primitive 2^28-th root of unity, Montgomery form (fr.hpp:27-30) */
static const orc_fe FR_ROOT_28 = { { 0x636e735580d13d9cULL, 0xa22bf3742445ffd6ULL, 0x56452ac01eb203d8ULL,
                                     0x1860ef942963f9e7ULL } };
/* fq: cube root of unity beta, Montgomery form (fq.hpp:21-24) */
static const orc_fe FQ_BETA = { { 0x71930c11d782e155ULL, 0xa6bb947cffbe3323ULL, 0xaa303344d4741444ULL,
                                  0x2c3b3f0d26594943ULL } };
/* g1 generator (1, 2) and curve constant b = 3 in Montgomery form (g1.hpp:13-16) */
static const orc_fe G1_ONE_Y = { { 0xa6ba871b8b1e1b3aULL, 0x14f1d651eb8e167bULL, 0xccdd46def0f28c58ULL,
                                   0x1c14ef83340fbe5eULL } };
static const orc_fe G1_B = { { 0x7a17caa950ad28d7ULL, 0x1f6ac17ae15521b9ULL, 0x334bea4e696bd284ULL,
                               0x2a1f6744ce179d8eULL } };

static inline const field_params* params(int field) { return field == ORC_FQ ? &FQ : &FR; }

/* ------------------------------------------------------------------------------------------
 * 256-bit helpers
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t add4(uint64_t* r, const uint64_t* a, const uint64_t* b)
{
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline uint64_t sub4(uint64_t* r, const uint64_t* a, const uint64_t* b)
{
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)t;
        borrow = (uint64_t)(t >> 64) & 1;
    }
    return borrow;
}
static inline int geq4(const uint64_t* a, const uint64_t* b)
{
    for (int i = 3; i >= 0; --i) {
        if (a[i] != b[i]) {
            return a[i] > b[i];
        }
    }
    return 1;
}
static inline void twice4(uint64_t* r, const uint64_t* a)
{
    r[3] = (a[3] << 1) | (a[2] >> 63);
    r[2] = (a[2] << 1) | (a[1] >> 63);
    r[1] = (a[1] << 1) | (a[0] >> 63);
    r[0] = a[0] << 1;
}

/* ------------------------------------------------------------------------------------------
 * Field arithmetic with the reference's coarse [0, 2p) representation
 * (bb/ecc/fields/field_impl.hpp:34-198, field_impl_generic.hpp:171-272, 392-442)
 * ---------------------------------------------------------------------------------------- */

/* reduce_once: conditional subtraction of p (field_impl_generic.hpp:171-194) */
static void f_reduce_once(const field_params* F, orc_fe* r, const orc_fe* a)
{
    uint64_t t[4];
    if (geq4(a->d, F->p)) {
        sub4(t, a->d, F->p);
        memcpy(r->d, t, 32);
    } else if (r != a) {
        *r = *a;
    }
}

/* Montgomery product a*b*R^-1 with NO final subtraction: inputs < 2p => output < 2p
 * (field_impl_generic.hpp:392-442; same integer (a*b + m*p)/2^256 as the x86 asm path) */
static void f_mul(const field_params* F, orc_fe* r, const orc_fe* a, const orc_fe* b)
{
    uint64_t t[6] = { 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a->d[i] * b->d[j] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);

        uint64_t m = t[0] * F->ninv;
        c = (u128)m * F->p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * F->p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    memcpy(r->d, t, 32);
}
static void f_sqr(const field_params* F, orc_fe* r, const orc_fe* a) { f_mul(F, r, a, a); }

/* a + b, then -2p if the sum is >= 2p (field_impl_generic.hpp:213-234) */
static void f_add(const field_params* F, orc_fe* r, const orc_fe* a, const orc_fe* b)
{
    uint64_t s[4], tp[4], t[4];
    add4(s, a->d, b->d); /* < 4p < 2^256: no carry out */
    twice4(tp, F->p);
    if (geq4(s, tp)) {
        sub4(t, s, tp);
        memcpy(r->d, t, 32);
    } else {
        memcpy(r->d, s, 32);
    }
}
/* a - b, then +2p on borrow (subtract_coarse, field_impl_generic.hpp:254-272) */
static void f_sub(const field_params* F, orc_fe* r, const orc_fe* a, const orc_fe* b)
{
    uint64_t s[4], tp[4];
    uint64_t borrow = sub4(s, a->d, b->d);
    if (borrow) {
        twice4(tp, F->p);
        add4(s, s, tp);
    }
    memcpy(r->d, s, 32);
}
/* -a = 2p - a (field_impl.hpp:148-157) */
static void f_neg(const field_params* F, orc_fe* r, const orc_fe* a)
{
    uint64_t tp[4], s[4];
    twice4(tp, F->p);
    sub4(s, tp, a->d);
    memcpy(r->d, s, 32);
}
/* zero or p both count as zero (field_impl.hpp:490-494) */
static int f_is_zero(const field_params* F, const orc_fe* a)
{
    if ((a->d[0] | a->d[1] | a->d[2] | a->d[3]) == 0) {
        return 1;
    }
    return memcmp(a->d, F->p, 32) == 0;
}
/* to_montgomery_form: three reduce_once, * R^2, reduce_once (field_impl.hpp:234-244) */
static void f_to_mont(const field_params* F, orc_fe* r, const orc_fe* a)
{
    orc_fe t = *a, r2;
    f_reduce_once(F, &t, &t);
    f_reduce_once(F, &t, &t);
    f_reduce_once(F, &t, &t);
    memcpy(r2.d, F->r2, 32);
    f_mul(F, &t, &t, &r2);
    f_reduce_once(F, r, &t);
}
/* from_montgomery_form: * 1 then reduce_once => canonical (field_impl.hpp:246-250) */
static void f_from_mont(const field_params* F, orc_fe* r, const orc_fe* a)
{
    orc_fe one = { { 1, 0, 0, 0 } }, t;
    f_mul(F, &t, a, &one);
    f_reduce_once(F, r, &t);
}
static void f_one(const field_params* F, orc_fe* r)
{
    orc_fe one = { { 1, 0, 0, 0 } };
    f_to_mont(F, r, &one);
}
static int f_eq(const field_params* F, const orc_fe* a, const orc_fe* b)
{
    orc_fe x, y;
    f_reduce_once(F, &x, a);
    f_reduce_once(F, &y, b);
    return memcmp(x.d, y.d, 32) == 0;
}
/* square-and-multiply, MSB first (field_impl.hpp:296-316) */
static void f_pow256(const field_params* F, orc_fe* r, const orc_fe* a, const uint64_t e[4])
{
    int msb = -1;
    for (int i = 255; i >= 0; --i) {
        if ((e[i >> 6] >> (i & 63)) & 1) {
            msb = i;
            break;
        }
    }
    if (f_is_zero(F, a)) {
        memset(r, 0, sizeof(*r));
        return;
    }
    if (msb < 0) {
        f_one(F, r);
        return;
    }
    orc_fe acc = *a;
    for (int i = msb - 1; i >= 0; --i) {
        f_sqr(F, &acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) {
            f_mul(F, &acc, &acc, a);
        }
    }
    *r = acc;
}
/* invert = a^(p-2) (field_impl.hpp:323-329); zero maps to zero here (reference throws) */
static void f_invert(const field_params* F, orc_fe* r, const orc_fe* a)
{
    uint64_t e[4], two[4] = { 2, 0, 0, 0 };
    sub4(e, F->p, two);
    f_pow256(F, r, a, e);
}

void orc_field_op(int field, int op, const orc_fe* a, const orc_fe* b, orc_fe* out)
{
    const field_params* F = params(field);
    orc_fe r;
    switch (op) {
    case 0: f_mul(F, &r, a, b); break;
    case 1: f_add(F, &r, a, b); break;
    case 2: f_sub(F, &r, a, b); break;
    case 3: f_sqr(F, &r, a); break;
    case 4: f_to_mont(F, &r, a); break;
    case 5: f_from_mont(F, &r, a); break;
    case 6: f_invert(F, &r, a); break;
    case 7: f_reduce_once(F, &r, a); break;
    case 8: f_neg(F, &r, a); break;
    default: memset(&r, 0, sizeof(r));
    }
    *out = r;
}
void orc_fr_pow(const orc_fe* a, uint64_t e, orc_fe* out)
{
    uint64_t ee[4] = { e, 0, 0, 0 };
    f_pow256(&FR, out, a, ee);
}
/* get_root_of_unity(k): the 2^28-th root squared (28 - k) times (field_impl.hpp:496-503) */
void orc_fr_root_of_unity(unsigned log2n, orc_fe* out)
{
    orc_fe r = FR_ROOT_28;
    for (unsigned i = 28; i > log2n; --i) {
        f_sqr(&FR, &r, &r);
    }
    *out = r;
}
/* coset_generator(0) = the multiplicative generator 5 (fr.hpp:44-59, evaluation_domain.cpp:69-70) */
void orc_fr_coset_generator(orc_fe* out)
{
    /* the reference stores this constant as a coarse (>= r) representative of 5 * R mod r */
    static const orc_fe g = { { 0x5eef048d8fffffe7ULL, 0x12ee50ec1ce401d0ULL, 0x0029312d5a5e5ee7ULL,
                                0x463456c802275bedULL } };
    *out = g;
}
void orc_fq_beta(orc_fe* out) { *out = FQ_BETA; }

/* ------------------------------------------------------------------------------------------
 * g1 (bb/ecc/groups/element_impl.hpp, affine_element_impl.hpp); y^2 = x^3 + 3
 * ---------------------------------------------------------------------------------------- */
#define MSB63 0x8000000000000000ULL
static inline int aff_is_inf(const orc_affine* a) { return (a->x.d[3] & MSB63) != 0; }
static inline int jac_is_inf(const orc_jac* a) { return (a->x.d[3] & MSB63) != 0; }
/* element::self_set_infinity: set bit 255 of x (element_impl.hpp:497-516) */
static inline void jac_set_inf(orc_jac* a) { a->x.d[3] |= MSB63; }

void orc_g1_one(orc_affine* out)
{
    f_one(&FQ, &out->x);
    out->y = G1_ONE_Y;
}
void orc_g1_set_infinity(orc_jac* out)
{
    f_one(&FQ, &out->x);
    out->y = G1_ONE_Y;
    f_one(&FQ, &out->z);
    jac_set_inf(out);
}

/* dbl-2009-l doubling for a = 0 (element_impl.hpp:70-139) */
static void jac_dbl(orc_jac* r, const orc_jac* a)
{
    const field_params* F = &FQ;
    if (jac_is_inf(a)) {
        *r = *a;
        jac_set_inf(r);
        return;
    }
    orc_fe xx, yy, yyyy, s, m, t, x3, y3, z3;
    f_sqr(F, &xx, &a->x);
    f_sqr(F, &yy, &a->y);
    f_sqr(F, &yyyy, &yy);
    f_add(F, &s, &yy, &a->x);     /* (x + yy)^2 - xx - yyyy = 2 x yy */
    f_sqr(F, &s, &s);
    f_add(F, &t, &xx, &yyyy);
    f_sub(F, &s, &s, &t);
    f_add(F, &s, &s, &s);         /* S = 4 x yy */
    f_add(F, &m, &xx, &xx);
    f_add(F, &m, &m, &xx);        /* M = 3 xx */
    f_add(F, &z3, &a->z, &a->z);
    f_mul(F, &z3, &z3, &a->y);    /* z3 = 2 y z */
    f_add(F, &t, &s, &s);
    f_sqr(F, &x3, &m);
    f_sub(F, &x3, &x3, &t);       /* x3 = M^2 - 2S */
    f_add(F, &yyyy, &yyyy, &yyyy);
    f_add(F, &yyyy, &yyyy, &yyyy);
    f_add(F, &yyyy, &yyyy, &yyyy); /* 8 yyyy */
    f_sub(F, &y3, &s, &x3);
    f_mul(F, &y3, &y3, &m);
    f_sub(F, &y3, &y3, &yyyy);
    r->x = x3;
    r->y = y3;
    r->z = z3;
}

/* madd-2007-bl mixed addition (element_impl.hpp:243-330) */
static void jac_mixed_add(orc_jac* r, const orc_jac* a, const orc_affine* b)
{
    const field_params* F = &FQ;
    if (jac_is_inf(a) || aff_is_inf(b)) {
        if (jac_is_inf(a)) {
            /* note: the reference copies (x, y, 1) even when b is infinity: the flag bit rides along in x */
            r->x = b->x;
            r->y = b->y;
            f_one(F, &r->z);
        } else {
            *r = *a;
        }
        return;
    }
    orc_fe z1z1, h, rr, t, hh, i4, j, v, x3, y3, z3;
    f_sqr(F, &z1z1, &a->z);
    f_mul(F, &h, &b->x, &z1z1);
    f_sub(F, &h, &h, &a->x);          /* H = x2 z1^2 - x1 */
    f_mul(F, &rr, &a->z, &z1z1);
    f_mul(F, &rr, &rr, &b->y);
    f_sub(F, &rr, &rr, &a->y);        /* y2 z1^3 - y1 */
    if (f_is_zero(F, &h)) {
        if (f_is_zero(F, &rr)) {
            jac_dbl(r, a);
        } else {
            *r = *a;
            jac_set_inf(r);
        }
        return;
    }
    f_add(F, &rr, &rr, &rr);          /* R = 2 (y2 z1^3 - y1) */
    f_sqr(F, &hh, &h);
    f_add(F, &z3, &a->z, &h);
    f_sqr(F, &z3, &z3);
    f_add(F, &t, &z1z1, &hh);
    f_sub(F, &z3, &z3, &t);           /* z3 = (z1 + H)^2 - z1z1 - HH */
    f_add(F, &i4, &hh, &hh);
    f_add(F, &i4, &i4, &i4);          /* I = 4 HH */
    f_mul(F, &j, &h, &i4);            /* J = H I */
    f_mul(F, &v, &i4, &a->x);         /* V = x1 I */
    f_add(F, &t, &v, &v);
    f_add(F, &t, &t, &j);
    f_sqr(F, &x3, &rr);
    f_sub(F, &x3, &x3, &t);           /* x3 = R^2 - J - 2V */
    f_sub(F, &y3, &v, &x3);
    f_mul(F, &y3, &y3, &rr);
    f_mul(F, &t, &j, &a->y);
    f_add(F, &t, &t, &t);
    f_sub(F, &y3, &y3, &t);           /* y3 = R (V - x3) - 2 y1 J */
    r->x = x3;
    r->y = y3;
    r->z = z3;
}

/* add-2007-bl full addition (element_impl.hpp:354-441) */
static void jac_add(orc_jac* r, const orc_jac* a, const orc_jac* b)
{
    const field_params* F = &FQ;
    int ai = jac_is_inf(a), bi = jac_is_inf(b);
    if (ai || bi) {
        if (ai && !bi) {
            *r = *b;
        } else if (bi && !ai) {
            *r = *a;
        } else {
            *r = *a;
            jac_set_inf(r);
        }
        return;
    }
    orc_fe z1z1, z2z2, u1, u2, s1, s2, h, f, i, j, t, x3, y3, z3;
    f_sqr(F, &z1z1, &a->z);
    f_sqr(F, &z2z2, &b->z);
    f_mul(F, &s2, &z1z1, &a->z);
    f_mul(F, &u2, &z1z1, &b->x);
    f_mul(F, &s2, &s2, &b->y);
    f_mul(F, &u1, &z2z2, &a->x);
    f_mul(F, &s1, &z2z2, &b->z);
    f_mul(F, &s1, &s1, &a->y);
    f_sub(F, &f, &s2, &s1);
    f_sub(F, &h, &u2, &u1);
    if (f_is_zero(F, &h)) {
        if (f_is_zero(F, &f)) {
            jac_dbl(r, a);
        } else {
            *r = *a;
            jac_set_inf(r);
        }
        return;
    }
    f_add(F, &f, &f, &f);
    f_add(F, &i, &h, &h);
    f_sqr(F, &i, &i);
    f_mul(F, &j, &h, &i);
    f_mul(F, &u1, &u1, &i);
    f_add(F, &u2, &u1, &u1);
    f_add(F, &u2, &u2, &j);
    f_sqr(F, &x3, &f);
    f_sub(F, &x3, &x3, &u2);
    f_mul(F, &j, &j, &s1);
    f_add(F, &j, &j, &j);
    f_sub(F, &y3, &u1, &x3);
    f_mul(F, &y3, &y3, &f);
    f_sub(F, &y3, &y3, &j);
    f_add(F, &z3, &a->z, &b->z);
    f_add(F, &t, &z1z1, &z2z2);
    f_sqr(F, &z3, &z3);
    f_sub(F, &z3, &z3, &t);
    f_mul(F, &z3, &z3, &h);
    r->x = x3;
    r->y = y3;
    r->z = z3;
}

void orc_g1_mixed_add(const orc_jac* a, const orc_affine* b, orc_jac* out)
{
    orc_jac r;
    jac_mixed_add(&r, a, b);
    *out = r;
}
void orc_g1_add(const orc_jac* a, const orc_jac* b, orc_jac* out)
{
    orc_jac r;
    jac_add(&r, a, b);
    *out = r;
}
void orc_g1_dbl(const orc_jac* a, orc_jac* out)
{
    orc_jac r;
    jac_dbl(&r, a);
    *out = r;
}

/* Jacobian -> affine: (x / z^2, y / z^3); infinity -> (0, 0) with the flag (element_impl.hpp:51-68) */
void orc_g1_to_affine(const orc_jac* a, orc_affine* out)
{
    const field_params* F = &FQ;
    if (jac_is_inf(a)) {
        memset(out, 0, sizeof(*out));
        out->x.d[3] |= MSB63;
        return;
    }
    orc_fe zi, zzi, zzzi;
    f_invert(F, &zi, &a->z);
    f_sqr(F, &zzi, &zi);
    f_mul(F, &zzzi, &zzi, &zi);
    f_mul(F, &out->x, &a->x, &zzi);
    f_mul(F, &out->y, &a->y, &zzzi);
}

/* big-endian canonical 32 bytes (fields/field.hpp:449-457 `write`, serialize_to_buffer) */
static void fq_to_be32(const orc_fe* a, uint8_t* buf)
{
    orc_fe c;
    f_from_mont(&FQ, &c, a);
    for (int limb = 0; limb < 4; ++limb) {
        uint64_t v = c.d[3 - limb];
        for (int b = 0; b < 8; ++b) {
            buf[limb * 8 + b] = (uint8_t)(v >> (56 - 8 * b));
        }
    }
}
/* affine_element::serialize_to_buffer: y then x, infinity flag in bit 7 of byte 0 (affine_element.hpp:38-45) */
void orc_g1_affine_to_buffer(const orc_affine* a, uint8_t* buf64)
{
    fq_to_be32(&a->y, buf64);
    fq_to_be32(&a->x, buf64 + 32);
    if (aff_is_inf(a)) {
        buf64[0] |= 0x80;
    }
}
void orc_g1_jac_to_buffer(const orc_jac* a, uint8_t* buf64)
{
    orc_affine t;
    orc_g1_to_affine(a, &t);
    orc_g1_affine_to_buffer(&t, buf64);
}
int orc_g1_on_curve(const orc_affine* a)
{
    const field_params* F = &FQ;
    if (aff_is_inf(a)) {
        return 1;
    }
    orc_fe lhs, rhs;
    f_sqr(F, &rhs, &a->x);
    f_mul(F, &rhs, &rhs, &a->x);
    f_add(F, &rhs, &rhs, &G1_B);
    f_sqr(F, &lhs, &a->y);
    return f_eq(F, &lhs, &rhs);
}
/* k * P by plain double-and-add on the canonical scalar (same group element as
 * element::operator*(fr), element_impl.hpp:592-663, which uses the GLV split + wNAF) */
void orc_g1_mul(const orc_affine* p, const orc_fe* scalar_mont, orc_jac* out)
{
    orc_fe k;
    orc_jac acc;
    f_from_mont(&FR, &k, scalar_mont);
    orc_g1_set_infinity(&acc);
    if (aff_is_inf(p)) {
        *out = acc;
        return;
    }
    for (int i = 255; i >= 0; --i) {
        jac_dbl(&acc, &acc);
        if ((k.d[i >> 6] >> (i & 63)) & 1) {
            jac_mixed_add(&acc, &acc, p);
        }
    }
    *out = acc;
}
/* c_bind.cpp:40-45 */
void orc_g1_sum(const orc_jac* pts, size_t n, orc_jac* out)
{
    orc_jac acc;
    orc_g1_set_infinity(&acc);
    for (size_t i = 0; i < n; ++i) {
        jac_add(&acc, &acc, &pts[i]);
    }
    *out = acc;
}

/* ------------------------------------------------------------------------------------------
 * SRS (bb/srs/io.cpp)
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t load_be64(const uint8_t* p)
{
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) {
        v = (v << 8) | p[i];
    }
    return v;
}
static inline uint32_t load_be32(const uint8_t* p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
/* each coordinate: 4 limbs least-significant first, each limb big-endian, non-Montgomery
 * => bswap every limb, then to_montgomery_form (io.cpp:47-67) */
void orc_read_g1_elements_from_buffer(orc_affine* elements, const uint8_t* buffer, size_t buffer_size)
{
    size_t n = buffer_size / 64;
    for (size_t i = 0; i < n; ++i) {
        orc_fe x, y;
        for (int l = 0; l < 4; ++l) {
            x.d[l] = load_be64(buffer + i * 64 + l * 8);
            y.d[l] = load_be64(buffer + i * 64 + 32 + l * 8);
        }
        f_to_mont(&FQ, &elements[i].x, &x);
        f_to_mont(&FQ, &elements[i].y, &y);
    }
}
/* monomials[0] = generator; then file points from transcript00.dat, transcript01.dat, ...
 * 28-byte manifest of 7 big-endian u32 (io.cpp:11-45, 134-162). returns 0 ok, 1 short/missing */
int orc_read_transcript_g1(orc_affine* monomials, size_t degree, const char* dir)
{
    if (degree == 0) {
        return 0;
    }
    orc_g1_one(&monomials[0]);
    size_t num_read = 1;
    for (int num = 0; num_read < degree; ++num) {
        char path[4096];
        snprintf(path, sizeof(path), "%s/transcript%02d.dat", dir, num);
        FILE* f = fopen(path, "rb");
        if (!f) {
            break;
        }
        uint8_t man[28];
        if (fread(man, 1, 28, f) != 28) {
            fclose(f);
            break;
        }
        size_t num_g1 = load_be32(man + 16); /* num_g1_points */
        size_t to_read = degree - num_read < num_g1 ? degree - num_read : num_g1;
        uint8_t* buf = (uint8_t*)malloc(to_read * 64 + 1);
        size_t got = fread(buf, 1, to_read * 64, f);
        fclose(f);
        orc_read_g1_elements_from_buffer(&monomials[num_read], buf, got);
        free(buf);
        num_read += got / 64;
        if (got != to_read * 64) {
            break;
        }
    }
    return num_read < degree ? 1 : 0;
}

/* table[2i] = P_i, table[2i+1] = (beta * x_i, -y_i); iterate backwards so table may alias points
 * (scalar_multiplication.cpp:104-112) */
void orc_generate_pippenger_point_table(const orc_affine* points, orc_affine* table, size_t n)
{
    for (size_t i = n; i-- > 0;) {
        orc_affine p = points[i];
        table[2 * i] = p;
        f_mul(&FQ, &table[2 * i + 1].x, &p.x, &FQ_BETA);
        f_neg(&FQ, &table[2 * i + 1].y, &p.y);
    }
}

/* ------------------------------------------------------------------------------------------
 * MSM.  The reference computes sum_i s_i * P_i with an endomorphism-split, signed-window,
 * affine-batched Pippenger (scalar_multiplication.cpp:188-906).  The Jacobian representative it
 * returns depends on addition order, so only the group element is comparable; the
 * oracle therefore restates the *bucket method itself* in its simplest form:
 *   for each c-bit window (MSB first): acc <<= c; drop each point in bucket[digit];
 *   acc += sum_b b * bucket[b] via the running-sum trick (scalar_multiplication.cpp:773-783).
 * ---------------------------------------------------------------------------------------- */
static unsigned window_bits_for(size_t n)
{
    unsigned c = 1;
    while (((size_t)1 << (c + 3)) < n && c < 16) {
        ++c;
    }
    return c < 4 ? 4 : c;
}
static inline unsigned get_window(const orc_fe* k, unsigned lo, unsigned c)
{
    unsigned limb = lo >> 6, off = lo & 63;
    uint64_t v = k->d[limb] >> off;
    if (off + c > 64 && limb < 3) {
        v |= k->d[limb + 1] << (64 - off);
    }
    return (unsigned)(v & (((uint64_t)1 << c) - 1));
}
void orc_pippenger(const orc_fe* scalars, const orc_affine* points, size_t n, size_t point_stride, orc_jac* out)
{
    orc_jac total;
    orc_g1_set_infinity(&total);
    if (n == 0) {
        *out = total;
        return;
    }
    const unsigned c = window_bits_for(n);
    const unsigned num_windows = (254 + c - 1) / c;
    const size_t num_buckets = ((size_t)1 << c) - 1;
    orc_fe* k = (orc_fe*)malloc(n * sizeof(orc_fe));
    for (size_t i = 0; i < n; ++i) {
        f_from_mont(&FR, &k[i], &scalars[i]); /* scalar_multiplication.cpp:224 */
    }
    orc_jac* window_sums = (orc_jac*)malloc(num_windows * sizeof(orc_jac));
#pragma omp parallel for schedule(dynamic, 1)
    for (unsigned w = 0; w < num_windows; ++w) {
        orc_jac* buckets = (orc_jac*)malloc(num_buckets * sizeof(orc_jac));
        for (size_t b = 0; b < num_buckets; ++b) {
            orc_g1_set_infinity(&buckets[b]);
        }
        for (size_t i = 0; i < n; ++i) {
            unsigned d = get_window(&k[i], w * c, c);
            if (d) {
                jac_mixed_add(&buckets[d - 1], &buckets[d - 1], &points[i * point_stride]);
            }
        }
        orc_jac running, sum;
        orc_g1_set_infinity(&running);
        orc_g1_set_infinity(&sum);
        for (size_t b = num_buckets; b-- > 0;) {
            jac_add(&running, &running, &buckets[b]);
            jac_add(&sum, &sum, &running);
        }
        window_sums[w] = sum;
        free(buckets);
    }
    for (unsigned w = num_windows; w-- > 0;) {
        for (unsigned i = 0; i < c; ++i) {
            jac_dbl(&total, &total);
        }
        jac_add(&total, &total, &window_sums[w]);
    }
    free(window_sums);
    free(k);
    *out = total;
}
/* sum_i P_i * s_i the slow way (the comparison used by scalar_multiplication.test.cpp:655-686) */
void orc_naive_msm(const orc_fe* scalars, const orc_affine* points, size_t n, size_t point_stride, orc_jac* out)
{
    orc_jac acc;
    orc_g1_set_infinity(&acc);
    for (size_t i = 0; i < n; ++i) {
        orc_jac t;
        orc_g1_mul(&points[i * point_stride], &scalars[i], &t);
        jac_add(&acc, &acc, &t);
    }
    *out = acc;
}

/* ------------------------------------------------------------------------------------------
 * NTT family (bb/polynomials/polynomial_arithmetic.cpp, evaluation_domain.cpp)
 * ---------------------------------------------------------------------------------------- */
static unsigned log2_exact(size_t n)
{
    unsigned l = 0;
    while (((size_t)1 << l) < n) {
        ++l;
    }
    return l;
}
/* polynomial_arithmetic.cpp:39-46 */
static uint32_t reverse_bits(uint32_t x, uint32_t bits)
{
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; ++i) {
        r = (r << 1) | ((x >> i) & 1);
    }
    return r;
}
/* radix-2 decimation-in-time: bit-reversal permutation, a twiddle-free first pass, then
 * log2(n)-1 passes  t = w^j * x[k+j+m]; x[k+j+m] = x[k+j] - t; x[k+j] += t  with
 * w = root^(n/2m) (fft_inner_serial :59-95; twiddles as compute_lookup_table_single,
 * evaluation_domain.cpp:33-54).  Natural order in, natural order out, lazy [0,2p) values. */
static void ntt_inner(orc_fe* x, size_t n, const orc_fe* root)
{
    const field_params* F = &FR;
    const unsigned lg = log2_exact(n);
    if (n <= 1) {
        return;
    }
    for (size_t i = 0; i < n; ++i) {
        size_t j = reverse_bits((uint32_t)i, lg);
        if (i < j) {
            orc_fe t = x[i];
            x[i] = x[j];
            x[j] = t;
        }
    }
    for (size_t k = 0; k < n; k += 2) {
        orc_fe t = x[k + 1];
        f_sub(F, &x[k + 1], &x[k], &t);
        f_add(F, &x[k], &x[k], &t);
    }
    orc_fe* tw = (orc_fe*)malloc((n / 2 ? n / 2 : 1) * sizeof(orc_fe));
    for (size_t m = 2; m < n; m *= 2) {
        orc_fe round_root;
        uint64_t e[4] = { (uint64_t)(n / (2 * m)), 0, 0, 0 };
        f_pow256(F, &round_root, root, e);
        f_one(F, &tw[0]);
        for (size_t j = 1; j < m; ++j) {
            f_mul(F, &tw[j], &tw[j - 1], &round_root);
        }
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (size_t b = 0; b < n / (2 * m); ++b) {
            const size_t k = b * 2 * m;
            for (size_t j = 0; j < m; ++j) {
                orc_fe t;
                f_mul(F, &t, &tw[j], &x[k + j + m]);
                f_sub(F, &x[k + j + m], &x[k + j], &t);
                f_add(F, &x[k + j], &x[k + j], &t);
            }
        }
    }
    free(tw);
}
/* target[i] = coeffs[i] * start * shift^i for i < generator_size ONLY (scale_by_generator :97-117) */
static void scale_by_generator(orc_fe* x, size_t generator_size, const orc_fe* start, const orc_fe* shift)
{
    orc_fe g = *start;
    for (size_t i = 0; i < generator_size; ++i) {
        f_mul(&FR, &x[i], &x[i], &g);
        f_mul(&FR, &g, &g, shift);
    }
}
static void scale_all(orc_fe* x, size_t n, const orc_fe* c)
{
    for (size_t i = 0; i < n; ++i) {
        f_mul(&FR, &x[i], &x[i], c);
    }
}
/* evaluation_domain ctor (evaluation_domain.cpp:57-76) */
void orc_domain_constants(size_t n, orc_fe* out6)
{
    orc_fe nn = { { (uint64_t)n, 0, 0, 0 } };
    orc_fr_root_of_unity(log2_exact(n), &out6[0]);
    f_invert(&FR, &out6[1], &out6[0]);
    f_to_mont(&FR, &out6[2], &nn);
    f_invert(&FR, &out6[3], &out6[2]);
    orc_fr_coset_generator(&out6[4]);
    f_invert(&FR, &out6[5], &out6[4]);
}
void orc_ntt(int kind, orc_fe* x, size_t n, size_t generator_size, const orc_fe* constant)
{
    orc_fe dc[6], one, t;
    if (generator_size == 0) {
        generator_size = n; /* evaluation_domain.cpp:64 */
    }
    orc_domain_constants(n, dc);
    f_one(&FR, &one);
    switch (kind) {
    case 0: /* fft :374-377 */
        ntt_inner(x, n, &dc[0]);
        break;
    case 1: /* ifft :379-385 */
        ntt_inner(x, n, &dc[1]);
        scale_all(x, n, &dc[3]);
        break;
    case 2: /* coset_fft :395-399 */
        scale_by_generator(x, generator_size, &one, &dc[4]);
        ntt_inner(x, n, &dc[0]);
        break;
    case 3: /* coset_ifft :480-484 (scales the WHOLE domain by g^-i) */
        ntt_inner(x, n, &dc[1]);
        scale_all(x, n, &dc[3]);
        scale_by_generator(x, n, &one, &dc[5]);
        break;
    case 4: /* fft_with_constant :387-393 */
        ntt_inner(x, n, &dc[0]);
        scale_all(x, n, constant);
        break;
    case 5: /* ifft_with_constant :471-478 */
        ntt_inner(x, n, &dc[1]);
        f_mul(&FR, &t, &dc[3], constant);
        scale_all(x, n, &t);
        break;
    case 6: /* coset_fft_with_constant :458-463 */
        scale_by_generator(x, generator_size, constant, &dc[4]);
        ntt_inner(x, n, &dc[0]);
        break;
    case 7: /* coset_fft_with_generator_shift :465-469 */
        f_mul(&FR, &t, &dc[4], constant);
        scale_by_generator(x, generator_size, &one, &t);
        ntt_inner(x, n, &dc[0]);
        break;
    default:
        break;
    }
}
/* coset_fft(coeffs, small, large, ext): ext independent n-point coset FFTs on the cosets
 * g * w_{ext n}^k, outputs interleaved out[ext*i + k] (:401-456). coeffs has ext*n entries,
 * the first n are the input polynomial. */
void orc_coset_fft_ext(orc_fe* x, size_t n, size_t ext)
{
    orc_fe dc[6], one, prim, gk;
    orc_domain_constants(n, dc);
    f_one(&FR, &one);
    orc_fr_root_of_unity(log2_exact(n) + log2_exact(ext), &prim);
    orc_fe* scratch = (orc_fe*)malloc(n * ext * sizeof(orc_fe));
    gk = dc[4];
    for (size_t k = 0; k < ext; ++k) {
        memcpy(scratch + k * n, x, n * sizeof(orc_fe));
        scale_by_generator(scratch + k * n, n, &one, &gk);
        ntt_inner(scratch + k * n, n, &dc[0]);
        f_mul(&FR, &gk, &gk, &prim);
    }
    for (size_t i = 0; i < n; ++i) {
        for (size_t k = 0; k < ext; ++k) {
            x[ext * i + k] = scratch[k * n + i];
        }
    }
    free(scratch);
}
/* Horner-equivalent evaluation sum_i c_i z^i (polynomial_arithmetic.cpp:507-538) */
void orc_evaluate(const orc_fe* coeffs, const orc_fe* z, size_t n, orc_fe* out)
{
    orc_fe acc;
    memset(&acc, 0, sizeof(acc));
    for (size_t i = n; i-- > 0;) {
        f_mul(&FR, &acc, &acc, z);
        f_add(&FR, &acc, &acc, &coeffs[i]);
    }
    *out = acc;
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
