// g1_team.cuh -- lane-cooperative XYZZ addition and doubling for the latency-bound tail of the MSM.
//
// A lone warp takes ~0.45 us per fq multiply (170 dependent-issue instructions, IMAD.WIDE at one per four cycles), so
// one XYZZ addition -- 14 multiplies in a row -- is ~6.3 us, and the bucket reduction / window combine are chains of a
// few dozen of them with far fewer workers than the GPU has schedulers.  The formulas have plenty of
// parallelism INSIDE one addition, though: their dependency depth is 4 multiplies (add) or 3 (doubling).  Here a TEAM of
// four adjacent lanes holds the same operands (replicated registers); in every "slot" each lane computes ONE of the
// independent products and the four results are all-gathered with shuffles (24 SHFL per slot).  An addition is then
// 4 multiply latencies + ~100 shuffles instead of 14 multiply latencies: ~3x shorter chains, at 4x the lanes -- exactly
// the trade wanted where lanes are idle anyway.
//
// All lanes of a team execute the same instruction stream with the same (replicated) data, so their control flow is
// identical; different teams of a warp may diverge from each other, hence every shuffle names only the team's lanes.
// Exceptional cases (infinity operands, P + P, P - P) are resolved after the cooperative part from replicated flags.
#pragma once
#include "g1.cuh"

namespace bbg {

struct Team {
    unsigned r;    // lane within the team, 0..3
    unsigned mask; // the team's four lanes
};
__device__ __forceinline__ Team team_of_lane()
{
    const unsigned lane = threadIdx.x & 31;
    Team t;
    t.r = lane & 3;
    t.mask = 0xFu << (lane & ~3u);
    return t;
}

__device__ __forceinline__ fq fq_sel(bool c, const fq& a, const fq& b)
{
    fq r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
__device__ __forceinline__ fq fq_sel4(unsigned r, const fq& v0, const fq& v1, const fq& v2, const fq& v3)
{
    return fq_sel(r & 2, fq_sel(r & 1, v3, v2), fq_sel(r & 1, v1, v0));
}
__device__ __forceinline__ fq fq_shfl_xor(const fq& v, int x, unsigned mask)
{
    fq r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_xor_sync(mask, v.l[i], x);
    return r;
}
// every lane contributes `mine`; afterwards every lane holds the four values in lane order
__device__ __forceinline__ void team_gather(const Team& t, const fq& mine, fq& v0, fq& v1, fq& v2, fq& v3)
{
    const fq other = fq_shfl_xor(mine, 1, t.mask);
    const fq lo = fq_sel(t.r & 1, other, mine); // the even lane's value of this pair
    const fq hi = fq_sel(t.r & 1, mine, other);
    const fq lo2 = fq_shfl_xor(lo, 2, t.mask);
    const fq hi2 = fq_shfl_xor(hi, 2, t.mask);
    v0 = fq_sel(t.r & 2, lo2, lo);
    v1 = fq_sel(t.r & 2, hi2, hi);
    v2 = fq_sel(t.r & 2, lo, lo2);
    v3 = fq_sel(t.r & 2, hi, hi2);
}

// a = 2 a ("dbl-2008-s-1", a = 0), operands replicated over the team: 3 multiply slots instead of 9 multiplies
__device__ __forceinline__ void xyzz_dbl_team(const Team& t, xyzz_t& a)
{
    const bool inf = xyzz_is_inf(a);
    const fq u = fe_dbl(a.y);
    fq g0, g1, g2, g3;
    // slot 1: v = u^2 | xx = x^2 | (idle) | (idle)
    {
        const fq l = fq_sel(t.r & 1, a.x, u);
        team_gather(t, fe_mul(l, l), g0, g1, g2, g3);
    }
    const fq v = g0, xx = g1;
    const fq m = fe_add(fe_dbl(xx), xx);
    // slot 2: w = u v | s = x v | mm = m^2 | zz' = v zz
    {
        const fq l = fq_sel4(t.r, u, a.x, m, v);
        const fq rr = fq_sel4(t.r, v, v, m, a.zz);
        team_gather(t, fe_mul(l, rr), g0, g1, g2, g3);
    }
    const fq w = g0, s = g1, mm = g2, zz3 = g3;
    const fq x3 = fe_sub(mm, fe_dbl(s));
    // slot 3: m (s - x3) | w y | zzz' = w zzz | (idle)
    {
        const fq l = fq_sel4(t.r, m, w, w, w);
        const fq rr = fq_sel4(t.r, fe_sub(s, x3), a.y, a.zzz, a.zzz);
        team_gather(t, fe_mul(l, rr), g0, g1, g2, g3);
    }
    if (!inf) {
        a.x = x3;
        a.y = fe_sub(g0, g1);
        a.zz = zz3;
        a.zzz = g2;
    }
}

// a += b ("add-2008-s"), both replicated over the team, all exceptional cases handled: 4 multiply slots instead of 14 multiplies
__device__ __forceinline__ void xyzz_add_team(const Team& t, xyzz_t& a, const xyzz_t& b)
{
    const bool a_inf = xyzz_is_inf(a), b_inf = xyzz_is_inf(b);
    fq g0, g1, g2, g3;
    // slot 1: u1 = a.x b.zz | u2 = b.x a.zz | s1 = a.y b.zzz | s2 = b.y a.zzz
    {
        const fq l = fq_sel4(t.r, a.x, b.x, a.y, b.y);
        const fq rr = fq_sel4(t.r, b.zz, a.zz, b.zzz, a.zzz);
        team_gather(t, fe_mul(l, rr), g0, g1, g2, g3);
    }
    const fq u1 = g0, s1 = g2;
    const fq p = fe_sub(g1, g0);
    const fq r = fe_sub(g3, g2);
    // slot 2: pp = p^2 | rr = r^2 | a.zz b.zz | a.zzz b.zzz
    {
        const fq l = fq_sel4(t.r, p, r, a.zz, a.zzz);
        const fq rr = fq_sel4(t.r, p, r, b.zz, b.zzz);
        team_gather(t, fe_mul(l, rr), g0, g1, g2, g3);
    }
    const fq pp = g0, r2 = g1, zz12 = g2, zzz12 = g3;
    // slot 3: ppp = p pp | q = u1 pp | zz' = zz12 pp | (idle)
    {
        const fq l = fq_sel4(t.r, p, u1, zz12, zz12);
        team_gather(t, fe_mul(l, pp), g0, g1, g2, g3);
    }
    const fq ppp = g0, q = g1, zz3 = g2;
    const fq x3 = fe_sub(fe_sub(r2, ppp), fe_dbl(q));
    // slot 4: r (q - x3) | s1 ppp | zzz' = zzz12 ppp | (idle)
    {
        const fq l = fq_sel4(t.r, r, s1, zzz12, zzz12);
        const fq rr = fq_sel4(t.r, fe_sub(q, x3), ppp, ppp, ppp);
        team_gather(t, fe_mul(l, rr), g0, g1, g2, g3);
    }
    // resolve (flags are replicated, so the whole team takes the same branch)
    if (b_inf) {
        return;
    }
    if (a_inf) {
        a = b;
        return;
    }
    if (__builtin_expect(fe_is_zero(p), 0)) {
        if (fe_is_zero(r)) {
            xyzz_dbl_team(t, a); // P + P
        } else {
            a = xyzz_infinity(); // P + (-P)
        }
        return;
    }
    a.x = x3;
    a.y = fe_sub(g0, g1);
    a.zz = zz3;
    a.zzz = g2;
}

} // namespace bbg
