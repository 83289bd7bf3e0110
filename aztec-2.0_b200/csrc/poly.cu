// poly.cu -- the prover's pointwise / scan arithmetic over fr that sits between the NTTs and the MSMs (SURVEY.md 8f
// ranks 2 and 3), so that a proof's polynomials can stay in HBM from the wire iFFTs to the quotient commitments.
//
//   kernel                      replaces (bb/ = barretenberg/src/aztec/)
//   k_turbo_quotient<KIND>      TransitionWidget::compute_quotient_contribution over the 4n coset domain
//                               (bb/plonk/proof_system/widgets/transition_widgets/transition_widget.hpp:293-307) with the
//                               TurboArithmetic / TurboFixedBase / TurboRange / TurboLogic gate identities
//                               (turbo_arithmetic_widget.hpp:17-143, turbo_fixed_base_widget.hpp:17-160,
//                               turbo_range_widget.hpp:30-161, turbo_logic_widget.hpp:17-183)
//   k_permutation_quotient      ProverPermutationWidget::compute_quotient_contribution
//                               (widgets/random_widgets/permutation_widget_impl.hpp:317-437), identity permutation polynomials
//   k_divide_vanishing          polynomial_arithmetic::divide_by_pseudo_vanishing_polynomial (bb/polynomials/polynomial_arithmetic.cpp:628-725)
//   k_lagrange_l1               polynomial_arithmetic::compute_lagrange_polynomial_fft (:546-626)
//   k_perm_terms + product scan + k_perm_finish   the grand product z(X) of compute_round_commitments (permutation_widget_impl.hpp:48-270)
//   k_eval_partial / k_eval_final                 polynomial_arithmetic::evaluate (:507-538)
//   k_opening_*                                   compute_kate_opening_coefficients / KateCommitmentScheme::compute_opening_polynomial
//                                                 (polynomial_arithmetic.cpp:727-751, commitment_scheme/kate_commitment_scheme.cpp:25-57)
//
// The gate identities are written here from their algebra (what each constraint says about the wire values), not from
// the reference's instruction sequence; every value is an element of fr, so the canonical results coincide.  All
// arithmetic is the register-resident Montgomery arithmetic of field.cuh; the kernels are elementwise over HBM-resident
// arrays (or log-depth scans), one thread per evaluation point, 128-bit loads and stores.
#include <algorithm>
#include <cstring>

#include "field.cuh"
#include "internal.hpp"
#include "poly.hpp"

namespace bbg {

using fr = Fe<FrParams>;

__device__ __forceinline__ fr ld(const fr* p, uint32_t i) { return fe_load_nc<FrParams>(p + i); }
__device__ __forceinline__ fr mul(const fr& a, const fr& b) { return fe_mul(a, b); }
__device__ __forceinline__ fr sqr(const fr& a) { return fe_sqr(a); }
__device__ __forceinline__ fr add(const fr& a, const fr& b) { return fe_add(a, b); }
__device__ __forceinline__ fr sub(const fr& a, const fr& b) { return fe_sub(a, b); }
__device__ __forceinline__ fr dbl(const fr& a) { return fe_add(a, a); }
__device__ __forceinline__ fr quad(const fr& a) { return dbl(dbl(a)); }

// ------------------------------------------------------------------------------------------------
// Turbo transition widgets: quotient[i] += gate identity at the i-th point of the 4n coset domain
// ------------------------------------------------------------------------------------------------
// x' denotes the value one gate later: X.w on the small domain = index i + 4 on the 4n domain (FFTGetter, transition_widget.hpp:160-169)
template <int KIND> __global__ void __launch_bounds__(128) k_turbo_quotient(const TurboParams P)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_large) return;
    const uint32_t is = (i + 4) & (P.n_large - 1);
    fr out;
    if (KIND == BBG_WIDGET_TURBO_ARITHMETIC) {
        // q_arith [ q_m w1 w2 + q_1 w1 + q_2 w2 + q_3 w3 + q_4 w4 + q_5 alpha w4 (w4 - 1)(w4 - 2) + q_c ] alpha_0
        //   + alpha_0 (q_arith^2 - q_arith) d (9 d - 2 d^2 - 7),  d = w3 - 4 w4   (the quad's high bit; active when q_arith = 2)
        const fr w1 = ld(P.p[BBG_POLY_W_1], i), w2 = ld(P.p[BBG_POLY_W_2], i), w3 = ld(P.p[BBG_POLY_W_3], i), w4 = ld(P.p[BBG_POLY_W_4], i);
        const fr qa = ld(P.p[BBG_POLY_Q_ARITHMETIC_SELECTOR], i);
        fr s = mul(mul(w1, w2), ld(P.p[BBG_POLY_Q_M], i));
        s = add(s, mul(w1, ld(P.p[BBG_POLY_Q_1], i)));
        s = add(s, mul(w2, ld(P.p[BBG_POLY_Q_2], i)));
        s = add(s, mul(w3, ld(P.p[BBG_POLY_Q_3], i)));
        s = add(s, mul(w4, ld(P.p[BBG_POLY_Q_4], i)));
        fr t = mul(sub(sqr(w4), w4), sub(w4, P.c_two));
        s = add(s, mul(mul(t, P.alpha), ld(P.p[BBG_POLY_Q_5], i)));
        s = add(s, ld(P.p[BBG_POLY_Q_C], i));
        s = mul(mul(s, qa), P.alpha_pow[0]);
        const fr d = sub(w3, quad(w4));
        const fr d9 = add(quad(dbl(d)), d); // 8 d + d
        fr h = sub(sub(d9, dbl(sqr(d))), P.c_seven);
        h = mul(mul(h, d), sub(sqr(qa), qa));
        out = add(s, mul(h, P.alpha_pow[0]));
    } else if (KIND == BBG_WIDGET_TURBO_FIXED_BASE) {
        // one step of the fixed-base scalar multiplication ladder on grumpkin (y^2 = x^3 - 17), 7 relations a_0 .. a_6
        const fr w1 = ld(P.p[BBG_POLY_W_1], i), w2 = ld(P.p[BBG_POLY_W_2], i), w3 = ld(P.p[BBG_POLY_W_3], i), w4 = ld(P.p[BBG_POLY_W_4], i);
        const fr w1n = ld(P.p[BBG_POLY_W_1], is), w2n = ld(P.p[BBG_POLY_W_2], is), w3n = ld(P.p[BBG_POLY_W_3], is), w4n = ld(P.p[BBG_POLY_W_4], is);
        const fr qc = ld(P.p[BBG_POLY_Q_C], i), qe = ld(P.p[BBG_POLY_Q_FIXED_BASE_SELECTOR], i);
        const fr d = sub(w4n, quad(w4)); // the next quad, in {-3, -1, 1, 3}
        const fr dq = mul(d, qe);
        // selector-weighted part
        fr lin = mul(mul(mul(sqr(d), qe), P.alpha_pow[1]), ld(P.p[BBG_POLY_Q_1], i));
        lin = add(lin, mul(mul(P.alpha_pow[1], qe), ld(P.p[BBG_POLY_Q_2], i)));
        const fr dw3n = mul(d, w3n);
        fr q3 = mul(mul(mul(sub(w1n, w1), dw3n), P.alpha_pow[3]), qe);
        q3 = add(q3, mul(dbl(mul(mul(dw3n, w2), P.alpha_pow[2])), qe));
        lin = add(lin, mul(q3, ld(P.p[BBG_POLY_Q_3], i)));
        const fr qeqc = mul(qe, qc);
        const fr w3qeqc = mul(w3, qeqc);
        lin = add(lin, mul(mul(w3qeqc, P.alpha_pow[5]), ld(P.p[BBG_POLY_Q_4], i)));
        lin = add(lin, mul(mul(mul(sub(P.c_one, w4), qeqc), P.alpha_pow[5]), ld(P.p[BBG_POLY_Q_5], i)));
        lin = add(lin, mul(mul(w3qeqc, P.alpha_pow[6]), ld(P.p[BBG_POLY_Q_M], i)));
        // selector-free part
        const fr acc_id = mul(mul(mul(add(d, P.c_one), add(d, P.c_three)), mul(sub(d, P.c_one), sub(d, P.c_three))), P.alpha_pow[0]);
        const fr xalpha_id = fe_neg(mul(w3n, P.alpha_pow[1]));
        const fr dx = sub(w3n, w1);
        fr xacc = mul(add(add(w1n, w1), w3n), sqr(dx));
        xacc = sub(xacc, sub(add(mul(sqr(w3n), w3n), sqr(w2)), P.c_17)); // - (x_alpha^3 + y^2 + b), b = -17
        xacc = add(xacc, dbl(mul(dq, w2)));
        xacc = mul(xacc, P.alpha_pow[2]);
        fr yacc = add(mul(add(w2n, w2), dx), mul(sub(w1, w1n), sub(w2, dq)));
        yacc = mul(yacc, P.alpha_pow[3]);
        const fr w4m1 = sub(w4, P.c_one);
        fr init = mul(mul(w4m1, sub(w4m1, w3)), P.alpha_pow[4]);
        init = sub(init, mul(mul(w1, w3), P.alpha_pow[5]));
        init = add(init, mul(sub(mul(sub(P.c_one, w4), qc), mul(w2, w3)), P.alpha_pow[6]));
        fr gate = add(add(add(mul(init, qc), acc_id), add(xalpha_id, xacc)), yacc);
        out = add(lin, mul(gate, qe));
    } else if (KIND == BBG_WIDGET_TURBO_RANGE) {
        // four base-4 digits per gate: d_k (d_k - 1)(d_k - 2)(d_k - 3) = 0 for the accumulator differences
        const fr w1 = ld(P.p[BBG_POLY_W_1], i), w2 = ld(P.p[BBG_POLY_W_2], i), w3 = ld(P.p[BBG_POLY_W_3], i), w4 = ld(P.p[BBG_POLY_W_4], i);
        const fr w4n = ld(P.p[BBG_POLY_W_4], is);
        fr acc = fe_zero<FrParams>();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const fr d = k == 0 ? sub(w3, quad(w4)) : (k == 1 ? sub(w2, quad(w3)) : (k == 2 ? sub(w1, quad(w2)) : sub(w4n, quad(w1))));
            fr t = mul(mul(sub(sqr(d), d), sub(d, P.c_two)), sub(d, P.c_three));
            acc = add(acc, mul(t, P.alpha_pow[k]));
        }
        out = mul(acc, ld(P.p[BBG_POLY_Q_RANGE_SELECTOR], i));
    } else {
        // AND / XOR of one base-4 digit of each operand per gate (q_c selects which): a, b = operand digits, c = output digit
        const fr w1 = ld(P.p[BBG_POLY_W_1], i), w2 = ld(P.p[BBG_POLY_W_2], i), w3 = ld(P.p[BBG_POLY_W_3], i), w4 = ld(P.p[BBG_POLY_W_4], i);
        const fr w1n = ld(P.p[BBG_POLY_W_1], is), w2n = ld(P.p[BBG_POLY_W_2], is), w4n = ld(P.p[BBG_POLY_W_4], is);
        const fr qc = ld(P.p[BBG_POLY_Q_C], i);
        const fr a = sub(w1n, quad(w1)), b = sub(w2n, quad(w2)), c = sub(w4n, quad(w4));
        const fr s = add(a, b);
        const fr a2 = sqr(a), b2 = sqr(b);
        const fr q = add(a2, b2);
        // alpha^3 (2 a b - 2 w3) + alpha^2 a(a-1)(a-2)(a-3) + alpha b(b-1)(b-2)(b-3) + [the digit identity]
        fr id = mul(sub(sub(sqr(s), q), dbl(w3)), P.alpha);
        const fr a2a = sub(a2, a);
        id = add(id, mul(add(sub(a2a, quad(a)), P.c_six), a2a)); // (a^2 - 5a + 6)(a^2 - a)
        id = mul(id, P.alpha);
        const fr b2b = sub(b2, b);
        id = add(id, mul(add(sub(b2b, quad(b)), P.c_six), b2b));
        id = mul(id, P.alpha);
        // 3 c + 3 s - 2 w3 ( w3 (4 w3 - 18 s + 81) + 18 q - 81 s + 83 ) + q_c (9 c - 3 s)
        const fr s3 = add(dbl(s), s);
        const fr s9 = add(dbl(s3), s3);
        const fr s18 = dbl(s9);
        const fr s81 = add(quad(s18), s9);
        const fr q3 = add(dbl(q), q);
        const fr q18 = dbl(add(dbl(q3), q3));
        fr inner = mul(add(sub(quad(w3), s18), P.c_81), w3);
        inner = add(inner, add(sub(q18, s81), P.c_83));
        inner = mul(inner, w3);
        const fr c3 = add(dbl(c), c);
        const fr c9 = add(dbl(c3), c3);
        fr tail = sub(add(c3, s3), dbl(inner));
        tail = add(tail, mul(sub(c9, s3), qc));
        id = mul(add(id, tail), P.alpha_pow[0]);
        out = mul(id, ld(P.p[BBG_POLY_Q_LOGIC_SELECTOR], i));
    }
    fe_store(P.quotient + i, fe_add(fe_load<FrParams>(P.quotient + i), out));
}

// ------------------------------------------------------------------------------------------------
// permutation argument: quotient[i] = alpha_0 (numerator - denominator), assignment (the first widget to run)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_permutation_quotient(const PermParams P)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_large) return;
    const uint32_t mask = P.n_large - 1;
    // X beta at this point of the coset: g w_{4n}^i beta; wire k uses k_k X beta with k_0 = 1, k_j = coset_generator(j - 1)
    const fr xb = mul(ld(P.roots, i), P.g_beta);
    fr num = fe_zero<FrParams>(), den = fe_zero<FrParams>();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < (int)P.width) {
            const fr wg = add(ld(P.wires[k], i), P.gamma);
            const fr idt = add(k == 0 ? xb : mul(P.coset_gen[k - 1], xb), wg);
            const fr sgt = add(mul(ld(P.sigmas[k], i), P.beta), wg);
            num = k == 0 ? idt : mul(num, idt);
            den = k == 0 ? sgt : mul(den, sgt);
        }
    }
    const fr z = ld(P.z, i), zn = ld(P.z, (i + 4) & mask);
    num = mul(num, z);
    den = mul(den, zn);
    // (z(X w) - delta) alpha_0 L_end(X) + (z(X) - 1) alpha_0^2 L_1(X)
    fr t = mul(mul(sub(zn, P.public_input_delta), P.alpha_base), ld(P.l_start, (i + 4 + 4 * P.roots_cut) & mask));
    num = add(num, t);
    t = mul(mul(sub(z, P.c_one), P.alpha_squared), ld(P.l_start, i));
    num = add(num, t);
    fe_store(P.quotient + i, mul(sub(num, den), P.alpha_base));
}

// ------------------------------------------------------------------------------------------------
// quotient[i] *= (prod_k (g w_{4n}^i - w_n^{-(k+1)})) / ((g w_{4n}^i)^n - 1)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_divide_vanishing(fr* __restrict__ q, const fr* __restrict__ roots, const VanishParams P)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_large) return;
    const fr x = mul(ld(roots, i), P.g); // g w_{4n}^i
    fr v = fe_load<FrParams>(q + i);
    const uint32_t j = i & (P.subgroup - 1);
    fr inv = P.inv_sub[0];
#pragma unroll
    for (int s = 1; s < 8; ++s) {
        if (j == (uint32_t)s) inv = P.inv_sub[s];
    }
    v = mul(v, inv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < (int)P.roots_cut) v = mul(v, add(x, P.numer[k]));
    }
    fe_store(q + i, v);
}

// ------------------------------------------------------------------------------------------------
// a^(p-2) for fr (field_impl.hpp:323-329); used by the per-thread batch inversions below
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ fr fr_inv(const fr& a)
{
    fr acc = fe_one<FrParams>();
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = FrParams::P(i);
    e[0] -= 2;
#pragma unroll 1
    for (int i = 253; i >= 0; --i) {
        acc = fe_sqr(acc);
        uint32_t limb = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) limb = (i >> 5) == k ? e[k] : limb;
        if ((limb >> (i & 31)) & 1) acc = fe_mul(acc, a);
    }
    return acc;
}

// l[i] = ((g^n w_ext^(i mod ext)) - 1) / n / (g w_{ext n}^i - 1): the ext*n-point coset evaluations of L_1
static constexpr int INV_CHUNK = 16;
__global__ void __launch_bounds__(128) k_lagrange_l1(fr* __restrict__ out, const fr* __restrict__ roots, const LagrangeParams P)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i0 = t * INV_CHUNK;
    if (i0 >= P.n_large) return;
    // Montgomery's trick over this thread's chunk (n_large is a power of two >= INV_CHUNK or the tail is guarded)
    fr den[INV_CHUNK];
    fr run = fe_one<FrParams>();
#pragma unroll 1
    for (int j = 0; j < INV_CHUNK; ++j) {
        if (i0 + j < P.n_large) {
            den[j] = sub(mul(ld(roots, i0 + j), P.g), P.c_one);
            const fr prev = run;
            run = mul(run, den[j]);
            fe_store(out + i0 + j, prev); // prefix product before this element
        }
    }
    fr inv = fr_inv(run);
#pragma unroll 1
    for (int j = INV_CHUNK - 1; j >= 0; --j) {
        if (i0 + j < P.n_large) {
            const fr prefix = fe_load<FrParams>(out + i0 + j);
            fr v = mul(inv, prefix); // 1 / den[j]
            inv = mul(inv, den[j]);
            const uint32_t s = (i0 + j) & (P.subgroup - 1);
            fr numer = P.numer_sub[0];
#pragma unroll
            for (int q = 1; q < 8; ++q) {
                if (s == (uint32_t)q) numer = P.numer_sub[q];
            }
            fe_store(out + i0 + j, mul(v, numer));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// grand product z: z[0] = 1, z[i] = prod_{j < i} num_j / den_j
// ------------------------------------------------------------------------------------------------
// terms: num[i] = prod_k (w_k[i] + gamma + beta k_k w^i), den[i] = prod_k (w_k[i] + gamma + beta sigma_k[i])
__global__ void __launch_bounds__(128) k_perm_terms(const GrandParams P, fr* __restrict__ num, fr* __restrict__ den)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const fr xb = mul(ld(P.roots, i << P.root_stride_log), P.beta); // w_n^i beta
    fr a = fe_zero<FrParams>(), b = fe_zero<FrParams>();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < (int)P.width) {
            const fr wg = add(ld(P.wires[k], i), P.gamma);
            const fr idt = add(k == 0 ? xb : mul(P.coset_gen[k - 1], xb), wg);
            const fr sgt = add(mul(ld(P.sigmas[k], i), P.beta), wg);
            a = k == 0 ? idt : mul(a, idt);
            b = k == 0 ? sgt : mul(b, sgt);
        }
    }
    fe_store(num + i, a);
    fe_store(den + i, b);
}

// inclusive prefix PRODUCT, three kernels (chunk products -> scan of the chunk products by one CTA -> apply); two arrays
// at a time (blockIdx.y selects num / den).
static constexpr int SCAN_CHUNK = 16;
static constexpr int SCAN_TOP_THREADS = 1024;
__global__ void __launch_bounds__(128) k_prod_chunks(const fr* __restrict__ a0, const fr* __restrict__ a1, uint32_t n, fr* __restrict__ tot0, fr* __restrict__ tot1)
{
    const fr* a = blockIdx.y ? a1 : a0;
    fr* tot = blockIdx.y ? tot1 : tot0;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i0 = t * SCAN_CHUNK;
    if (i0 >= n) return;
    fr run = fe_load<FrParams>(a + i0);
#pragma unroll 1
    for (int j = 1; j < SCAN_CHUNK; ++j) {
        if (i0 + j < n) run = mul(run, fe_load<FrParams>(a + i0 + j));
    }
    fe_store(tot + t, run);
}
// exclusive prefix product of m chunk totals in place (tot[t] <- prod_{u < t} tot[u]); one CTA per array
__global__ void __launch_bounds__(SCAN_TOP_THREADS) k_prod_top(fr* __restrict__ tot0, fr* __restrict__ tot1, uint32_t m)
{
    __shared__ fr sm[SCAN_TOP_THREADS];
    fr* tot = blockIdx.y ? tot1 : tot0;
    const uint32_t per = (m + SCAN_TOP_THREADS - 1) / SCAN_TOP_THREADS;
    const uint32_t lo = threadIdx.x * per;
    fr run = fe_one<FrParams>();
    for (uint32_t j = 0; j < per; ++j) {
        if (lo + j < m) run = mul(run, fe_load<FrParams>(tot + lo + j));
    }
    sm[threadIdx.x] = run;
    __syncthreads();
    // Hillis-Steele over the thread products (10 steps)
    for (uint32_t d = 1; d < SCAN_TOP_THREADS; d <<= 1) {
        fr v = sm[threadIdx.x];
        const bool take = threadIdx.x >= d;
        fr o = take ? sm[threadIdx.x - d] : fe_one<FrParams>();
        __syncthreads();
        if (take) sm[threadIdx.x] = mul(v, o);
        __syncthreads();
    }
    fr pre = threadIdx.x ? sm[threadIdx.x - 1] : fe_one<FrParams>();
    for (uint32_t j = 0; j < per; ++j) {
        if (lo + j < m) {
            const fr v = fe_load<FrParams>(tot + lo + j);
            fe_store(tot + lo + j, pre);
            pre = mul(pre, v);
        }
    }
}
// z[i + 1] = N_i / D_i with N, D the inclusive prefix products of num / den (i < n - 1), z[0] = 1; the division is a
// per-thread Montgomery batch inversion over the chunk
__global__ void __launch_bounds__(128) k_perm_finish(const fr* __restrict__ num, const fr* __restrict__ den, uint32_t n, const fr* __restrict__ tot_num,
                                                     const fr* __restrict__ tot_den, fr* __restrict__ z)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i0 = t * SCAN_CHUNK;
    if (i0 >= n) return;
    fr N[SCAN_CHUNK], D[SCAN_CHUNK], pre[SCAN_CHUNK];
    fr rn = fe_load<FrParams>(tot_num + t), rd = fe_load<FrParams>(tot_den + t);
    fr run = fe_one<FrParams>();
#pragma unroll 1
    for (int j = 0; j < SCAN_CHUNK; ++j) {
        if (i0 + j < n) {
            rn = mul(rn, fe_load<FrParams>(num + i0 + j));
            rd = mul(rd, fe_load<FrParams>(den + i0 + j));
            N[j] = rn;
            D[j] = rd;
            pre[j] = run;
            run = mul(run, rd);
        }
    }
    fr inv = fr_inv(run);
#pragma unroll 1
    for (int j = SCAN_CHUNK - 1; j >= 0; --j) {
        if (i0 + j < n) {
            const fr dinv = mul(inv, pre[j]);
            inv = mul(inv, D[j]);
            if (i0 + j + 1 < n) fe_store(z + i0 + j + 1, mul(N[j], dinv));
        }
    }
    if (t == 0) fe_store(z, fe_one<FrParams>());
}

// ------------------------------------------------------------------------------------------------
// evaluate: sum_i c_i z^i
// ------------------------------------------------------------------------------------------------
static constexpr int EVAL_CHUNK = 16;
static constexpr int EVAL_THREADS = 256;
// blockIdx.y selects the polynomial (batched evaluation: the ~30 opening evaluations of round 5 in one launch)
__global__ void __launch_bounds__(EVAL_THREADS) k_eval_partial(const EvalBatchParams P, fr* __restrict__ partial)
{
    __shared__ fr sm[EVAL_THREADS];
    const uint32_t y = blockIdx.y;
    const fr* c = P.coeffs[y];
    const uint32_t n = P.n[y];
    const fr z = P.z[y];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i0 = t * EVAL_CHUNK;
    fr acc = fe_zero<FrParams>();
    if (i0 < n) {
        // Horner over the chunk, then shift by z^(i0)
#pragma unroll 1
        for (int j = EVAL_CHUNK - 1; j >= 0; --j) {
            acc = mul(acc, z);
            if (i0 + j < n) acc = add(acc, fe_load_nc<FrParams>(c + i0 + j));
        }
        acc = mul(acc, fe_pow(P.z_chunk[y], (uint64_t)t));
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t d = EVAL_THREADS / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) sm[threadIdx.x] = add(sm[threadIdx.x], sm[threadIdx.x + d]);
        __syncthreads();
    }
    if (threadIdx.x == 0) fe_store(partial + (size_t)y * gridDim.x + blockIdx.x, sm[0]);
}
__global__ void __launch_bounds__(EVAL_THREADS) k_eval_final(const fr* __restrict__ partial, uint32_t m, fr* __restrict__ out)
{
    __shared__ fr sm[EVAL_THREADS];
    partial += (size_t)blockIdx.x * m;
    fr acc = fe_zero<FrParams>();
    for (uint32_t i = threadIdx.x; i < m; i += EVAL_THREADS) acc = add(acc, fe_load<FrParams>(partial + i));
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t d = EVAL_THREADS / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) sm[threadIdx.x] = add(sm[threadIdx.x], sm[threadIdx.x + d]);
        __syncthreads();
    }
    if (threadIdx.x == 0) fe_store(out + blockIdx.x, fe_reduce_once(sm[0]));
}

// ------------------------------------------------------------------------------------------------
// opening polynomial W(X) = (F(X) - F(z)) / (X - z): w_i = sum_{j > i} f_j z^(j - i - 1), i < n_out; F has n_in coefficients.
// The reference runs the recurrence from the constant term up (dest[i] = (src[i] - dest[i-1]) / -z, kate_commitment_scheme.cpp:48-53);
// with F(z) exact the two recurrences define the same polynomial, and this direction needs no inversion and
// yields F(z) = f_0 + z w_0 as a by-product.  Suffix scan in three kernels like the prefix product above.
// ------------------------------------------------------------------------------------------------
static constexpr int OPEN_CHUNK = 16;
// chunk t covers indices [i0, i0 + CHUNK); S_t = sum_{j in chunk} f_j z^(j - i0)
__global__ void __launch_bounds__(128) k_open_chunks(const fr* __restrict__ f, uint32_t n_in, fr z, fr* __restrict__ tot)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i0 = t * OPEN_CHUNK;
    if (i0 >= n_in) return;
    fr acc = fe_zero<FrParams>();
#pragma unroll 1
    for (int j = OPEN_CHUNK - 1; j >= 0; --j) {
        acc = mul(acc, z);
        if (i0 + j < n_in) acc = add(acc, fe_load_nc<FrParams>(f + i0 + j));
    }
    fe_store(tot + t, acc);
}
// carry[t] = sum_{u > t} S_u z^((u - t - 1) CHUNK)  (what flows into chunk t from above), one CTA, in place
static constexpr int OPEN_TOP_THREADS = 512; // two fr arrays of shared memory: 32 KB
__global__ void __launch_bounds__(OPEN_TOP_THREADS) k_open_top(fr* __restrict__ tot, uint32_t m, fr zc /* z^CHUNK */)
{
    __shared__ fr sm[OPEN_TOP_THREADS];
    __shared__ fr pw[OPEN_TOP_THREADS];
    const uint32_t per = (m + OPEN_TOP_THREADS - 1) / OPEN_TOP_THREADS;
    // thread x owns chunks [lo, lo + per), processed from the top; A_x = sum_{u in own} S_u zc^(u - lo)
    const uint32_t lo = threadIdx.x * per;
    fr acc = fe_zero<FrParams>();
    for (int j = (int)per - 1; j >= 0; --j) {
        acc = mul(acc, zc);
        if (lo + j < m) acc = add(acc, fe_load<FrParams>(tot + lo + j));
    }
    const fr zp = fe_pow(zc, (uint64_t)per); // zc^per: one thread's span
    sm[threadIdx.x] = acc;
    pw[threadIdx.x] = zp;
    __syncthreads();
    // suffix scan of (A, span) pairs: combined(x) = A_x + span_x * combined(x + d)
    for (uint32_t d = 1; d < OPEN_TOP_THREADS; d <<= 1) {
        const bool take = threadIdx.x + d < OPEN_TOP_THREADS;
        fr a = sm[threadIdx.x], p = pw[threadIdx.x];
        fr oa = take ? sm[threadIdx.x + d] : fe_zero<FrParams>();
        fr op = take ? pw[threadIdx.x + d] : fe_one<FrParams>();
        __syncthreads();
        if (take) {
            sm[threadIdx.x] = add(a, mul(p, oa));
            pw[threadIdx.x] = mul(p, op);
        }
        __syncthreads();
    }
    // carry into this thread's TOP chunk from everything above = combined(x + 1)
    fr carry = threadIdx.x + 1 < OPEN_TOP_THREADS ? sm[threadIdx.x + 1] : fe_zero<FrParams>();
    for (int j = (int)per - 1; j >= 0; --j) {
        if (lo + j < m) {
            const fr s = fe_load<FrParams>(tot + lo + j);
            fe_store(tot + lo + j, carry);
            carry = add(s, mul(carry, zc));
        }
    }
}
__global__ void __launch_bounds__(128) k_open_apply(const fr* __restrict__ f, uint32_t n_in, uint32_t n_out, fr z, const fr* __restrict__ carry,
                                                    fr* __restrict__ w, fr* __restrict__ f_at_z)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i0 = t * OPEN_CHUNK;
    if (i0 >= n_in) return;
    // r = sum_{j > i} f_j z^(j - i - 1), walking down from the top of the chunk
    fr r = fe_load<FrParams>(carry + t);
#pragma unroll 1
    for (int j = OPEN_CHUNK - 1; j >= 0; --j) {
        const uint32_t i = i0 + j;
        if (i < n_in) {
            const fr fi = fe_load_nc<FrParams>(f + i);
            if (i < n_out) fe_store(w + i, r);
            r = add(fi, mul(r, z));
            if (i == 0 && f_at_z != nullptr) fe_store(f_at_z, fe_reduce_once(r));
        }
    }
}

// dest[i] = (base ? base[i] : 0) + sum_k polys[k][i] * scalars[k]   (the accumulation of KateCommitmentScheme::batch_open,
// bb/plonk/proof_system/commitment_scheme/kate_commitment_scheme.cpp:213-222)
__global__ void __launch_bounds__(128) k_linear_combination(const LinCombParams P)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    fr acc = P.base ? fe_load_nc<FrParams>(P.base + i) : fe_zero<FrParams>();
#pragma unroll 1
    for (uint32_t k = 0; k < P.count; ++k) {
        acc = add(acc, mul(fe_load_nc<FrParams>(P.polys[k] + i), P.scalars[k]));
    }
    fe_store(P.dest + i, acc);
}

// dst[i] = i < n ? src[i] : 0 for i < total (copy_polynomial + zero padding, polynomial_arithmetic.cpp copy_polynomial)
__global__ void __launch_bounds__(256) k_copy_pad(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16, size_t total16)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total16) return;
    dst[i] = i < n16 ? src[i] : make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static fr dev_fr(const hf::Fr& a)
{
    fr r;
    for (int i = 0; i < 4; ++i) {
        r.l[2 * i] = (uint32_t)a.d[i];
        r.l[2 * i + 1] = (uint32_t)(a.d[i] >> 32);
    }
    return r;
}
static hf::Fr neg_small(uint64_t k) { return hf::sub(hf::zero(), hf::from_u64(k)); }

int poly_turbo_quotient_device(Context* ctx, int kind, const void* const* d_polys, size_t n_large, const void* alpha_base, const void* alpha,
                               void* d_quotient, cudaStream_t st)
{
    if (n_large == 0 || (n_large & (n_large - 1)) || n_large >= (1ull << 32)) {
        set_last_error("turbo_quotient: the large domain must be a power of two below 2^32");
        return BBG_ERR_ARG;
    }
    static const int need[4][16] = {
        { BBG_POLY_W_1, BBG_POLY_W_2, BBG_POLY_W_3, BBG_POLY_W_4, BBG_POLY_Q_1, BBG_POLY_Q_2, BBG_POLY_Q_3, BBG_POLY_Q_4, BBG_POLY_Q_5, BBG_POLY_Q_M,
          BBG_POLY_Q_C, BBG_POLY_Q_ARITHMETIC_SELECTOR, -1 },
        { BBG_POLY_W_1, BBG_POLY_W_2, BBG_POLY_W_3, BBG_POLY_W_4, BBG_POLY_Q_1, BBG_POLY_Q_2, BBG_POLY_Q_3, BBG_POLY_Q_4, BBG_POLY_Q_5, BBG_POLY_Q_M,
          BBG_POLY_Q_C, BBG_POLY_Q_FIXED_BASE_SELECTOR, -1 },
        { BBG_POLY_W_1, BBG_POLY_W_2, BBG_POLY_W_3, BBG_POLY_W_4, BBG_POLY_Q_RANGE_SELECTOR, -1 },
        { BBG_POLY_W_1, BBG_POLY_W_2, BBG_POLY_W_3, BBG_POLY_W_4, BBG_POLY_Q_C, BBG_POLY_Q_LOGIC_SELECTOR, -1 },
    };
    if (kind < 0 || kind > 3) {
        set_last_error("turbo_quotient: unknown widget kind");
        return BBG_ERR_ARG;
    }
    TurboParams P;
    memset(&P, 0, sizeof(P));
    for (int k = 0; need[kind][k] >= 0; ++k) {
        if (d_polys[need[kind][k]] == nullptr) {
            set_last_error("turbo_quotient: a polynomial this widget reads is missing (index " + std::to_string(need[kind][k]) + ")");
            return BBG_ERR_ARG;
        }
    }
    for (int k = 0; k < BBG_POLY_COUNT; ++k) P.p[k] = (const fr*)d_polys[k];
    P.quotient = (fr*)d_quotient;
    P.n_large = (uint32_t)n_large;
    const hf::Fr a0 = hf::load(alpha_base), a = hf::load(alpha);
    hf::Fr pw = a0;
    for (int k = 0; k < 7; ++k) {
        P.alpha_pow[k] = dev_fr(hf::reduce(pw));
        pw = hf::mul(pw, a);
    }
    P.alpha = dev_fr(hf::reduce(a));
    P.c_one = dev_fr(hf::one());
    P.c_two = dev_fr(hf::from_u64(2));
    P.c_three = dev_fr(hf::from_u64(3));
    P.c_six = dev_fr(hf::from_u64(6));
    P.c_seven = dev_fr(hf::from_u64(7));
    P.c_17 = dev_fr(hf::from_u64(17));
    P.c_81 = dev_fr(hf::from_u64(81));
    P.c_83 = dev_fr(hf::from_u64(83));
    const unsigned blocks = div_up(n_large, 128);
    switch (kind) {
    case BBG_WIDGET_TURBO_ARITHMETIC: k_turbo_quotient<BBG_WIDGET_TURBO_ARITHMETIC><<<blocks, 128, 0, st>>>(P); break;
    case BBG_WIDGET_TURBO_FIXED_BASE: k_turbo_quotient<BBG_WIDGET_TURBO_FIXED_BASE><<<blocks, 128, 0, st>>>(P); break;
    case BBG_WIDGET_TURBO_RANGE: k_turbo_quotient<BBG_WIDGET_TURBO_RANGE><<<blocks, 128, 0, st>>>(P); break;
    default: k_turbo_quotient<BBG_WIDGET_TURBO_LOGIC><<<blocks, 128, 0, st>>>(P); break;
    }
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

static unsigned log2u(size_t n)
{
    unsigned lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    return lg;
}

int poly_permutation_quotient_device(Context* ctx, const PermArgs& A, cudaStream_t st)
{
    if (A.width < 1 || A.width > 4 || A.n_large == 0 || (A.n_large & (A.n_large - 1)) || A.n_large >= (1ull << 32)) {
        set_last_error("permutation_quotient: program width 1..4 and a power-of-two domain below 2^32");
        return BBG_ERR_ARG;
    }
    const fr* roots = nullptr;
    int rc = ntt_root_table(ctx, log2u(A.n_large), (const void**)&roots, st);
    if (rc) return rc;
    PermParams P;
    memset(&P, 0, sizeof(P));
    for (unsigned k = 0; k < A.width; ++k) {
        P.wires[k] = (const fr*)A.d_wires[k];
        P.sigmas[k] = (const fr*)A.d_sigmas[k];
    }
    P.z = (const fr*)A.d_z;
    P.l_start = (const fr*)A.d_l_start;
    P.roots = roots;
    P.quotient = (fr*)A.d_quotient;
    P.n_large = (uint32_t)A.n_large;
    P.width = A.width;
    P.roots_cut = A.roots_cut;
    const hf::Fr g = hf::from_u64(5); // evaluation_domain::generator = fr::coset_generator(0)
    P.g_beta = dev_fr(hf::reduce(hf::mul(g, A.beta)));
    P.beta = dev_fr(hf::reduce(A.beta));
    P.gamma = dev_fr(hf::reduce(A.gamma));
    P.alpha_base = dev_fr(hf::reduce(A.alpha_base));
    P.alpha_squared = dev_fr(hf::reduce(hf::sqr(A.alpha_base)));
    P.public_input_delta = dev_fr(hf::reduce(A.public_input_delta));
    P.c_one = dev_fr(hf::one());
    for (int k = 0; k < 3; ++k) P.coset_gen[k] = dev_fr(hf::from_u64(5 + k)); // fr::coset_generator(k), bb/ecc/curves/bn254/fr.hpp:44-59
    k_permutation_quotient<<<div_up(A.n_large, 128), 128, 0, st>>>(P);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

int poly_divide_vanishing_device(Context* ctx, void* d_q, size_t n_small, size_t n_large, unsigned roots_cut, cudaStream_t st)
{
    if (n_small == 0 || n_large < n_small || (n_large & (n_large - 1)) || (n_small & (n_small - 1)) || n_large / n_small > 8 || roots_cut > 4 ||
        n_large >= (1ull << 32)) {
        set_last_error("divide_by_pseudo_vanishing_polynomial: power-of-two domains, extension <= 8, <= 4 roots cut out");
        return BBG_ERR_ARG;
    }
    const fr* roots = nullptr;
    int rc = ntt_root_table(ctx, log2u(n_large), (const void**)&roots, st);
    if (rc) return rc;
    VanishParams P;
    memset(&P, 0, sizeof(P));
    P.n_large = (uint32_t)n_large;
    P.subgroup = (uint32_t)(n_large / n_small);
    P.roots_cut = roots_cut;
    const hf::Fr g = hf::from_u64(5);
    P.g = dev_fr(g);
    // (g^n w_ext^j - 1)^-1  (compute_multiplicative_subgroup :119-138 + batch_invert)
    hf::Fr gn = g;
    for (unsigned i = 0; i < log2u(n_small); ++i) gn = hf::sqr(gn);
    const hf::Fr wsub = ntt_root_of_unity(log2u(P.subgroup));
    hf::Fr cur = gn;
    for (unsigned j = 0; j < P.subgroup; ++j) {
        P.inv_sub[j] = dev_fr(hf::reduce(hf::invert(hf::sub(cur, hf::one()))));
        cur = hf::mul(cur, wsub);
    }
    // -w_n^-(k+1)
    const hf::Fr winv = hf::invert(ntt_root_of_unity(log2u(n_small)));
    hf::Fr c = hf::sub(hf::zero(), winv);
    for (unsigned k = 0; k < roots_cut; ++k) {
        P.numer[k] = dev_fr(hf::reduce(c));
        c = hf::mul(c, winv);
    }
    k_divide_vanishing<<<div_up(n_large, 128), 128, 0, st>>>((fr*)d_q, roots, P);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

int poly_lagrange_l1_device(Context* ctx, void* d_out, size_t n_small, size_t n_large, cudaStream_t st)
{
    if (n_small == 0 || n_large < n_small || (n_large & (n_large - 1)) || (n_small & (n_small - 1)) || n_large / n_small > 8 ||
        n_large >= (1ull << 32)) {
        set_last_error("compute_lagrange_polynomial_fft: power-of-two domains, extension <= 8");
        return BBG_ERR_ARG;
    }
    const fr* roots = nullptr;
    int rc = ntt_root_table(ctx, log2u(n_large), (const void**)&roots, st);
    if (rc) return rc;
    LagrangeParams P;
    memset(&P, 0, sizeof(P));
    P.n_large = (uint32_t)n_large;
    P.subgroup = (uint32_t)(n_large / n_small);
    const hf::Fr g = hf::from_u64(5);
    P.g = dev_fr(g);
    P.c_one = dev_fr(hf::one());
    hf::Fr gn = g;
    for (unsigned i = 0; i < log2u(n_small); ++i) gn = hf::sqr(gn);
    const hf::Fr wsub = ntt_root_of_unity(log2u(P.subgroup));
    const hf::Fr n_inv = hf::invert(hf::from_u64(n_small));
    hf::Fr cur = gn;
    for (unsigned j = 0; j < P.subgroup; ++j) {
        P.numer_sub[j] = dev_fr(hf::reduce(hf::mul(hf::sub(cur, hf::one()), n_inv)));
        cur = hf::mul(cur, wsub);
    }
    k_lagrange_l1<<<div_up(div_up(n_large, INV_CHUNK), 128), 128, 0, st>>>((fr*)d_out, roots, P);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

int poly_grand_product_device(Context* ctx, const GrandArgs& A, cudaStream_t st)
{
    if (A.width < 1 || A.width > 4 || A.n < 2 || (A.n & (A.n - 1)) || A.n >= (1ull << 31)) {
        set_last_error("grand_product: program width 1..4 and a power-of-two domain");
        return BBG_ERR_ARG;
    }
    const unsigned lg = log2u(A.n);
    // w_n^i: every (N / n)-th entry of the largest root table we already hold, or this size's own
    const fr* roots = nullptr;
    unsigned stride_log = 0;
    int rc = ntt_root_table_at_least(ctx, lg, (const void**)&roots, &stride_log, st);
    if (rc) return rc;
    const size_t m = (A.n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if ((rc = ctx->poly_tmp.reserve((2 * A.n + 2 * m) * sizeof(fr)))) return rc;
    fr* num = (fr*)ctx->poly_tmp.p;
    fr* den = num + A.n;
    fr* tot_num = den + A.n;
    fr* tot_den = tot_num + m;
    GrandParams P;
    memset(&P, 0, sizeof(P));
    for (unsigned k = 0; k < A.width; ++k) {
        P.wires[k] = (const fr*)A.d_wires[k];
        P.sigmas[k] = (const fr*)A.d_sigmas[k];
    }
    P.roots = roots;
    P.root_stride_log = stride_log;
    P.n = (uint32_t)A.n;
    P.width = A.width;
    P.beta = dev_fr(hf::reduce(A.beta));
    P.gamma = dev_fr(hf::reduce(A.gamma));
    for (int k = 0; k < 3; ++k) P.coset_gen[k] = dev_fr(hf::from_u64(5 + k));
    k_perm_terms<<<div_up(A.n, 128), 128, 0, st>>>(P, num, den);
    k_prod_chunks<<<dim3(div_up(m, 128), 2), 128, 0, st>>>(num, den, (uint32_t)A.n, tot_num, tot_den);
    k_prod_top<<<dim3(1, 2), SCAN_TOP_THREADS, 0, st>>>(tot_num, tot_den, (uint32_t)m);
    k_perm_finish<<<div_up(m, 128), 128, 0, st>>>(num, den, (uint32_t)A.n, tot_num, tot_den, (fr*)A.d_z);
    ctx->launches += 4;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

// out[k] = sum_i coeffs_k[i] z_k^i, k < count <= EVAL_BATCH_MAX; one launch for the whole batch
int poly_evaluate_batch_device(Context* ctx, const void* const* d_coeffs, const size_t* n, const void* zs, size_t count, void* d_out, cudaStream_t st)
{
    if (count == 0) return BBG_OK;
    if (count > EVAL_BATCH_MAX) {
        set_last_error("evaluate: at most " + std::to_string(EVAL_BATCH_MAX) + " polynomials per batch");
        return BBG_ERR_ARG;
    }
    EvalBatchParams P;
    memset(&P, 0, sizeof(P));
    size_t n_max = 1;
    for (size_t k = 0; k < count; ++k) {
        if (n[k] >= (1ull << 32)) {
            set_last_error("evaluate: too many coefficients");
            return BBG_ERR_ARG;
        }
        n_max = std::max(n_max, n[k]);
        P.coeffs[k] = (const fr*)d_coeffs[k];
        P.n[k] = (uint32_t)n[k];
        const hf::Fr z = hf::reduce(hf::load((const char*)zs + 32 * k));
        hf::Fr zc = z;
        for (int i = 0; i < 4; ++i) zc = hf::sqr(zc); // z^16
        P.z[k] = dev_fr(z);
        P.z_chunk[k] = dev_fr(hf::reduce(zc));
    }
    static_assert(EVAL_CHUNK == 16, "z^EVAL_CHUNK is computed by four squarings");
    const size_t threads = (n_max + EVAL_CHUNK - 1) / EVAL_CHUNK;
    const unsigned blocks = std::max(1u, div_up(threads, EVAL_THREADS));
    int rc = ctx->poly_tmp.reserve((size_t)blocks * count * sizeof(fr));
    if (rc) return rc;
    k_eval_partial<<<dim3(blocks, (unsigned)count), EVAL_THREADS, 0, st>>>(P, (fr*)ctx->poly_tmp.p);
    k_eval_final<<<(unsigned)count, EVAL_THREADS, 0, st>>>((const fr*)ctx->poly_tmp.p, blocks, (fr*)d_out);
    ctx->launches += 2;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

int poly_evaluate_device(Context* ctx, const void* d_coeffs, size_t n, const hf::Fr& z, void* d_out, cudaStream_t st)
{
    return poly_evaluate_batch_device(ctx, &d_coeffs, &n, z.d, 1, d_out, st);
}

int poly_opening_device(Context* ctx, const void* d_src, size_t n_in, size_t n_out, const hf::Fr& z, void* d_dest, void* d_f_at_z, cudaStream_t st)
{
    if (n_in == 0 || n_in >= (1ull << 32) || n_out > n_in) {
        set_last_error("compute_opening_polynomial: bad sizes");
        return BBG_ERR_ARG;
    }
    const size_t m = (n_in + OPEN_CHUNK - 1) / OPEN_CHUNK;
    int rc = ctx->poly_tmp.reserve(m * sizeof(fr));
    if (rc) return rc;
    fr* tot = (fr*)ctx->poly_tmp.p;
    hf::Fr zc = z;
    for (int i = 0; i < 4; ++i) zc = hf::sqr(zc);
    static_assert(OPEN_CHUNK == 16, "z^OPEN_CHUNK is computed by four squarings");
    const fr zd = dev_fr(hf::reduce(z));
    k_open_chunks<<<div_up(m, 128), 128, 0, st>>>((const fr*)d_src, (uint32_t)n_in, zd, tot);
    k_open_top<<<1, OPEN_TOP_THREADS, 0, st>>>(tot, (uint32_t)m, dev_fr(hf::reduce(zc)));
    k_open_apply<<<div_up(m, 128), 128, 0, st>>>((const fr*)d_src, (uint32_t)n_in, (uint32_t)n_out, zd, tot, (fr*)d_dest, (fr*)d_f_at_z);
    ctx->launches += 3;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

int poly_linear_combination_device(Context* ctx, void* d_dest, const void* d_base, const void* const* d_polys, const void* scalars, size_t count,
                                   size_t n, cudaStream_t st)
{
    if (count > LINCOMB_MAX || n >= (1ull << 32)) {
        set_last_error("linear_combination: at most " + std::to_string(LINCOMB_MAX) + " terms");
        return BBG_ERR_ARG;
    }
    if (n == 0) return BBG_OK;
    LinCombParams P;
    memset(&P, 0, sizeof(P));
    P.dest = (fr*)d_dest;
    P.base = (const fr*)d_base;
    P.n = (uint32_t)n;
    P.count = (uint32_t)count;
    for (size_t k = 0; k < count; ++k) {
        P.polys[k] = (const fr*)d_polys[k];
        P.scalars[k] = dev_fr(hf::reduce(hf::load((const char*)scalars + 32 * k)));
    }
    k_linear_combination<<<div_up(n, 128), 128, 0, st>>>(P);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

int poly_copy_pad_device(Context* ctx, const void* d_src, void* d_dst, size_t n, size_t total, cudaStream_t st)
{
    if (total == 0) return BBG_OK;
    k_copy_pad<<<div_up(total * 2, 256), 256, 0, st>>>((const uint4*)d_src, (uint4*)d_dst, n * 2, total * 2);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

} // namespace bbg
