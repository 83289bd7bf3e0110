// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" wrapper around the UNMODIFIED reference implementation
// (barretenberg, /root/reference/barretenberg/src/aztec) so that tests/, smoke()
// and bench.py's cpu_baseline / --impl reference arms can call the reference's
// own CPU path through ctypes.  This file is ours; it only *includes* reference
// headers and is linked against reference translation units compiled where they
// lie (see oracle/Makefile).  Nothing under aztec-2.0_b200/ may link it.
//
// Reference entry points wrapped here:
//   ecc/curves/bn254/scalar_multiplication/scalar_multiplication.hpp:94,139-148
//   ecc/curves/bn254/scalar_multiplication/pippenger.hpp:35-52
//   polynomials/polynomial_arithmetic.hpp:23-39
//   polynomials/evaluation_domain.hpp:6-59
//   srs/io.hpp:10-18
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <ecc/curves/bn254/fq.hpp>
#include <ecc/curves/bn254/fr.hpp>
#include <ecc/curves/bn254/g1.hpp>
#include <ecc/curves/bn254/scalar_multiplication/pippenger.hpp>
#include <ecc/curves/bn254/scalar_multiplication/scalar_multiplication.hpp>
#include <numeric/random/engine.hpp>
#include <polynomials/evaluation_domain.hpp>
#include <polynomials/polynomial_arithmetic.hpp>
#include <srs/io.hpp>
// header-only gate kernels of the TurboPLONK widgets (templates over Field / Getters; nothing from plonk is linked)
#include <plonk/proof_system/widgets/transition_widgets/turbo_arithmetic_widget.hpp>
#include <plonk/proof_system/widgets/transition_widgets/turbo_fixed_base_widget.hpp>
#include <plonk/proof_system/widgets/transition_widgets/turbo_logic_widget.hpp>
#include <plonk/proof_system/widgets/transition_widgets/turbo_range_widget.hpp>

#ifndef NO_MULTITHREADING
#include <omp.h>
#endif

using namespace barretenberg;

namespace {
template <typename F> void field_binop(int op, const void* a, const void* b, void* out)
{
    F x = *reinterpret_cast<const F*>(a);
    F y = b ? *reinterpret_cast<const F*>(b) : F::zero();
    F r;
    switch (op) {
    case 0: r = x * y; break;
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = x.sqr(); break;
    case 4: r = x.to_montgomery_form(); break;
    case 5: r = x.from_montgomery_form(); break;
    case 6: r = x.invert(); break;
    case 7: r = x.reduce_once(); break;
    case 8: r = -x; break;
    default: r = F::zero();
    }
    *reinterpret_cast<F*>(out) = r;
}
} // namespace

extern "C" {

int ref_num_threads()
{
#ifndef NO_MULTITHREADING
    return (int)max_threads::compute_num_threads();
#else
    return 1;
#endif
}

void ref_set_num_threads(int n)
{
#ifndef NO_MULTITHREADING
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// ---- L0: field ops on raw 4xu64 limbs (Montgomery form in/out unless op says otherwise)
void ref_fr_op(int op, const void* a, const void* b, void* out) { field_binop<fr>(op, a, b, out); }
void ref_fq_op(int op, const void* a, const void* b, void* out) { field_binop<fq>(op, a, b, out); }

void ref_fr_pow(const void* a, uint64_t e, void* out)
{
    *reinterpret_cast<fr*>(out) = reinterpret_cast<const fr*>(a)->pow(e);
}

// k (Montgomery) -> k1 | k2<<128 (non-Montgomery), field.hpp:236-282 as used by scalar_multiplication.cpp:224-226
void ref_fr_split_endo(const void* k, void* k1_k2)
{
    fr x = reinterpret_cast<const fr*>(k)->from_montgomery_form();
    fr k1{ 0, 0, 0, 0 }, k2{ 0, 0, 0, 0 };
    fr::split_into_endomorphism_scalars(x, k1, k2);
    uint64_t* o = reinterpret_cast<uint64_t*>(k1_k2);
    o[0] = k1.data[0]; o[1] = k1.data[1]; o[2] = k2.data[0]; o[3] = k2.data[1];
}

void ref_fr_constants(void* root_of_unity_28, void* coset_generator0, void* beta)
{
    *reinterpret_cast<fr*>(root_of_unity_28) = fr::get_root_of_unity(28);
    *reinterpret_cast<fr*>(coset_generator0) = fr::coset_generator(0);
    *reinterpret_cast<fr*>(beta) = fr::beta();
}

void ref_fq_constants(void* beta, void* g1_b, void* one_x, void* one_y)
{
    *reinterpret_cast<fq*>(beta) = fq::beta();
    *reinterpret_cast<fq*>(g1_b) = g1::curve_b;
    *reinterpret_cast<fq*>(one_x) = g1::affine_one.x;
    *reinterpret_cast<fq*>(one_y) = g1::affine_one.y;
}

// deterministic reference RNG stream (numeric/random/engine.cpp:124-137, field_impl.hpp:505-515);
// used only to reproduce SURVEY.md Appendix B and prove the staging is faithful.
void ref_debug_random_frs(void* out, size_t n)
{
    auto& e = numeric::random::get_debug_engine(true);
    fr* o = reinterpret_cast<fr*>(out);
    for (size_t i = 0; i < n; ++i) {
        o[i] = fr::random_element(&e);
    }
}

// ---- L0: group ops. Jacobian = 96 B {x,y,z}; affine = 64 B {x,y}
void ref_g1_mixed_add(const void* jac, const void* aff, void* out)
{
    g1::element r = *reinterpret_cast<const g1::element*>(jac);
    r += *reinterpret_cast<const g1::affine_element*>(aff);
    *reinterpret_cast<g1::element*>(out) = r;
}
void ref_g1_add(const void* a, const void* b, void* out)
{
    g1::element r = *reinterpret_cast<const g1::element*>(a);
    r += *reinterpret_cast<const g1::element*>(b);
    *reinterpret_cast<g1::element*>(out) = r;
}
void ref_g1_dbl(const void* a, void* out)
{
    g1::element r = *reinterpret_cast<const g1::element*>(a);
    r.self_dbl();
    *reinterpret_cast<g1::element*>(out) = r;
}
void ref_g1_set_infinity(void* out)
{
    g1::element r = g1::one;
    r.self_set_infinity();
    *reinterpret_cast<g1::element*>(out) = r;
}
// element -> affine (element_impl.hpp:51-68)
void ref_g1_to_affine(const void* jac, void* aff)
{
    *reinterpret_cast<g1::affine_element*>(aff) = g1::affine_element(*reinterpret_cast<const g1::element*>(jac));
}
// affine_element::to_buffer(): y || x big-endian canonical, infinity flag bit7 of byte 0 (affine_element.hpp:38-63)
void ref_g1_affine_to_buffer(const void* aff, uint8_t* buf64)
{
    g1::affine_element::serialize_to_buffer(*reinterpret_cast<const g1::affine_element*>(aff), buf64);
}
// the exact bytes work_queue.hpp:231,239 pushes into the transcript for an MSM result
void ref_g1_jac_to_buffer(const void* jac, uint8_t* buf64)
{
    g1::affine_element a(*reinterpret_cast<const g1::element*>(jac));
    g1::affine_element::serialize_to_buffer(a, buf64);
}
// scalar in Montgomery form; element * fr (mul_with_endomorphism, element_impl.hpp:592-663)
void ref_g1_mul(const void* aff, const void* scalar, void* out_jac)
{
    g1::element p(*reinterpret_cast<const g1::affine_element*>(aff));
    *reinterpret_cast<g1::element*>(out_jac) = p * *reinterpret_cast<const fr*>(scalar);
}
int ref_g1_on_curve(const void* aff) { return reinterpret_cast<const g1::affine_element*>(aff)->on_curve() ? 1 : 0; }
void ref_g1_hash_to_curve(uint64_t seed, void* aff)
{
    *reinterpret_cast<g1::affine_element*>(aff) = g1::affine_element::hash_to_curve(seed);
}
// c_bind.cpp:40-45
void ref_g1_sum(const void* jacs, size_t n, void* out)
{
    auto points = reinterpret_cast<const g1::element*>(jacs);
    g1::element r = g1::one;
    r.self_set_infinity();
    r = std::accumulate(points, points + n, r);
    *reinterpret_cast<g1::element*>(out) = r;
}

// ---- SRS (srs/io.cpp:134-162, 47-67)
int ref_read_transcript_g1(void* monomials, size_t degree, const char* dir)
{
    try {
        io::read_transcript_g1(reinterpret_cast<g1::affine_element*>(monomials), degree, std::string(dir));
    } catch (std::exception const&) {
        return 1;
    }
    return 0;
}
void ref_read_g1_elements_from_buffer(void* elements, const char* buffer, size_t buffer_size)
{
    io::read_g1_elements_from_buffer(reinterpret_cast<g1::affine_element*>(elements), buffer, buffer_size);
}

// ---- MSM
size_t ref_point_table_size(size_t n) { return scalar_multiplication::point_table_size(n); }

// scalar_multiplication.cpp:104-112 (table may alias points)
void ref_generate_pippenger_point_table(void* points, void* table, size_t n)
{
    scalar_multiplication::generate_pippenger_point_table(
        reinterpret_cast<g1::affine_element*>(points), reinterpret_cast<g1::affine_element*>(table), n);
}

void* ref_new_runtime_state(size_t n) { return new scalar_multiplication::pippenger_runtime_state(n); }
void ref_delete_runtime_state(void* s) { delete reinterpret_cast<scalar_multiplication::pippenger_runtime_state*>(s); }

// scalar_multiplication.cpp:853-906 / 923-929. table = 2n interleaved entries (+slack). state may be null.
int ref_pippenger(const void* scalars, void* table, size_t n, void* state, int unsafe, void* out_jac)
{
    try {
        scalar_multiplication::pippenger_runtime_state* st =
            reinterpret_cast<scalar_multiplication::pippenger_runtime_state*>(state);
        bool own = false;
        if (st == nullptr) {
            st = new scalar_multiplication::pippenger_runtime_state(n);
            own = true;
        }
        fr* s = const_cast<fr*>(reinterpret_cast<const fr*>(scalars));
        g1::affine_element* t = reinterpret_cast<g1::affine_element*>(table);
        g1::element r = unsafe ? scalar_multiplication::pippenger_unsafe(s, t, n, *st)
                               : scalar_multiplication::pippenger(s, t, n, *st);
        *reinterpret_cast<g1::element*>(out_jac) = r;
        if (own) {
            delete st;
        }
    } catch (std::exception const&) {
        return 1;
    }
    return 0;
}

// sum_i points[i] * scalars[i] the slow way (what scalar_multiplication.test.cpp:655-686 compares against)
void ref_naive_msm(const void* scalars, const void* affine_points, size_t n, size_t point_stride, void* out_jac)
{
    const fr* s = reinterpret_cast<const fr*>(scalars);
    const g1::affine_element* p = reinterpret_cast<const g1::affine_element*>(affine_points);
    g1::element acc = g1::one;
    acc.self_set_infinity();
    for (size_t i = 0; i < n; ++i) {
        g1::element t = g1::element(p[i * point_stride]) * s[i];
        acc += t;
    }
    *reinterpret_cast<g1::element*>(out_jac) = acc;
}

// ---- Pippenger class (pippenger.hpp:35-52): SRS owner + MSM over a point range
void* ref_new_pippenger_from_path(const char* dir, size_t num_points)
{
    try {
        return new scalar_multiplication::Pippenger(std::string(dir), num_points);
    } catch (std::exception const&) {
        return nullptr;
    }
}
void ref_delete_pippenger(void* p) { delete reinterpret_cast<scalar_multiplication::Pippenger*>(p); }
size_t ref_pippenger_num_points(void* p) { return reinterpret_cast<scalar_multiplication::Pippenger*>(p)->get_num_points(); }
void ref_pippenger_copy_table(void* p, void* out2n)
{
    auto* pp = reinterpret_cast<scalar_multiplication::Pippenger*>(p);
    std::memcpy(out2n, pp->get_point_table(), pp->get_num_points() * 2 * sizeof(g1::affine_element));
}
int ref_pippenger_class_unsafe(void* p, const void* scalars, size_t from, size_t range, void* out_jac)
{
    try {
        g1::element r = reinterpret_cast<scalar_multiplication::Pippenger*>(p)->pippenger_unsafe(
            const_cast<fr*>(reinterpret_cast<const fr*>(scalars)), from, range);
        *reinterpret_cast<g1::element*>(out_jac) = r;
    } catch (std::exception const&) {
        return 1;
    }
    return 0;
}

// ---- NTT
// kind: 0 fft, 1 ifft, 2 coset_fft, 3 coset_ifft, 4 fft_with_constant, 5 ifft_with_constant,
//       6 coset_fft_with_constant, 7 coset_fft_with_generator_shift   (polynomial_arithmetic.cpp:374-484)
struct ref_domain {
    evaluation_domain d;
    ref_domain(size_t n, size_t gs)
        : d(n, gs)
    {
        d.compute_lookup_table();
    }
};
void* ref_new_domain(size_t n, size_t generator_size) { return new ref_domain(n, generator_size); }
void ref_delete_domain(void* d) { delete reinterpret_cast<ref_domain*>(d); }
// root, root_inverse, domain, domain_inverse, generator, generator_inverse (6 x 32 B)
void ref_domain_constants(void* dom, void* out6)
{
    const evaluation_domain& d = reinterpret_cast<ref_domain*>(dom)->d;
    fr* o = reinterpret_cast<fr*>(out6);
    o[0] = d.root; o[1] = d.root_inverse; o[2] = d.domain; o[3] = d.domain_inverse;
    o[4] = d.generator; o[5] = d.generator_inverse;
}
void ref_ntt(void* dom, int kind, void* coeffs, const void* constant)
{
    const evaluation_domain& d = reinterpret_cast<ref_domain*>(dom)->d;
    fr* c = reinterpret_cast<fr*>(coeffs);
    const fr k = constant ? *reinterpret_cast<const fr*>(constant) : fr::one();
    switch (kind) {
    case 0: polynomial_arithmetic::fft(c, d); break;
    case 1: polynomial_arithmetic::ifft(c, d); break;
    case 2: polynomial_arithmetic::coset_fft(c, d); break;
    case 3: polynomial_arithmetic::coset_ifft(c, d); break;
    case 4: polynomial_arithmetic::fft_with_constant(c, d, k); break;
    case 5: polynomial_arithmetic::ifft_with_constant(c, d, k); break;
    case 6: polynomial_arithmetic::coset_fft_with_constant(c, d, k); break;
    case 7: polynomial_arithmetic::coset_fft_with_generator_shift(c, d, k); break;
    default: break;
    }
}
// coset_fft(coeffs, small, large, ext) (polynomial_arithmetic.cpp:401-456); coeffs has ext*n entries
void ref_coset_fft_ext(void* small_dom, void* large_dom, void* coeffs, size_t ext)
{
    polynomial_arithmetic::coset_fft(reinterpret_cast<fr*>(coeffs),
                                     reinterpret_cast<ref_domain*>(small_dom)->d,
                                     reinterpret_cast<ref_domain*>(large_dom)->d,
                                     ext);
}
// fr evaluate(coeffs, z, n) -- Horner, used by the reference's fft_with_small_degree test
void ref_evaluate(const void* coeffs, const void* z, size_t n, void* out)
{
    *reinterpret_cast<fr*>(out) =
        polynomial_arithmetic::evaluate(reinterpret_cast<const fr*>(coeffs), *reinterpret_cast<const fr*>(z), n);
}


// ---- quotient-stage pointwise functions
void ref_divide_by_pseudo_vanishing_polynomial(void* evals, void* small_dom, void* large_dom, size_t num_roots_cut)
{
    polynomial_arithmetic::divide_by_pseudo_vanishing_polynomial(reinterpret_cast<fr*>(evals), reinterpret_cast<ref_domain*>(small_dom)->d,
                                                                 reinterpret_cast<ref_domain*>(large_dom)->d, num_roots_cut);
}
void ref_compute_lagrange_polynomial_fft(void* l1, void* small_dom, void* large_dom)
{
    polynomial_arithmetic::compute_lagrange_polynomial_fft(reinterpret_cast<fr*>(l1), reinterpret_cast<ref_domain*>(small_dom)->d,
                                                           reinterpret_cast<ref_domain*>(large_dom)->d);
}
// fr compute_kate_opening_coefficients(src, dest, z, n) (polynomial_arithmetic.cpp:727-751); returns F(z) in f_out
void ref_compute_kate_opening_coefficients(const void* src, void* dest, const void* z, size_t n, void* f_out)
{
    *reinterpret_cast<fr*>(f_out) = polynomial_arithmetic::compute_kate_opening_coefficients(
        reinterpret_cast<const fr*>(src), reinterpret_cast<fr*>(dest), *reinterpret_cast<const fr*>(z), n);
}

} // extern "C"

// The reference's gate kernels instantiated over raw arrays: exactly the three calls per evaluation point that
// TransitionWidget::compute_quotient_contribution makes (transition_widget.hpp:293-307), with FFTGetter's indexing
// ((index + 4) & block_mask for the shifted wires, :160-169).
namespace {
typedef waffle::widget::containers::poly_ptr_array<fr> raw_polys;
struct RawGetters {
    template <bool use_shifted_evaluation, waffle::PolynomialIndex id>
    inline static const fr& get_polynomial(const raw_polys& polynomials, const size_t index = 0)
    {
        if constexpr (use_shifted_evaluation) {
            return polynomials.coefficients[id][(index + 4) & polynomials.block_mask];
        }
        return polynomials.coefficients[id][index];
    }
};
template <template <typename, typename, typename> typename KernelBase>
void run_turbo_kernel(raw_polys& polynomials, size_t n_large, const fr& alpha_base, const fr& alpha, fr* quotient)
{
    typedef KernelBase<fr, RawGetters, raw_polys> Kernel;
    constexpr size_t R = Kernel::num_independent_relations;
    waffle::widget::containers::challenge_array<fr, R> challenges{};
    challenges.elements[waffle::widget::ChallengeIndex::ALPHA] = alpha;
    challenges.alpha_powers[0] = alpha_base;
    for (size_t i = 1; i < R; ++i) challenges.alpha_powers[i] = challenges.alpha_powers[i - 1] * alpha;
    for (size_t i = 0; i < n_large; ++i) {
        waffle::widget::containers::coefficient_array<fr> linear_terms;
        Kernel::compute_linear_terms(polynomials, challenges, linear_terms, i);
        quotient[i] += Kernel::sum_linear_terms(polynomials, challenges, linear_terms, i);
        Kernel::compute_non_linear_terms(polynomials, challenges, quotient[i], i);
    }
}
} // namespace

extern "C" {
// kind: 0 arithmetic, 1 fixed base, 2 range, 3 logic; polys: MAX_NUM_POLYNOMIALS pointers indexed by waffle::PolynomialIndex
void ref_turbo_quotient(int kind, const void* const* polys, size_t n_large, const void* alpha_base, const void* alpha, void* quotient)
{
    raw_polys p;
    for (size_t k = 0; k < waffle::PolynomialIndex::MAX_NUM_POLYNOMIALS; ++k) {
        p.coefficients[k] = const_cast<fr*>(reinterpret_cast<const fr*>(polys[k]));
    }
    p.block_mask = n_large - 1;
    const fr a0 = *reinterpret_cast<const fr*>(alpha_base), a = *reinterpret_cast<const fr*>(alpha);
    fr* q = reinterpret_cast<fr*>(quotient);
    switch (kind) {
    case 0: run_turbo_kernel<waffle::widget::TurboArithmeticKernel>(p, n_large, a0, a, q); break;
    case 1: run_turbo_kernel<waffle::widget::TurboFixedBaseKernel>(p, n_large, a0, a, q); break;
    case 2: run_turbo_kernel<waffle::widget::TurboRangeKernel>(p, n_large, a0, a, q); break;
    default: run_turbo_kernel<waffle::widget::TurboLogicKernel>(p, n_large, a0, a, q); break;
    }
}
} // extern "C"
