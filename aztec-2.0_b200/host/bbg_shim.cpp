// bbg_shim.cpp -- barretenberg's own C++ entry points for the prover hot path, implemented on libbbg's C-ABI.
//
// Compile this ONE file against barretenberg's headers (-I <barretenberg>/src/aztec) and link it, with
// libbbg.so, in place of the definitions it replaces.  Callers (plonk::work_queue::process_queue,
// composer key generation, the verifier's small MSMs, Pippenger-backed reference strings) are untouched:
// same namespaces, same signatures, same g1::affine_element / fr memory layout, same exceptions.
//
//   replaced definition                                            reference location
//   scalar_multiplication::pippenger                               bb/ecc/curves/bn254/scalar_multiplication/scalar_multiplication.cpp:853-906
//   scalar_multiplication::pippenger_unsafe                        ... :923-929
//   scalar_multiplication::Pippenger::{ctor x2, dtor, pippenger_unsafe}   bb/ecc/curves/bn254/scalar_multiplication/pippenger.cpp:7-36
//   polynomial_arithmetic::fft / ifft / *_with_constant            bb/polynomials/polynomial_arithmetic.cpp:374-393, 471-478
//   polynomial_arithmetic::coset_fft (both) / coset_ifft / coset_fft_with_constant / coset_fft_with_generator_shift
//                                                                  ... :395-469, 480-484
//
// Everything else in those translation units (evaluate, divide_by_pseudo_vanishing_polynomial, the
// CPU-side generate_pippenger_point_table, ...) keeps its reference definition: INTEGRATION.md shows the
// two ways to link (objcopy --weaken-symbol on the reference objects, as oracle/Makefile's `ref_gpu`
// target does, or excluding the functions at source level).
#include <map>
#include <mutex>
#include <string>

#include <common/mem.hpp>
#include <common/throw_or_abort.hpp>
#include <ecc/curves/bn254/scalar_multiplication/pippenger.hpp>
#include <ecc/curves/bn254/scalar_multiplication/scalar_multiplication.hpp>
#include <polynomials/polynomial_arithmetic.hpp>

#include "../../include/bbg.h"

namespace {

void check(int rc)
{
    if (rc != BBG_OK) {
        throw_or_abort(std::string("libbbg: ") + bbg_last_error());
    }
}

// Pippenger's header fixes its data members (monomials_, num_points_), so the device handle of each object
// lives beside it, keyed by the object's address.
std::mutex g_mu;
std::map<const void*, void*> g_handles;

void* handle_of(const void* obj)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_handles.find(obj);
    return it == g_handles.end() ? nullptr : it->second;
}

} // namespace

namespace barretenberg {
namespace scalar_multiplication {

// `points` is the 2n-entry interleaved table of ProverReferenceString::get_monomials(); libbbg uses the resident
// device copy when the pointer lies inside a table it knows (bbg_pippenger_bind_host_table below, or
// BBG_AUTO_ADOPT=1), otherwise it uploads the even entries for this call.  The runtime state has no device
// counterpart; callers keep passing theirs.
g1::element pippenger(fr* scalars, g1::affine_element* points, const size_t num_points, pippenger_runtime_state&, bool handle_edge_cases)
{
    g1::element result;
    check(bbg_pippenger(scalars, points, num_points, handle_edge_cases ? 1 : 0, &result));
    return result;
}

g1::element pippenger_unsafe(fr* scalars, g1::affine_element* points, const size_t num_initial_points, pippenger_runtime_state&)
{
    g1::element result;
    check(bbg_pippenger(scalars, points, num_initial_points, 0, &result));
    return result;
}

// The transcript is decoded (byte swap + to-Montgomery) and the 2n table built ON THE DEVICE; the host copy that
// get_point_table() must return is read back once, so it is byte-identical to what the CPU path would hold.
Pippenger::Pippenger(uint8_t const* points, size_t num_points)
    : monomials_(point_table_alloc<g1::affine_element>(num_points))
    , num_points_(num_points)
{
    void* h = bbg_new_pippenger(points, num_points);
    if (h == nullptr) {
        throw_or_abort(std::string("libbbg: ") + bbg_last_error());
    }
    check(bbg_pippenger_get_point_table(h, monomials_));
    check(bbg_pippenger_bind_host_table(h, monomials_));
    std::lock_guard<std::mutex> lk(g_mu);
    g_handles[this] = h;
}

Pippenger::Pippenger(std::string const& path, size_t num_points)
    : monomials_(point_table_alloc<g1::affine_element>(num_points))
    , num_points_(num_points)
{
    void* h = bbg_new_pippenger_from_path(path.c_str(), num_points);
    if (h == nullptr) {
        throw_or_abort(std::string("libbbg: ") + bbg_last_error()); // includes io.cpp's "Is your srs large enough?"
    }
    check(bbg_pippenger_get_point_table(h, monomials_));
    check(bbg_pippenger_bind_host_table(h, monomials_));
    std::lock_guard<std::mutex> lk(g_mu);
    g_handles[this] = h;
}

Pippenger::~Pippenger()
{
    void* h = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_handles.find(this);
        if (it != g_handles.end()) {
            h = it->second;
            g_handles.erase(it);
        }
    }
    bbg_delete_pippenger(h);
    aligned_free(monomials_);
}

g1::element Pippenger::pippenger_unsafe(fr* scalars, size_t from, size_t range)
{
    g1::element result;
    check(bbg_pippenger_unsafe(handle_of(this), scalars, from, range, &result));
    return result;
}

} // namespace scalar_multiplication

namespace polynomial_arithmetic {

void fft(fr* coeffs, const evaluation_domain& domain)
{
    check(bbg_ntt(coeffs, domain.size, BBG_FFT, domain.generator_size, nullptr));
}
void fft_with_constant(fr* coeffs, const evaluation_domain& domain, const fr& value)
{
    check(bbg_ntt(coeffs, domain.size, BBG_FFT_WITH_CONSTANT, domain.generator_size, &value));
}
void ifft(fr* coeffs, const evaluation_domain& domain)
{
    check(bbg_ntt(coeffs, domain.size, BBG_IFFT, domain.generator_size, nullptr));
}
void ifft_with_constant(fr* coeffs, const evaluation_domain& domain, const fr& value)
{
    check(bbg_ntt(coeffs, domain.size, BBG_IFFT_WITH_CONSTANT, domain.generator_size, &value));
}
void coset_fft(fr* coeffs, const evaluation_domain& domain)
{
    check(bbg_ntt(coeffs, domain.size, BBG_COSET_FFT, domain.generator_size, nullptr));
}
void coset_fft(fr* coeffs, const evaluation_domain& small_domain, const evaluation_domain&, const size_t domain_extension)
{
    check(bbg_coset_fft_ext(coeffs, small_domain.size, domain_extension));
}
void coset_fft_with_constant(fr* coeffs, const evaluation_domain& domain, const fr& constant)
{
    check(bbg_ntt(coeffs, domain.size, BBG_COSET_FFT_WITH_CONSTANT, domain.generator_size, &constant));
}
void coset_fft_with_generator_shift(fr* coeffs, const evaluation_domain& domain, const fr& constant)
{
    check(bbg_ntt(coeffs, domain.size, BBG_COSET_FFT_WITH_GENERATOR_SHIFT, domain.generator_size, &constant));
}
void coset_ifft(fr* coeffs, const evaluation_domain& domain)
{
    // the result stays in the array's device mirror only if the mirror is already ahead of host memory (the prover shim's
    // quotient chain left it there: widgets -> divide_by_pseudo_vanishing_polynomial -> this transform -> commitments)
    check(bbg_ntt_ex(coeffs, domain.size, BBG_COSET_IFFT, domain.generator_size, nullptr, BBG_KEEP_IF_AHEAD));
}

} // namespace polynomial_arithmetic
} // namespace barretenberg
