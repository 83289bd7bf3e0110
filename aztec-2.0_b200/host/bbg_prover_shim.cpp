// bbg_prover_shim.cpp -- the PLONK prover's own callers of the hot path, re-pointed at libbbg so that a proof's
// polynomials stay in HBM between the NTTs, the pointwise quotient stage and the commitment MSMs.
//
// Link this file NEXT TO bbg_shim.cpp (it needs the plonk headers; bbg_shim.cpp alone still gives the L1 drop-in).  Every
// function below has the reference's own symbol, so the unmodified prover objects call it instead of their inline /
// template-instantiated (weak) copies; nothing in barretenberg is edited.
//
//   replaced definition                                                          reference location (bb/plonk/proof_system/)
//   work_queue::process_queue                                                    prover/work_queue.hpp:208-282
//   ProverPermutationWidget<4, false>::compute_round_commitments  (round 3)      widgets/random_widgets/permutation_widget_impl.hpp:48-313
//   ProverPermutationWidget<4, false>::compute_quotient_contribution             ... :317-437
//   TransitionWidget<fr, {turbo,unrolled_turbo}_settings, Turbo{Arithmetic,FixedBase,Range,Logic}Kernel>::compute_quotient_contribution
//                                                                                widgets/transition_widgets/transition_widget.hpp:293-307
//   polynomial_arithmetic::divide_by_pseudo_vanishing_polynomial                 bb/polynomials/polynomial_arithmetic.cpp:628-725
//   polynomial_arithmetic::evaluate                                              bb/polynomials/polynomial_arithmetic.cpp:507-538
//   polynomial_arithmetic::compute_lagrange_polynomial_fft  (key generation)     bb/polynomials/polynomial_arithmetic.cpp:546-626
//   KateCommitmentScheme<{turbo,unrolled_turbo}_settings>::batch_open            commitment_scheme/kate_commitment_scheme.cpp:133-237
//   KateCommitmentScheme<{turbo,unrolled_turbo}_settings>::add_opening_evaluations_to_transcript   ... :373-436
//
// What changes for a TurboPLONK proof (program width 4; every widget it uses is replaced here):
//   * work items are submitted in batches: the four wire commitments / four quotient commitments of a round go to
//     bbg_pippenger_batch (two streams, tails overlapped);
//   * resident polynomials are switched on: a wire polynomial is uploaded once for ifft -> commitment -> coset FFT;
//   * the 4n-point coset FFTs of the wires and of z, the whole quotient computation (permutation + four gate widgets +
//     division by Z_H*) and the grand product z run on the device and their outputs are NOT copied back: key->wire_ffts
//     and key->quotient_large in host memory are stale until quotient_large.coset_ifft() brings t(X) home.
//     Nothing else in the reference reads those arrays in between (prover.cpp:275-363).
// A prover that uses a widget this file does not replace (standard PLONK, MiMC, ...) never reaches the deferred paths:
// its FFT work items are written back like in round 1 (see `turbo_key`).
// BBG_PROVER_SHIM=0 in the environment makes process_queue run the reference's item-by-item logic again (the widget and
// division replacements keep running on the device, with their results written back to host memory every call);
// BBG_RESIDENT=0 keeps resident polynomials off (every call then uploads its inputs and downloads its outputs).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#define private public // work_queue keeps key / witness / transcript private and has no accessors
#include <plonk/proof_system/prover/work_queue.hpp>
#undef private
#include <common/throw_or_abort.hpp>
#include <plonk/proof_system/commitment_scheme/kate_commitment_scheme.hpp>
#include <plonk/proof_system/public_inputs/public_inputs.hpp>
#include <plonk/proof_system/types/program_settings.hpp>
#include <plonk/proof_system/types/prover_settings.hpp>
#include <plonk/proof_system/widgets/random_widgets/permutation_widget.hpp>
#include <plonk/proof_system/widgets/transition_widgets/turbo_arithmetic_widget.hpp>
#include <plonk/proof_system/widgets/transition_widgets/turbo_fixed_base_widget.hpp>
#include <plonk/proof_system/widgets/transition_widgets/turbo_logic_widget.hpp>
#include <plonk/proof_system/widgets/transition_widgets/turbo_range_widget.hpp>
#include <polynomials/polynomial_arithmetic.hpp>

#include "../../include/bbg.h"

using barretenberg::fr;
using barretenberg::g1;
using barretenberg::polynomial;

namespace {

void check(int rc)
{
    if (rc != BBG_OK) {
        throw_or_abort(std::string("libbbg: ") + bbg_last_error());
    }
}

bool shim_enabled()
{
    static const bool on = [] {
        const char* v = getenv("BBG_PROVER_SHIM");
        return !(v && *v == '0');
    }();
    return on;
}

// resident polynomials: on for the life of the process once the prover shim is in use
void ensure_resident()
{
    static bool done = false;
    if (!done) {
        const char* v = getenv("BBG_RESIDENT");
        if (!(v && *v == '0')) check(bbg_resident_mode(1));
        done = true;
    }
}

// BBG_SHIM_TRACE=1: one stderr line per replaced call (wall ms inside, ms since the previous replaced call returned = the
// reference CPU code in between)
struct Trace {
    const char* name;
    std::chrono::steady_clock::time_point t0;
    static std::chrono::steady_clock::time_point& last()
    {
        static std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
        return t;
    }
    static bool on()
    {
        static const bool v = [] {
            const char* e = getenv("BBG_SHIM_TRACE");
            return e && *e && *e != '0';
        }();
        return v;
    }
    explicit Trace(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
    ~Trace()
    {
        if (!on()) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "{\"shim\": \"%s\", \"ms\": %.3f, \"cpu_gap_before_ms\": %.3f}\n", name,
                std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t0 - last()).count());
        last() = t1;
    }
};

// a TurboPLONK key: width 4 and the four turbo selectors -- the configuration whose widgets are ALL replaced below
bool turbo_key(const waffle::proving_key* key)
{
    if (!key || key->wire_ffts.count("w_4_fft") == 0) return false;
    bool arith = false, fixed = false, range = false, logic = false, other = false;
    for (const auto& info : key->polynomial_manifest) {
        switch (info.index) {
        case waffle::PolynomialIndex::Q_ARITHMETIC_SELECTOR: arith = true; break;
        case waffle::PolynomialIndex::Q_FIXED_BASE_SELECTOR: fixed = true; break;
        case waffle::PolynomialIndex::Q_RANGE_SELECTOR: range = true; break;
        case waffle::PolynomialIndex::Q_LOGIC_SELECTOR: logic = true; break;
        case waffle::PolynomialIndex::Q_SORT_SELECTOR:
        case waffle::PolynomialIndex::TABLE_1:
        case waffle::PolynomialIndex::Q_MIMC_SELECTOR:
        case waffle::PolynomialIndex::Q_ELLIPTIC:
        case waffle::PolynomialIndex::S:
        case waffle::PolynomialIndex::Z_LOOKUP: other = true; break;
        default: break;
        }
    }
    return arith && fixed && range && logic && !other;
}

// the "_fft" polynomial of every manifest entry, indexed by PolynomialIndex (FFTGetter::get_fft_polynomials,
// transition_widget.hpp:131-158), null where the key has none
void fft_polynomial_table(waffle::proving_key* key, const void** table)
{
    for (size_t i = 0; i < BBG_NUM_POLYNOMIALS; ++i) table[i] = nullptr;
    for (const auto& info : key->polynomial_manifest) {
        const std::string label = std::string(info.polynomial_label) + "_fft";
        const fr* poly = nullptr;
        switch (info.source) {
        case waffle::PolynomialSource::WITNESS: poly = &key->wire_ffts.at(label)[0]; break;
        case waffle::PolynomialSource::SELECTOR: poly = &key->constraint_selector_ffts.at(label)[0]; break;
        case waffle::PolynomialSource::PERMUTATION: poly = &key->permutation_selector_ffts.at(label)[0]; break;
        }
        if ((size_t)info.index < BBG_NUM_POLYNOMIALS) table[info.index] = poly;
    }
}
static_assert((size_t)waffle::PolynomialIndex::MAX_NUM_POLYNOMIALS == BBG_NUM_POLYNOMIALS, "bbg.h polynomial table size");

// ---- the reference bodies, for BBG_PROVER_SHIM=0 and for configurations this file does not accelerate ----
void reference_process_item(waffle::work_queue* q, const waffle::work_queue::work_item& item)
{
    using WorkType = waffle::work_queue::WorkType;
    auto* key = q->key;
    auto* witness = q->witness;
    switch (item.work_type) {
    case WorkType::SCALAR_MULTIPLICATION: {
        const size_t n = key->small_domain.size + (item.constant == fr(1) ? 1 : 0);
        if (item.constant == fr(1)) {
            auto state = barretenberg::scalar_multiplication::pippenger_runtime_state(n);
            g1::affine_element r(barretenberg::scalar_multiplication::pippenger_unsafe(item.mul_scalars, key->reference_string->get_monomials(), n, state));
            q->transcript->add_element(item.tag, r.to_buffer());
        } else {
            g1::affine_element r(barretenberg::scalar_multiplication::pippenger_unsafe(item.mul_scalars, key->reference_string->get_monomials(), n,
                                                                                       key->pippenger_runtime_state));
            q->transcript->add_element(item.tag, r.to_buffer());
        }
        break;
    }
    case WorkType::SMALL_FFT: {
        const size_t n = key->n;
        polynomial& wire = witness->wires.at(item.tag);
        polynomial& wire_fft = key->wire_ffts.at(item.tag + "_fft");
        polynomial wire_copy(wire, n);
        wire_copy.coset_fft_with_generator_shift(key->small_domain, item.constant);
        for (size_t i = 0; i < n; ++i) wire_fft[4 * i + item.index] = wire_copy[i];
        wire_fft[4 * n + item.index] = wire_copy[0];
        break;
    }
    case WorkType::FFT: {
        polynomial& wire = witness->wires.at(item.tag);
        polynomial& wire_fft = key->wire_ffts.at(item.tag + "_fft");
        barretenberg::polynomial_arithmetic::copy_polynomial(&wire[0], &wire_fft[0], key->n, 4 * key->n + 4);
        wire_fft.coset_fft(key->large_domain);
        for (size_t k = 0; k < 4; ++k) wire_fft.add_lagrange_base_coefficient(wire_fft[k]);
        break;
    }
    case WorkType::IFFT: {
        witness->wires.at(item.tag).ifft(key->small_domain);
        break;
    }
    default: break;
    }
}

} // namespace

// ------------------------------------------------------------------------------------------------------------------
// work_queue::process_queue.  The member is defined inline in the reference header, so its replacement is a free
// function carrying the member's mangled name (Itanium ABI: `this` is the first integer argument).
// ------------------------------------------------------------------------------------------------------------------
extern "C" void bbg_shim_process_queue(waffle::work_queue* self) asm("_ZN6waffle10work_queue13process_queueEv");
extern "C" void bbg_shim_process_queue(waffle::work_queue* self)
{
    using WorkType = waffle::work_queue::WorkType;
    Trace trace("process_queue");
    auto& items = self->work_item_queue;
    auto* key = self->key;
    auto* witness = self->witness;
    if (!shim_enabled()) {
        for (const auto& item : items) reference_process_item(self, item);
        items = std::vector<waffle::work_queue::work_item>();
        return;
    }
    ensure_resident();
    const bool defer = turbo_key(key) && bbg_resident_mode(-1) == 1;
    const size_t n = key->small_domain.size;
    for (size_t i = 0; i < items.size();) {
        const auto& item = items[i];
        switch (item.work_type) {
        case WorkType::SCALAR_MULTIPLICATION: {
            // a run of commitments over the same bases and size: one batched call
            const bool plus_one = item.constant == fr(1);
            size_t j = i;
            std::vector<const void*> scalars;
            while (j < items.size() && items[j].work_type == WorkType::SCALAR_MULTIPLICATION && (items[j].constant == fr(1)) == plus_one) {
                scalars.push_back(items[j].mul_scalars);
                ++j;
            }
            const size_t count = j - i;
            const size_t num_points = n + (plus_one ? 1 : 0);
            std::vector<g1::element> results(count);
            g1::affine_element* monomials = key->reference_string->get_monomials();
            int rc = count > 1 ? bbg_pippenger_batch(scalars.data(), count, monomials, num_points, results.data()) : BBG_ERR_ARG;
            if (rc != BBG_OK) {
                // a single commitment, or bases libbbg has not adopted (MemReferenceString): one call each
                for (size_t k = 0; k < count; ++k) check(bbg_pippenger(scalars[k], monomials, num_points, 0, &results[k]));
            }
            for (size_t k = 0; k < count; ++k) {
                g1::affine_element r(results[k]);
                self->transcript->add_element(items[i + k].tag, r.to_buffer());
            }
            i = j;
            break;
        }
        case WorkType::IFFT: {
            // first work of a proof: the prover has just rewritten its wires and the Lagrange copies in wire_ffts
            // (prover.cpp:184-186) behind any mirror kept from the previous proof.  All IFFT items in a row go down together.
            std::vector<void*> cols;
            std::vector<const void*> copies;
            std::vector<polynomial*> polys;
            size_t j = i;
            while (j < items.size() && items[j].work_type == WorkType::IFFT && cols.size() < 16) {
                polynomial& wire = witness->wires.at(items[j].tag);
                polynomial& wire_fft = key->wire_ffts.at(items[j].tag + "_fft");
                check(bbg_resident_invalidate(&wire_fft[0], wire_fft.get_max_size() * sizeof(fr)));
                // ... and the wire itself: its mirror was left AHEAD of host memory by the previous proof (coefficients kept on
                // the device), so the library would trust it over the witness the prover has just written
                check(bbg_resident_invalidate(&wire[0], wire.get_max_size() * sizeof(fr)));
                // what polynomial::ifft does (polynomial.cpp:312-320), plus: the upload also seeds the mirror of the Lagrange
                // copy in wire_fft[0, n), which round 3's grand product reads
                if (n > wire.get_max_size()) wire.reserve(n);
                cols.push_back(&wire[0]);
                copies.push_back(&wire_fft[0]);
                polys.push_back(&wire);
                ++j;
            }
            check(bbg_wire_ifft_batch(cols.data(), n, copies.data(), cols.size(), defer ? BBG_KEEP_ON_DEVICE : 0));
            for (polynomial* w : polys) w->resize_unsafe(n);
            i = j;
            break;
        }
        case WorkType::FFT: {
            polynomial& wire = witness->wires.at(item.tag);
            polynomial& wire_fft = key->wire_ffts.at(item.tag + "_fft");
            wire_fft.resize_unsafe(4 * key->n + 4); // the size coset_fft + 4 x add_lagrange_base_coefficient leave behind
            check(bbg_wire_coset_fft(&wire[0], &wire_fft[0], key->n, 4, defer ? BBG_KEEP_ON_DEVICE : 0));
            ++i;
            break;
        }
        default:
            reference_process_item(self, item);
            ++i;
            break;
        }
    }
    items = std::vector<waffle::work_queue::work_item>();
}

// ------------------------------------------------------------------------------------------------------------------
// polynomial_arithmetic::divide_by_pseudo_vanishing_polynomial
// ------------------------------------------------------------------------------------------------------------------
namespace barretenberg {
namespace polynomial_arithmetic {
void divide_by_pseudo_vanishing_polynomial(fr* coeffs, const evaluation_domain& src_domain, const evaluation_domain& target_domain,
                                           const size_t num_roots_cut_out_of_vanishing_polynomial)
{
    Trace trace("divide_by_pseudo_vanishing_polynomial");
    // the result stays in the array's device mirror only if the mirror is already ahead of host memory (the widgets above
    // left it there); a caller that holds the data on the host gets it back on the host
    check(bbg_divide_by_pseudo_vanishing_polynomial(coeffs, src_domain.size, target_domain.size, (unsigned)num_roots_cut_out_of_vanishing_polynomial,
                                                    BBG_KEEP_IF_AHEAD));
}
} // namespace polynomial_arithmetic
} // namespace barretenberg

// ------------------------------------------------------------------------------------------------------------------
// polynomial_arithmetic::evaluate: the ~30 opening evaluations of round 5 (kate_commitment_scheme.cpp:373-436) and
// t(zeta) over 4n coefficients (prover.cpp:379) read polynomials whose mirrors are already on the device
// ------------------------------------------------------------------------------------------------------------------
namespace barretenberg {
namespace polynomial_arithmetic {
void compute_lagrange_polynomial_fft(fr* l_1_coefficients, const evaluation_domain& src_domain, const evaluation_domain& target_domain)
{
    Trace trace("compute_lagrange_polynomial_fft");
    check(bbg_compute_lagrange_polynomial_fft(l_1_coefficients, src_domain.size, target_domain.size));
}

fr evaluate(const fr* coeffs, const fr& z, const size_t n)
{
    Trace trace("evaluate");
    ensure_resident();
    fr result;
    check(bbg_evaluate(coeffs, n, &z, &result));
    return result;
}
} // namespace polynomial_arithmetic
} // namespace barretenberg

// ------------------------------------------------------------------------------------------------------------------
// KateCommitmentScheme::batch_open for program width 4.  The class is explicitly instantiated in the reference (extern
// template in its header), so like process_queue the replacement is a free function carrying the member's mangled name;
// the two std::shared_ptr arguments are passed by invisible reference (Itanium ABI), i.e. as pointers.
// ------------------------------------------------------------------------------------------------------------------
namespace {
template <typename settings>
void batch_open_width4(const transcript::StandardTranscript& transcript, waffle::work_queue& queue, const std::shared_ptr<waffle::proving_key>& input_key,
                       const std::shared_ptr<waffle::program_witness>& witness)
{
    static_assert(settings::program_width == 4, "the n + 1 coefficient of standard PLONK's t_high is not handled here");
    Trace trace("batch_open");
    using waffle::PolynomialSource;
    std::vector<const void*> at_zeta, at_zeta_omega;
    std::vector<fr> nu_zeta, nu_zeta_omega;
    // the same tuples, in the same order, as kate_commitment_scheme.cpp:160-211
    for (size_t i = 0; i < input_key->polynomial_manifest.size(); ++i) {
        const auto& info = input_key->polynomial_manifest[i];
        const std::string poly_label(info.polynomial_label);
        fr* poly = nullptr;
        switch (info.source) {
        case PolynomialSource::WITNESS: poly = &witness->wires.at(poly_label)[0]; break;
        case PolynomialSource::SELECTOR: poly = &input_key->constraint_selectors.at(poly_label)[0]; break;
        case PolynomialSource::PERMUTATION: poly = &input_key->permutation_selectors.at(poly_label)[0]; break;
        }
        if (!info.is_linearised || !settings::use_linearisation) {
            at_zeta.push_back(poly);
            nu_zeta.push_back(transcript.get_challenge_field_element_from_map("nu", poly_label));
        }
        if (info.requires_shifted_evaluation) {
            at_zeta_omega.push_back(poly);
            nu_zeta_omega.push_back(transcript.get_challenge_field_element_from_map("nu", poly_label + "_omega"));
        }
    }
    const fr zeta = transcript.get_challenge_field_element("z");
    const size_t n = input_key->small_domain.size;
    for (size_t i = 1; i < settings::program_width; ++i) {
        const size_t offset = i * n;
        at_zeta.push_back(&input_key->quotient_large[offset]);
        nu_zeta.push_back(zeta.pow(static_cast<uint64_t>(offset)));
    }
    if constexpr (settings::use_linearisation) {
        at_zeta.push_back(&input_key->linear_poly[0]);
        nu_zeta.push_back(transcript.get_challenge_field_element_from_map("nu", "r"));
    }
    barretenberg::polynomial& opening_poly = input_key->opening_poly;
    barretenberg::polynomial& shifted_opening_poly = input_key->shifted_opening_poly;
    ensure_resident();
    // both opening polynomials are only ever read again as MSM scalars (PI_Z, PI_Z_OMEGA): they stay on the device
    const unsigned keep = bbg_resident_mode(-1) == 1 ? BBG_KEEP_ON_DEVICE : 0;
    check(bbg_linear_combination(&opening_poly[0], &input_key->quotient_large[0], at_zeta.data(), nu_zeta.data(), at_zeta.size(), n, keep));
    check(bbg_linear_combination(&shifted_opening_poly[0], nullptr, at_zeta_omega.data(), nu_zeta_omega.data(), at_zeta_omega.size(), n, keep));
    const fr zeta_omega = zeta * input_key->small_domain.root;
    check(bbg_compute_opening_polynomial(&opening_poly[0], &opening_poly[0], &zeta, n, n, nullptr, keep));
    queue.add_to_queue({ waffle::work_queue::WorkType::SCALAR_MULTIPLICATION, &opening_poly[0], "PI_Z", fr(0), 0 });
    check(bbg_compute_opening_polynomial(&shifted_opening_poly[0], &shifted_opening_poly[0], &zeta_omega, n, n, nullptr, keep));
    queue.add_to_queue({ waffle::work_queue::WorkType::SCALAR_MULTIPLICATION, &shifted_opening_poly[0], "PI_Z_OMEGA", fr(0), 0 });
}

// every opening evaluation of round 5 in ONE device launch (the reference evaluates the ~30 polynomials one by one)
template <typename settings>
void opening_evaluations(transcript::StandardTranscript& transcript, const std::shared_ptr<waffle::proving_key>& input_key,
                         const std::shared_ptr<waffle::program_witness>& witness, bool in_lagrange_form)
{
    Trace trace("add_opening_evaluations_to_transcript");
    using waffle::PolynomialSource;
    const fr zeta = fr::serialize_from_buffer(transcript.get_challenge("z").begin());
    const fr shifted_z = zeta * input_key->small_domain.root;
    const size_t n = input_key->small_domain.size;
    std::vector<const void*> polys;
    std::vector<size_t> ns;
    std::vector<fr> zs;
    std::vector<std::string> labels;
    for (size_t i = 0; i < input_key->polynomial_manifest.size(); ++i) {
        const auto& info = input_key->polynomial_manifest[i];
        const std::string poly_label(info.polynomial_label);
        fr* poly = nullptr;
        switch (info.source) {
        case PolynomialSource::WITNESS: poly = &witness->wires.at(poly_label)[0]; break;
        case PolynomialSource::SELECTOR: poly = &input_key->constraint_selectors.at(poly_label)[0]; break;
        case PolynomialSource::PERMUTATION: poly = &input_key->permutation_selectors.at(poly_label)[0]; break;
        }
        if (!info.is_linearised || !settings::use_linearisation) {
            polys.push_back(poly);
            ns.push_back(n);
            zs.push_back(zeta);
            labels.push_back(poly_label);
        }
        if (info.requires_shifted_evaluation) {
            polys.push_back(poly);
            ns.push_back(n);
            // like the reference (kate_commitment_scheme.cpp:427-431): the Lagrange-form branch evaluates at zeta
            zs.push_back(in_lagrange_form ? zeta : shifted_z);
            labels.push_back(poly_label + "_omega");
        }
    }
    std::vector<fr> evals(polys.size());
    if (in_lagrange_form) {
        for (size_t k = 0; k < polys.size(); ++k) {
            evals[k] = barretenberg::polynomial_arithmetic::compute_barycentric_evaluation(const_cast<fr*>(static_cast<const fr*>(polys[k])), n, zs[k], input_key->small_domain);
        }
    } else {
        ensure_resident();
        for (size_t k = 0; k < polys.size(); k += 40) {
            const size_t m = std::min<size_t>(40, polys.size() - k);
            check(bbg_evaluate_batch(polys.data() + k, ns.data() + k, m, zs.data() + k, evals.data() + k));
        }
    }
    for (size_t k = 0; k < polys.size(); ++k) transcript.add_element(labels[k], evals[k].to_buffer());
}
} // namespace

#define BBG_KATE_EVALS(NAME, SETTINGS, MANGLED)                                                                                             \
    extern "C" void NAME(void* self, transcript::StandardTranscript* transcript, std::shared_ptr<waffle::proving_key>* key,                  \
                         std::shared_ptr<waffle::program_witness>* witness, bool in_lagrange_form) asm(MANGLED);                             \
    extern "C" void NAME(void*, transcript::StandardTranscript* transcript, std::shared_ptr<waffle::proving_key>* key,                       \
                         std::shared_ptr<waffle::program_witness>* witness, bool in_lagrange_form)                                           \
    {                                                                                                                                       \
        opening_evaluations<SETTINGS>(*transcript, *key, *witness, in_lagrange_form);                                                       \
    }
BBG_KATE_EVALS(bbg_shim_opening_evals_turbo, waffle::turbo_settings,
               "_ZN6waffle20KateCommitmentSchemeINS_14turbo_settingsEE37add_opening_evaluations_to_transcriptERN10transcript18StandardTranscriptESt10shared_ptrINS_11proving_keyEES6_INS_15program_witnessEEb")
BBG_KATE_EVALS(bbg_shim_opening_evals_unrolled_turbo, waffle::unrolled_turbo_settings,
               "_ZN6waffle20KateCommitmentSchemeINS_23unrolled_turbo_settingsEE37add_opening_evaluations_to_transcriptERN10transcript18StandardTranscriptESt10shared_ptrINS_11proving_keyEES6_INS_15program_witnessEEb")
#undef BBG_KATE_EVALS

extern "C" void bbg_shim_batch_open_unrolled_turbo(void* self, const transcript::StandardTranscript* transcript, waffle::work_queue* queue,
                                                   std::shared_ptr<waffle::proving_key>* key, std::shared_ptr<waffle::program_witness>* witness)
    asm("_ZN6waffle20KateCommitmentSchemeINS_23unrolled_turbo_settingsEE10batch_openERKN10transcript18StandardTranscriptERNS_10work_queueESt10shared_ptrINS_11proving_keyEES9_INS_15program_witnessEE");
extern "C" void bbg_shim_batch_open_unrolled_turbo(void*, const transcript::StandardTranscript* transcript, waffle::work_queue* queue,
                                                   std::shared_ptr<waffle::proving_key>* key, std::shared_ptr<waffle::program_witness>* witness)
{
    batch_open_width4<waffle::unrolled_turbo_settings>(*transcript, *queue, *key, *witness);
}
extern "C" void bbg_shim_batch_open_turbo(void* self, const transcript::StandardTranscript* transcript, waffle::work_queue* queue,
                                          std::shared_ptr<waffle::proving_key>* key, std::shared_ptr<waffle::program_witness>* witness)
    asm("_ZN6waffle20KateCommitmentSchemeINS_14turbo_settingsEE10batch_openERKN10transcript18StandardTranscriptERNS_10work_queueESt10shared_ptrINS_11proving_keyEES9_INS_15program_witnessEE");
extern "C" void bbg_shim_batch_open_turbo(void*, const transcript::StandardTranscript* transcript, waffle::work_queue* queue,
                                          std::shared_ptr<waffle::proving_key>* key, std::shared_ptr<waffle::program_witness>* witness)
{
    batch_open_width4<waffle::turbo_settings>(*transcript, *queue, *key, *witness);
}

// ------------------------------------------------------------------------------------------------------------------
// widgets
// ------------------------------------------------------------------------------------------------------------------
namespace waffle {

namespace {
template <class Settings, template <typename, typename, typename> typename KernelBase>
fr turbo_quotient(widget::TransitionWidget<fr, Settings, KernelBase>* self, int kind, const fr& alpha_base, const transcript::StandardTranscript& transcript)
{
    typedef widget::TransitionWidget<fr, Settings, KernelBase> W;
    Trace trace("turbo_quotient");
    auto* key = self->key;
    auto challenges = W::FFTGetter::get_challenges(transcript, alpha_base);
    const void* table[BBG_NUM_POLYNOMIALS];
    fft_polynomial_table(key, table);
    ensure_resident();
    check(bbg_turbo_quotient(kind, table, key->large_domain.size, &challenges.alpha_powers[0], &challenges.elements[widget::ChallengeIndex::ALPHA],
                             &key->quotient_large[0], BBG_KEEP_IF_AHEAD));
    return W::FFTGetter::update_alpha(challenges, W::FFTKernel::num_independent_relations);
}
} // namespace

namespace widget {
#define BBG_TURBO_WIDGET(SETTINGS, KERNEL, KIND)                                                                                        \
    template <>                                                                                                                         \
    fr TransitionWidget<fr, SETTINGS, KERNEL>::compute_quotient_contribution(const fr& alpha_base, const transcript::StandardTranscript& transcript) \
    {                                                                                                                                   \
        return turbo_quotient<SETTINGS, KERNEL>(this, KIND, alpha_base, transcript);                                                    \
    }
BBG_TURBO_WIDGET(turbo_settings, TurboArithmeticKernel, BBG_WIDGET_TURBO_ARITHMETIC)
BBG_TURBO_WIDGET(turbo_settings, TurboFixedBaseKernel, BBG_WIDGET_TURBO_FIXED_BASE)
BBG_TURBO_WIDGET(turbo_settings, TurboRangeKernel, BBG_WIDGET_TURBO_RANGE)
BBG_TURBO_WIDGET(turbo_settings, TurboLogicKernel, BBG_WIDGET_TURBO_LOGIC)
BBG_TURBO_WIDGET(unrolled_turbo_settings, TurboArithmeticKernel, BBG_WIDGET_TURBO_ARITHMETIC)
BBG_TURBO_WIDGET(unrolled_turbo_settings, TurboFixedBaseKernel, BBG_WIDGET_TURBO_FIXED_BASE)
BBG_TURBO_WIDGET(unrolled_turbo_settings, TurboRangeKernel, BBG_WIDGET_TURBO_RANGE)
BBG_TURBO_WIDGET(unrolled_turbo_settings, TurboLogicKernel, BBG_WIDGET_TURBO_LOGIC)
#undef BBG_TURBO_WIDGET
} // namespace widget

// ---- permutation argument, program width 4, identity permutation polynomials
template <>
fr ProverPermutationWidget<4, false, 4>::compute_quotient_contribution(const fr& alpha_base, const transcript::StandardTranscript& transcript)
{
    Trace trace("permutation_quotient");
    const fr beta = fr::serialize_from_buffer(transcript.get_challenge("beta").begin());
    const fr gamma = fr::serialize_from_buffer(transcript.get_challenge("beta", 1).begin());
    const void* wires[4];
    const void* sigmas[4];
    for (size_t i = 0; i < 4; ++i) {
        wires[i] = &key->wire_ffts.at("w_" + std::to_string(i + 1) + "_fft")[0];
        sigmas[i] = &key->permutation_selector_ffts.at("sigma_" + std::to_string(i + 1) + "_fft")[0];
    }
    std::vector<fr> public_inputs = many_from_buffer<fr>(transcript.get_element("public_inputs"));
    const fr delta = compute_public_input_delta<fr>(public_inputs, beta, gamma, key->small_domain.root);
    ensure_resident();
    // assignment into quotient_large, left on the device for the gate widgets that follow when this is a Turbo key
    check(bbg_permutation_quotient(wires, sigmas, 4, &key->wire_ffts.at("z_fft")[0], &key->lagrange_1[0], key->large_domain.size, 4, &alpha_base, &beta,
                                   &gamma, &delta, &key->quotient_large[0],
                                   (turbo_key(key) && bbg_resident_mode(-1) == 1) ? BBG_KEEP_ON_DEVICE : 0));
    return alpha_base.sqr().sqr();
}

template <>
void ProverPermutationWidget<4, false, 4>::compute_round_commitments(transcript::StandardTranscript& transcript, const size_t round_number,
                                                                       work_queue& queue)
{
    if (round_number != 3) {
        return;
    }
    Trace trace("permutation_grand_product");
    const size_t n = key->n;
    polynomial& z = witness->wires.at("z");
    const fr beta = fr::serialize_from_buffer(transcript.get_challenge("beta").begin());
    const fr gamma = fr::serialize_from_buffer(transcript.get_challenge("beta", 1).begin());
    const void* wires[4];
    const void* sigmas[4];
    for (size_t i = 0; i < 4; ++i) {
        wires[i] = &key->wire_ffts.at("w_" + std::to_string(i + 1) + "_fft")[0]; // Lagrange-base copies made in the preamble
        sigmas[i] = &key->permutation_selectors_lagrange_base.at("sigma_" + std::to_string(i + 1))[0];
    }
    ensure_resident();
    const bool resident = bbg_resident_mode(-1) == 1;
    check(bbg_permutation_grand_product(wires, sigmas, 4, n, &beta, &gamma, &z[0], resident ? BBG_KEEP_ON_DEVICE : 0));
    // blinding scalars: same positions and the same three draws as permutation_widget_impl.hpp:285-291
    fr blind[3];
    for (size_t k = 0; k < 3; ++k) blind[k] = fr::random_element();
    check(bbg_poly_write(&z[0], (n - 4) + 1, blind, 3));
    if (resident && turbo_key(key) && shim_enabled()) {
        // coefficient form in the mirror only: the commitment, the coset FFT item of THIS file's process_queue (the
        // reference's copies z on the host first), the opening evaluations and batch_open of a Turbo proof all read it
        // there (polynomial::ifft's size bookkeeping, polynomial.cpp:312-320, by hand)
        if (n > z.get_max_size()) z.reserve(n);
        check(bbg_ntt_ex(&z[0], n, BBG_IFFT, 0, nullptr, BBG_KEEP_ON_DEVICE));
        z.resize_unsafe(n);
    } else {
        z.ifft(key->small_domain);
    }
    queue.add_to_queue({ work_queue::WorkType::SCALAR_MULTIPLICATION, z.get_coefficients(), "Z", fr(0), 0 });
    queue.add_to_queue({ work_queue::WorkType::FFT, nullptr, "z", fr(0), 0 });
}

} // namespace waffle
