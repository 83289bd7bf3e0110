// poly_api.cu -- extern "C" entry points of the pointwise / scan kernels (poly.cu), host-pointer semantics.
//
// Every array argument is a caller-owned HOST array, like everywhere else in include/bbg.h.  With resident polynomials
// on (resident.cu) an input whose device mirror is valid is not uploaded again and an output can stay in its mirror
// (BBG_KEEP_ON_DEVICE) until bbg_resident_flush(): that is what lets a proof's polynomials live in HBM from the wire
// iFFTs to the quotient commitments.  With residency off every call uploads its inputs and downloads its outputs.
#include <algorithm>
#include <cstring>

#include "api_common.hpp"
#include "poly.hpp"

namespace bbg {

// One array argument of a call: where it lives on the host, whether the kernel needs its content and whether the kernel
// writes it; bind() fills in the device address (resident mirror or a slot of the staging buffer).
struct Arg {
    const void* host = nullptr;
    size_t bytes = 0;
    bool need_data = true;
    bool written = false;
    void* d = nullptr;
    bool resident = false;
};

static int bind(Context* ctx, Arg* args, int n, cudaStream_t st, uint64_t* h2d)
{
    int rc;
    size_t staged = 0;
    for (int i = 0; i < n; ++i) {
        Arg& a = args[i];
        if (a.host == nullptr || a.bytes == 0) continue;
        // the same array passed twice (in place): share the binding
        bool dup = false;
        for (int j = 0; j < i; ++j) {
            if (args[j].host == a.host && args[j].bytes >= a.bytes && args[j].d != nullptr && args[j].resident) {
                a.d = args[j].d;
                a.resident = true;
                dup = true;
                break;
            }
        }
        if (dup) continue;
        bool hit = false;
        if ((rc = resident_acquire(ctx, a.host, a.bytes, a.need_data, &a.d, &hit, st))) return rc;
        a.resident = a.d != nullptr;
        if (a.resident) {
            if (a.need_data && !hit) *h2d += a.bytes;
        } else {
            staged += (a.bytes + 255) & ~(size_t)255;
        }
    }
    if (staged) {
        if ((rc = ctx->poly_stage.reserve(staged))) return rc;
        size_t off = 0;
        for (int i = 0; i < n; ++i) {
            Arg& a = args[i];
            if (a.host == nullptr || a.bytes == 0 || a.resident) continue;
            // in-place arguments share a staging slot
            bool dup = false;
            for (int j = 0; j < i; ++j) {
                if (args[j].host == a.host && !args[j].resident && args[j].d != nullptr) {
                    a.d = args[j].d;
                    dup = true;
                    break;
                }
            }
            if (dup) continue;
            a.d = (char*)ctx->poly_stage.p + off;
            off += (a.bytes + 255) & ~(size_t)255;
            if (a.need_data) {
                if ((rc = g_staging.h2d(a.d, a.host, a.bytes, st))) return rc;
                *h2d += a.bytes;
            }
        }
    }
    return BBG_OK;
}

// after the kernels: bring written arrays home (or leave them in their mirror), then wait for the stream
static int finish(Context* ctx, Arg* args, int n, unsigned flags, cudaStream_t st, uint64_t* d2h)
{
    int rc;
    for (int i = 0; i < n; ++i) {
        Arg& a = args[i];
        if (!a.written || a.host == nullptr || a.bytes == 0) continue;
        if (a.resident) {
            const bool keep = (flags & BBG_KEEP_ON_DEVICE) != 0 || ((flags & BBG_KEEP_IF_AHEAD) != 0 && resident_is_ahead(ctx, a.host, a.bytes));
            if ((rc = resident_commit(ctx, a.host, a.bytes, !keep, st))) return rc;
            if (!keep) *d2h += a.bytes;
        } else {
            if ((rc = g_staging.d2h((void*)a.host, a.d, a.bytes, st))) return rc;
            *d2h += a.bytes;
        }
    }
    BBG_CUDA(cudaStreamSynchronize(st));
    return BBG_OK;
}

struct PolyScope {
    StatScope stat;
    DeviceTimer tm;
    PolyScope(Context* ctx) : stat(STAT_POLY, ctx, 0, 0), tm(ctx) {}
    void account(uint64_t h2d, uint64_t d2h)
    {
        if (stat.row) {
            stat.row->bytes_h2d += h2d;
            stat.row->bytes_d2h += d2h;
        }
        stat.h2d = h2d;
        stat.d2h = d2h;
    }
};

} // namespace bbg

using namespace bbg;

extern "C" {

int bbg_turbo_quotient(int kind, const void* const* polys, size_t n_large, const void* alpha_base, const void* alpha, void* quotient,
                       unsigned flags)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!polys || !alpha_base || !alpha || !quotient) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    Arg args[BBG_POLY_COUNT + 1];
    for (int k = 0; k < BBG_POLY_COUNT; ++k) {
        args[k].host = polys[k];
        args[k].bytes = n_large * 32;
    }
    // only the polynomials this widget reads travel
    static const uint64_t reads[4] = {
        (1ull << BBG_POLY_W_1) | (1ull << BBG_POLY_W_2) | (1ull << BBG_POLY_W_3) | (1ull << BBG_POLY_W_4) | (1ull << BBG_POLY_Q_1) |
            (1ull << BBG_POLY_Q_2) | (1ull << BBG_POLY_Q_3) | (1ull << BBG_POLY_Q_4) | (1ull << BBG_POLY_Q_5) | (1ull << BBG_POLY_Q_M) |
            (1ull << BBG_POLY_Q_C) | (1ull << BBG_POLY_Q_ARITHMETIC_SELECTOR),
        (1ull << BBG_POLY_W_1) | (1ull << BBG_POLY_W_2) | (1ull << BBG_POLY_W_3) | (1ull << BBG_POLY_W_4) | (1ull << BBG_POLY_Q_1) |
            (1ull << BBG_POLY_Q_2) | (1ull << BBG_POLY_Q_3) | (1ull << BBG_POLY_Q_4) | (1ull << BBG_POLY_Q_5) | (1ull << BBG_POLY_Q_M) |
            (1ull << BBG_POLY_Q_C) | (1ull << BBG_POLY_Q_FIXED_BASE_SELECTOR),
        (1ull << BBG_POLY_W_1) | (1ull << BBG_POLY_W_2) | (1ull << BBG_POLY_W_3) | (1ull << BBG_POLY_W_4) | (1ull << BBG_POLY_Q_RANGE_SELECTOR),
        (1ull << BBG_POLY_W_1) | (1ull << BBG_POLY_W_2) | (1ull << BBG_POLY_W_3) | (1ull << BBG_POLY_W_4) | (1ull << BBG_POLY_Q_C) |
            (1ull << BBG_POLY_Q_LOGIC_SELECTOR),
    };
    if (kind < 0 || kind > 3) {
        set_last_error("turbo_quotient: unknown widget kind");
        return BBG_ERR_ARG;
    }
    for (int k = 0; k < BBG_POLY_COUNT; ++k) {
        if (!((reads[kind] >> k) & 1)) args[k].host = nullptr;
    }
    Arg& q = args[BBG_POLY_COUNT];
    q.host = quotient;
    q.bytes = n_large * 32;
    q.written = true;
    uint64_t h2d = 0, d2h = 0;
    int rc;
    if ((rc = bind(ctx, args, BBG_POLY_COUNT + 1, ctx->stream, &h2d))) return rc;
    PolyScope scope(ctx);
    const void* d_polys[BBG_POLY_COUNT];
    for (int k = 0; k < BBG_POLY_COUNT; ++k) d_polys[k] = args[k].d;
    if ((rc = poly_turbo_quotient_device(ctx, kind, d_polys, n_large, alpha_base, alpha, q.d, ctx->stream))) return rc;
    scope.tm.stop();
    if ((rc = finish(ctx, args, BBG_POLY_COUNT + 1, flags, ctx->stream, &d2h))) return rc;
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

int bbg_permutation_quotient(const void* const* wire_ffts, const void* const* sigma_ffts, unsigned program_width, const void* z_fft,
                             const void* lagrange_1, size_t n_large, unsigned num_roots_cut, const void* alpha_base, const void* beta,
                             const void* gamma, const void* public_input_delta, void* quotient, unsigned flags)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!wire_ffts || !sigma_ffts || !z_fft || !lagrange_1 || !alpha_base || !beta || !gamma || !public_input_delta || !quotient ||
        program_width < 1 || program_width > 4) {
        set_last_error("permutation_quotient: null argument or program width outside 1..4");
        return BBG_ERR_ARG;
    }
    Arg args[11];
    for (unsigned k = 0; k < program_width; ++k) {
        args[k].host = wire_ffts[k];
        args[k].bytes = n_large * 32;
        args[4 + k].host = sigma_ffts[k];
        args[4 + k].bytes = n_large * 32;
    }
    args[8].host = z_fft;
    args[8].bytes = n_large * 32;
    args[9].host = lagrange_1;
    args[9].bytes = n_large * 32;
    args[10].host = quotient;
    args[10].bytes = n_large * 32;
    args[10].need_data = false; // assignment: the first widget to run (permutation_widget_impl.hpp:430)
    args[10].written = true;
    uint64_t h2d = 0, d2h = 0;
    int rc;
    if ((rc = bind(ctx, args, 11, ctx->stream, &h2d))) return rc;
    PolyScope scope(ctx);
    PermArgs A;
    memset(&A, 0, sizeof(A));
    for (unsigned k = 0; k < program_width; ++k) {
        A.d_wires[k] = args[k].d;
        A.d_sigmas[k] = args[4 + k].d;
    }
    A.d_z = args[8].d;
    A.d_l_start = args[9].d;
    A.d_quotient = args[10].d;
    A.n_large = n_large;
    A.width = program_width;
    A.roots_cut = num_roots_cut;
    A.alpha_base = hf::load(alpha_base);
    A.beta = hf::load(beta);
    A.gamma = hf::load(gamma);
    A.public_input_delta = hf::load(public_input_delta);
    if ((rc = poly_permutation_quotient_device(ctx, A, ctx->stream))) return rc;
    scope.tm.stop();
    if ((rc = finish(ctx, args, 11, flags, ctx->stream, &d2h))) return rc;
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

int bbg_divide_by_pseudo_vanishing_polynomial(void* evaluations, size_t n_small, size_t n_large, unsigned num_roots_cut, unsigned flags)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!evaluations) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    Arg a;
    a.host = evaluations;
    a.bytes = n_large * 32;
    a.written = true;
    uint64_t h2d = 0, d2h = 0;
    int rc;
    if ((rc = bind(ctx, &a, 1, ctx->stream, &h2d))) return rc;
    PolyScope scope(ctx);
    if ((rc = poly_divide_vanishing_device(ctx, a.d, n_small, n_large, num_roots_cut, ctx->stream))) return rc;
    scope.tm.stop();
    if ((rc = finish(ctx, &a, 1, flags, ctx->stream, &d2h))) return rc;
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

int bbg_compute_lagrange_polynomial_fft(void* l_1_coefficients, size_t n_small, size_t n_large)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!l_1_coefficients) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    Arg a;
    a.host = l_1_coefficients;
    a.bytes = n_large * 32;
    a.need_data = false;
    a.written = true;
    uint64_t h2d = 0, d2h = 0;
    int rc;
    if ((rc = bind(ctx, &a, 1, ctx->stream, &h2d))) return rc;
    PolyScope scope(ctx);
    if ((rc = poly_lagrange_l1_device(ctx, a.d, n_small, n_large, ctx->stream))) return rc;
    scope.tm.stop();
    if ((rc = finish(ctx, &a, 1, 0, ctx->stream, &d2h))) return rc;
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

int bbg_permutation_grand_product(const void* const* wires_lagrange, const void* const* sigmas_lagrange, unsigned program_width, size_t n,
                                  const void* beta, const void* gamma, void* z, unsigned flags)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!wires_lagrange || !sigmas_lagrange || !beta || !gamma || !z || program_width < 1 || program_width > 4) {
        set_last_error("grand_product: null argument or program width outside 1..4");
        return BBG_ERR_ARG;
    }
    Arg args[9];
    for (unsigned k = 0; k < program_width; ++k) {
        args[k].host = wires_lagrange[k];
        args[k].bytes = n * 32;
        args[4 + k].host = sigmas_lagrange[k];
        args[4 + k].bytes = n * 32;
    }
    args[8].host = z;
    args[8].bytes = n * 32;
    args[8].need_data = false;
    args[8].written = true;
    uint64_t h2d = 0, d2h = 0;
    int rc;
    if ((rc = bind(ctx, args, 9, ctx->stream, &h2d))) return rc;
    PolyScope scope(ctx);
    GrandArgs A;
    memset(&A, 0, sizeof(A));
    for (unsigned k = 0; k < program_width; ++k) {
        A.d_wires[k] = args[k].d;
        A.d_sigmas[k] = args[4 + k].d;
    }
    A.d_z = args[8].d;
    A.n = n;
    A.width = program_width;
    A.beta = hf::load(beta);
    A.gamma = hf::load(gamma);
    if ((rc = poly_grand_product_device(ctx, A, ctx->stream))) return rc;
    scope.tm.stop();
    if ((rc = finish(ctx, args, 9, flags, ctx->stream, &d2h))) return rc;
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

int bbg_evaluate(const void* coeffs, size_t n, const void* z, void* result)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!z || !result || (n && !coeffs)) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    if (n == 0) {
        memset(result, 0, 32);
        return BBG_OK;
    }
    Arg a;
    a.host = coeffs;
    a.bytes = n * 32;
    uint64_t h2d = 0, d2h = 32;
    int rc;
    if ((rc = bind(ctx, &a, 1, ctx->stream, &h2d))) return rc;
    if ((rc = ctx->msm_ws[0].result.reserve(96))) return rc;
    PolyScope scope(ctx);
    if ((rc = poly_evaluate_device(ctx, a.d, n, hf::load(z), ctx->msm_ws[0].result.p, ctx->stream))) return rc;
    scope.tm.stop();
    BBG_CUDA(cudaMemcpyAsync(result, ctx->msm_ws[0].result.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

// count evaluations in one launch: results[k] = sum_i polys[k][i] z_k^i over ns[k] coefficients (count <= 40)
int bbg_evaluate_batch(const void* const* polys, const size_t* ns, size_t count, const void* zs, void* results)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (count == 0) return BBG_OK;
    if (!polys || !ns || !zs || !results || count > EVAL_BATCH_MAX) {
        set_last_error("evaluate_batch: null argument or more than 40 polynomials");
        return BBG_ERR_ARG;
    }
    Arg args[EVAL_BATCH_MAX];
    for (size_t k = 0; k < count; ++k) {
        args[k].host = polys[k];
        args[k].bytes = ns[k] * 32;
    }
    uint64_t h2d = 0;
    int rc;
    if ((rc = bind(ctx, args, (int)count, ctx->stream, &h2d))) return rc;
    if ((rc = ctx->poly_out.reserve(EVAL_BATCH_MAX * 32))) return rc;
    PolyScope scope(ctx);
    const void* d_polys[EVAL_BATCH_MAX];
    for (size_t k = 0; k < count; ++k) d_polys[k] = args[k].d;
    if ((rc = poly_evaluate_batch_device(ctx, d_polys, ns, zs, count, ctx->poly_out.p, ctx->stream))) return rc;
    scope.tm.stop();
    BBG_CUDA(cudaMemcpyAsync(results, ctx->poly_out.p, count * 32, cudaMemcpyDeviceToHost, ctx->stream));
    scope.account(h2d, count * 32);
    return scope.tm.finish();
}

int bbg_compute_opening_polynomial(const void* src, void* dest, const void* z, size_t n_eval, size_t n, void* f_at_z, unsigned flags)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!src || !dest || !z || n_eval == 0 || n > n_eval) {
        set_last_error("compute_opening_polynomial: null argument or n > n_eval");
        return BBG_ERR_ARG;
    }
    Arg args[2];
    args[0].host = src;
    args[0].bytes = n_eval * 32;
    args[1].host = dest;
    args[1].bytes = n * 32;
    args[1].need_data = (dest == src);
    args[1].written = true;
    if (dest == src) {
        // in place: one binding, written
        args[0].written = false;
        args[1].bytes = n_eval * 32;
    }
    uint64_t h2d = 0, d2h = 0;
    int rc;
    if ((rc = bind(ctx, args, 2, ctx->stream, &h2d))) return rc;
    if ((rc = ctx->msm_ws[0].result.reserve(96))) return rc;
    PolyScope scope(ctx);
    if ((rc = poly_opening_device(ctx, args[0].d, n_eval, n, hf::load(z), args[1].d, ctx->msm_ws[0].result.p, ctx->stream))) return rc;
    scope.tm.stop();
    if (f_at_z) BBG_CUDA(cudaMemcpyAsync(f_at_z, ctx->msm_ws[0].result.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = finish(ctx, args, 2, flags, ctx->stream, &d2h))) return rc;
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

int bbg_linear_combination(void* dest, const void* base, const void* const* polys, const void* scalars, size_t count, size_t n, unsigned flags)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!dest || (count && (!polys || !scalars)) || count > LINCOMB_MAX) {
        set_last_error("linear_combination: null argument or more than 48 terms");
        return BBG_ERR_ARG;
    }
    Arg args[LINCOMB_MAX + 2];
    for (size_t k = 0; k < count; ++k) {
        args[k].host = polys[k];
        args[k].bytes = n * 32;
    }
    Arg& b = args[LINCOMB_MAX];
    b.host = base;
    b.bytes = n * 32;
    Arg& d = args[LINCOMB_MAX + 1];
    d.host = dest;
    d.bytes = n * 32;
    d.need_data = false;
    d.written = true;
    // dest aliasing an input: keep its content
    for (size_t k = 0; k < count; ++k) {
        if (polys[k] == dest) d.need_data = true;
    }
    if (base == dest) d.need_data = true;
    uint64_t h2d = 0, d2h = 0;
    int rc;
    if ((rc = bind(ctx, args, LINCOMB_MAX + 2, ctx->stream, &h2d))) return rc;
    PolyScope scope(ctx);
    const void* d_polys[LINCOMB_MAX];
    for (size_t k = 0; k < count; ++k) d_polys[k] = args[k].d;
    if ((rc = poly_linear_combination_device(ctx, d.d, base ? b.d : nullptr, d_polys, scalars, count, n, ctx->stream))) return rc;
    scope.tm.stop();
    if ((rc = finish(ctx, args, LINCOMB_MAX + 2, flags, ctx->stream, &d2h))) return rc;
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

// work_queue FFT item (bb/plonk/proof_system/prover/work_queue.hpp:260-270): wire_fft[0, ext n + ext) = the ext n-point
// coset FFT of the n wire coefficients, followed by its first `ext` values again
int bbg_wire_coset_fft(const void* wire, void* wire_fft, size_t n, size_t ext, unsigned flags)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    unsigned lg = 0, lge = 0;
    while (((size_t)1 << lg) < n) ++lg;
    while (((size_t)1 << lge) < ext) ++lge;
    if (!wire || !wire_fft || n == 0 || ((size_t)1 << lg) != n || ((size_t)1 << lge) != ext || lg + lge > 28) {
        set_last_error("wire_coset_fft: n and the extension must be powers of two with n * ext <= 2^28");
        return BBG_ERR_ARG;
    }
    const size_t big = n * ext;
    Arg args[2];
    args[0].host = wire;
    args[0].bytes = n * 32;
    args[1].host = wire_fft;
    args[1].bytes = (big + ext) * 32;
    args[1].need_data = false;
    args[1].written = true;
    uint64_t h2d = 0, d2h = 0;
    int rc;
    if ((rc = bind(ctx, args, 2, ctx->stream, &h2d))) return rc;
    PolyScope scope(ctx);
    if ((rc = poly_copy_pad_device(ctx, args[0].d, args[1].d, n, big, ctx->stream))) return rc;
    NttScale pro, epi;
    pro.present = true;
    pro.start = hf::one();
    pro.has_shift = true;
    pro.shift = hf::from_u64(5);
    pro.size = n; // large domains are built with generator_size = n (proving_key.cpp:20-22): the rest is zero padding
    if ((rc = ntt_device(ctx, args[1].d, args[1].d, lg + lge, false, pro, epi, 0, 0, ctx->stream))) return rc;
    BBG_CUDA(cudaMemcpyAsync((char*)args[1].d + big * 32, args[1].d, ext * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    scope.tm.stop();
    if ((rc = finish(ctx, args, 2, flags, ctx->stream, &d2h))) return rc;
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

// work_queue IFFT item (work_queue.hpp:272-276) for a wire whose Lagrange-base copy the prover keeps in `lagrange_copy`:
// wire <- ifft(wire) in place, and the device mirror of lagrange_copy[0, n) is seeded from the data uploaded for the
// transform (prover.cpp:184-186 memcpy'd it from `wire` just before), so round 3's grand product finds it on the device
int bbg_wire_ifft(void* wire, size_t n, const void* lagrange_copy, unsigned flags)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!wire || n == 0 || (n & (n - 1))) {
        set_last_error("wire_ifft: null argument or n not a power of two");
        return BBG_ERR_ARG;
    }
    Arg args[2];
    args[0].host = wire;
    args[0].bytes = n * 32;
    args[0].written = true;
    args[1].host = lagrange_copy;
    args[1].bytes = n * 32;
    args[1].need_data = false;
    uint64_t h2d = 0, d2h = 0;
    int rc;
    if ((rc = bind(ctx, args, 2, ctx->stream, &h2d))) return rc;
    PolyScope scope(ctx);
    if (lagrange_copy != nullptr && args[1].resident) {
        BBG_CUDA(cudaMemcpyAsync(args[1].d, args[0].d, n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        resident_adopt(ctx, lagrange_copy, n * 32);
    }
    if ((rc = ntt_run_kind(ctx, args[0].d, n, BBG_IFFT, 0, nullptr, ctx->stream))) return rc;
    scope.tm.stop();
    if ((rc = finish(ctx, args, 2, flags, ctx->stream, &d2h))) return rc;
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

// The IFFT items of one queue flush together (the four wires of a proof, work_queue.hpp:272-276).  The witness columns live
// in pageable memory; a cudaMemcpyAsync from there is staged by the driver on the calling thread (~10 GB/s, and the GPU idles
// meanwhile).  Here every column is copied into a pinned slot by the staging pool's threads and goes up with a true
// asynchronous copy, so column k + 1 is being staged while column k is in flight and being transformed; one stream
// synchronisation at the end instead of one per column.
int bbg_wire_ifft_batch(void* const* wires, size_t n, const void* const* lagrange_copies, size_t count, unsigned flags)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!wires || n == 0 || (n & (n - 1)) || count == 0 || count > 16) {
        set_last_error("wire_ifft_batch: null argument, n not a power of two, or more than 16 columns");
        return BBG_ERR_ARG;
    }
    const size_t bytes = n * 32;
    if (count * bytes > ctx->pinned_cap) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_cap = 0;
        BBG_CUDA(cudaHostAlloc(&ctx->pinned, count * bytes, cudaHostAllocDefault));
        ctx->pinned_cap = count * bytes;
    }
    int rc;
    if ((rc = g_staging.ensure())) return rc;
    uint64_t h2d = 0, d2h = 0;
    PolyScope scope(ctx);
    Arg args[32];
    for (size_t k = 0; k < count; ++k) {
        if (!wires[k]) {
            set_last_error("wire_ifft_batch: null column");
            return BBG_ERR_ARG;
        }
        Arg* a = args + 2 * k;
        a[0].host = wires[k];
        a[0].bytes = bytes;
        a[0].written = true;
        a[0].need_data = false; // uploaded below, from the pinned slot
        a[1].host = lagrange_copies ? lagrange_copies[k] : nullptr;
        a[1].bytes = bytes;
        a[1].need_data = false;
        if ((rc = bind(ctx, a, 2, ctx->stream, &h2d))) return rc;
        char* slot = (char*)ctx->pinned + k * bytes;
        if (g_staging.pool != nullptr) {
            g_staging.pool->copy(slot, wires[k], bytes);
        } else {
            memcpy(slot, wires[k], bytes);
        }
        BBG_CUDA(cudaMemcpyAsync(a[0].d, slot, bytes, cudaMemcpyHostToDevice, ctx->stream));
        h2d += bytes;
        if (a[1].host != nullptr && a[1].resident) {
            BBG_CUDA(cudaMemcpyAsync(a[1].d, a[0].d, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            resident_adopt(ctx, a[1].host, bytes);
        }
        if ((rc = ntt_run_kind(ctx, a[0].d, n, BBG_IFFT, 0, nullptr, ctx->stream))) return rc;
        if (!a[0].resident) {
            // no mirrors (resident polynomials off): the column sits in the shared staging slot, bring it home before the next
            // column reuses the slot
            if ((rc = finish(ctx, a, 2, flags, ctx->stream, &d2h))) return rc;
            a[0].written = false;
        }
    }
    scope.tm.stop();
    if ((rc = finish(ctx, args, (int)(2 * count), flags, ctx->stream, &d2h))) return rc;
    scope.account(h2d, d2h);
    return scope.tm.finish();
}

// host[elem_offset, elem_offset + count) = values, in host memory AND in the array's device mirror if it has one
// (the prover's blinding scalars, prover.cpp:181-183 / permutation_widget_impl.hpp:289-291, written between two device steps)
int bbg_poly_write(void* host_array, size_t elem_offset, const void* values, size_t count)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!host_array || (count && !values)) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    char* dst = (char*)host_array + elem_offset * 32;
    if (resident_enabled(ctx)) {
        for (Context::Resident& e : ctx->resident) {
            if (dst >= e.host && dst + count * 32 <= e.host + e.bytes) {
                BBG_CUDA(cudaMemcpyAsync((char*)e.d + (dst - e.host), values, count * 32, cudaMemcpyHostToDevice, ctx->stream));
                BBG_CUDA(cudaStreamSynchronize(ctx->stream));
                memcpy(dst, values, count * 32);
                // refresh the fingerprint words that fall inside the range
                for (uint32_t k = 0; k < e.n_samples; ++k) {
                    const uint64_t off = e.sample_off[k];
                    if (off >= (uint64_t)(dst - e.host) && off + 8 <= (uint64_t)(dst - e.host) + count * 32) {
                        memcpy(&e.sample_val[k], e.host + off, 8);
                    }
                }
                return BBG_OK;
            }
        }
    }
    memcpy(dst, values, count * 32);
    return BBG_OK;
}

} // extern "C"
