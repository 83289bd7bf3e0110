// api.cu -- the extern "C" boundary of libbbg.so (declared in include/bbg.h).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>

#include "api_common.hpp"

namespace bbg {

// ---- error text + context
static thread_local std::string g_last_error;
static std::string g_last_error_global;
void set_last_error(const std::string& s)
{
    g_last_error = s;
    g_last_error_global = s;
}

static Context* g_ctx = nullptr;
static std::mutex g_ctx_mu;

static int create_context(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_last_error("no CUDA device available (libbbg has no CPU fallback)");
        return BBG_ERR_NO_DEVICE;
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) {
            device = 0;
        }
    }
    if (device >= count) {
        set_last_error("bbg_init: device index out of range");
        return BBG_ERR_ARG;
    }
    BBG_CUDA(cudaSetDevice(device));
    Context* c = new Context();
    c->device = device;
    cudaDeviceProp prop;
    BBG_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    if (getenv("BBG_STACK")) BBG_CUDA(cudaDeviceSetLimit(cudaLimitStackSize, (size_t)atoi(getenv("BBG_STACK"))));
    BBG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    BBG_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto& a : c->aux_stream) BBG_CUDA(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    for (auto& e : c->ev_piece) BBG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    BBG_CUDA(cudaEventCreate(&c->ev_a));
    BBG_CUDA(cudaEventCreate(&c->ev_b));
    BBG_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (auto& e : c->ev_join) BBG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    {
        int prio_low = 0, prio_high = 0;
        BBG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
        for (auto& a : c->part_stream) BBG_CUDA(cudaStreamCreateWithPriority(&a, cudaStreamNonBlocking, prio_high));
        BBG_CUDA(cudaEventCreateWithFlags(&c->ev_part_fork, cudaEventDisableTiming));
        for (auto& e : c->ev_part_join) BBG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    BBG_CUDA(cudaEventCreateWithFlags(&c->last_use, cudaEventDisableTiming));
    g_ctx = c;
    return BBG_OK;
}

// Callers may come from any thread and may have another device current (a PyTorch process): every entry point binds the
// context's device for its own duration and puts the caller's device back on exit (DeviceGuard inside GET_CTX).
int get_context(Context** out)
{
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (g_ctx == nullptr) {
        int rc = create_context(-1);
        if (rc) return rc;
    }
    *out = g_ctx;
    return BBG_OK;
}
DeviceGuard::DeviceGuard(int device)
{
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) {
        cudaSetDevice(device);
    } else {
        prev = -1; // nothing to restore
    }
}
DeviceGuard::~DeviceGuard()
{
    if (prev >= 0) cudaSetDevice(prev);
}

Staging g_staging;
HostStats g_stats;
std::vector<PippengerObj*> g_pippengers;

static int g_auto_adopt = -1; // -1: read BBG_AUTO_ADOPT on first use

static int upload_even_entries(Context* ctx, const void* table2n, size_t n, affine_t* d_points)
{
    // strided copy: every other 64-byte entry
    BBG_CUDA(cudaMemcpy2DAsync(d_points, 64, table2n, 128, 64, n, cudaMemcpyHostToDevice, ctx->stream));
    return BBG_OK;
}

static const uint64_t G1_ONE_X[4] = { 0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL };
static const uint64_t G1_ONE_Y[4] = { 0xa6ba871b8b1e1b3aULL, 0x14f1d651eb8e167bULL, 0xccdd46def0f28c58ULL, 0x1c14ef83340fbe5eULL };

// monomials[0] = generator (1, 2); the file points follow (bb/srs/io.cpp:134-139, pippenger.cpp:7-16)
static int decode_raw_points(Context* ctx, const uint8_t* raw, size_t num_file_points, affine_t* d_points_after_generator)
{
    if (num_file_points == 0) {
        return BBG_OK;
    }
    int rc = ctx->msm_points.reserve(num_file_points * 64);
    if (rc) return rc;
    BBG_CUDA(cudaMemcpyAsync(ctx->msm_points.p, raw, num_file_points * 64, cudaMemcpyHostToDevice, ctx->stream));
    return srs_decode_device(ctx, ctx->msm_points.p, num_file_points, d_points_after_generator, ctx->stream);
}

static int set_generator(Context* ctx, affine_t* d_point0)
{
    uint64_t g[8];
    memcpy(g, G1_ONE_X, 32);
    memcpy(g + 4, G1_ONE_Y, 32);
    BBG_CUDA(cudaMemcpyAsync(d_point0, g, 64, cudaMemcpyHostToDevice, ctx->stream));
    BBG_CUDA(cudaStreamSynchronize(ctx->stream));
    return BBG_OK;
}

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// Reads up to `degree - 1` file points from <dir>/transcript00.dat, transcript01.dat, ... into a host buffer
// (raw, undecoded).  Manifest = 7 big-endian u32 (bb/srs/io.cpp:11-45); field 4 = num_g1_points.
static int read_transcript_raw(const char* dir, size_t degree, std::vector<uint8_t>& raw)
{
    raw.clear();
    size_t need = degree > 0 ? degree - 1 : 0;
    raw.reserve(need * 64);
    for (int num = 0; raw.size() / 64 < need; ++num) {
        char path[4096];
        snprintf(path, sizeof(path), "%s/transcript%02d.dat", dir, num);
        std::ifstream f(path, std::ifstream::binary);
        if (!f.good()) {
            break;
        }
        uint8_t man[28];
        f.read((char*)man, 28);
        if (f.gcount() != 28) {
            break;
        }
        size_t num_g1 = be32(man + 16);
        size_t to_read = std::min(num_g1, need - raw.size() / 64);
        size_t old = raw.size();
        raw.resize(old + to_read * 64);
        f.read((char*)raw.data() + old, (std::streamsize)(to_read * 64));
        if ((size_t)f.gcount() != to_read * 64) {
            raw.resize(old + (size_t)f.gcount() / 64 * 64);
            break;
        }
    }
    if (raw.size() / 64 < need) {
        // same condition and wording as bb/srs/io.cpp:159-161
        set_last_error("Only read " + std::to_string(raw.size() / 64 + 1) + " points but require " + std::to_string(degree) +
                       ". Is your srs large enough?");
        return BBG_ERR_SRS;
    }
    return BBG_OK;
}

static PippengerObj* new_obj(Context* ctx, size_t n)
{
    PippengerObj* o = new PippengerObj();
    o->n = n;
    // fixed-base levels: as many as fit in half of the free HBM (all of them for every size the prover uses)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)8 << 30;
    o->lv = msm_levels_plan(std::max<size_t>(n, 1), free_b / 2);
    if (cudaMalloc(&o->d_points, std::max<size_t>(n, 1) * 64 * o->lv.L) != cudaSuccess) {
        set_last_error("cudaMalloc failed for SRS points");
        delete o;
        return nullptr;
    }
    (void)ctx;
    g_pippengers.push_back(o);
    return o;
}

// probe kernels ------------------------------------------------------------------------------------
template <class F> __global__ void k_field_op(int op, const Fe<F>* a, const Fe<F>* b, Fe<F>* out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fe<F> x = fe_load<F>(a + i);
    Fe<F> y = b ? fe_load<F>(b + i) : fe_zero<F>();
    Fe<F> r;
    switch (op) {
    case 0: r = fe_mul(x, y); break;
    case 1: r = fe_add(x, y); break;
    case 2: r = fe_sub(x, y); break;
    case 3: r = fe_sqr(x); break;
    case 4: r = fe_to_mont(fe_reduce_once(x)); break;
    case 5: r = fe_from_mont(x); break;
    case 7: r = fe_reduce_once(x); break;
    case 8: r = fe_neg(x); break;
    default: r = fe_zero<F>();
    }
    fe_store(out + i, r);
}
__global__ void k_g1_op(int op, const jac_t* a, const void* b, jac_t* out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    jac_t ja;
    ja.x = fe_load<FqParams>(&a[i].x);
    ja.y = fe_load<FqParams>(&a[i].y);
    ja.z = fe_load<FqParams>(&a[i].z);
    xyzz_t p = xyzz_from_jacobian(ja);
    if (op == 0) {
        affine_t q = affine_load(reinterpret_cast<const affine_t*>(b) + i);
        if (!affine_is_inf(q)) {
            xyzz_madd(p, q);
        }
    } else if (op == 1) {
        const jac_t* bj = reinterpret_cast<const jac_t*>(b) + i;
        jac_t jb;
        jb.x = fe_load<FqParams>(&bj->x);
        jb.y = fe_load<FqParams>(&bj->y);
        jb.z = fe_load<FqParams>(&bj->z);
        xyzz_t q = xyzz_from_jacobian(jb);
        xyzz_add(p, q);
    } else {
        p = xyzz_dbl(p);
    }
    jac_t r = xyzz_to_jacobian(p);
    fe_store(&out[i].x, r.x);
    fe_store(&out[i].y, r.y);
    fe_store(&out[i].z, r.z);
}

// g1::affine_element(element) (bb/ecc/groups/element_impl.hpp:51-68): Jacobian -> canonical affine, one inversion each
__global__ void __launch_bounds__(64) k_g1_normalize(const jac_t* __restrict__ in, affine_t* __restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    jac_t j;
    j.x = fe_load<FqParams>(&in[i].x);
    j.y = fe_load<FqParams>(&in[i].y);
    j.z = fe_load<FqParams>(&in[i].z);
    affine_t r = xyzz_to_affine(xyzz_from_jacobian(j));
    fe_store(&out[i].x, r.x);
    fe_store(&out[i].y, r.y);
}

template <class F> __global__ void __launch_bounds__(256) k_bench_mul(Fe<F>* out, int iters)
{
    Fe<F> x, y;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x.l[i] = threadIdx.x * 7 + i + 1;
        y.l[i] = blockIdx.x + i * 3 + 5;
    }
    x.l[7] &= 0x0fffffff;
    y.l[7] &= 0x0fffffff;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        x = fe_mul(x, y);
        y = fe_mul(y, x);
    }
    if (x.l[0] == 0x12345678 && y.l[3] == 0x9abcdef0) fe_store(out + threadIdx.x, x); // never true in practice; keeps the chain live
}
__global__ void __launch_bounds__(128) k_g1_add_affine(const affine_t* __restrict__ in, affine_t* __restrict__ out, size_t n, affine_t q)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t p = affine_load(in + i);
    xyzz_t acc = xyzz_from_affine(p);
    if (!affine_is_inf(q)) xyzz_madd(acc, q);
    affine_t r = xyzz_to_affine(acc);
    fe_store(&out[i].x, r.x);
    fe_store(&out[i].y, r.y);
}

// NTT kind -> prologue/epilogue scalings (bb/polynomials/polynomial_arithmetic.cpp:374-484)
static hf::Fr coset_generator()
{
    return hf::from_u64(5); // fr::coset_generator(0), bb/ecc/curves/bn254/fr.hpp:44-59
}
static int ntt_kind_params(int kind, unsigned log_n, size_t generator_size, const void* constant, bool& inverse, NttScale& pro, NttScale& epi)
{
    const uint64_t n = 1ull << log_n;
    if (generator_size == 0 || generator_size > n) {
        generator_size = n; // evaluation_domain.cpp:64
    }
    hf::Fr k = hf::one();
    if (kind >= 4 && kind <= 7) {
        if (constant == nullptr) {
            set_last_error("ntt: this kind needs a constant");
            return BBG_ERR_ARG;
        }
        k = hf::reduce(hf::load(constant));
    }
    const hf::Fr g = coset_generator();
    const hf::Fr n_inv = hf::invert(hf::from_u64(n));
    inverse = false;
    pro = NttScale();
    epi = NttScale();
    switch (kind) {
    case BBG_FFT: break;
    case BBG_IFFT:
        inverse = true;
        epi.present = true;
        epi.start = n_inv;
        break;
    case BBG_COSET_FFT:
        pro.present = true;
        pro.start = hf::one();
        pro.has_shift = true;
        pro.shift = g;
        pro.size = generator_size;
        break;
    case BBG_COSET_IFFT:
        inverse = true;
        epi.present = true;
        epi.start = n_inv;
        epi.has_shift = true;
        epi.shift = hf::invert(g);
        break;
    case BBG_FFT_WITH_CONSTANT:
        epi.present = true;
        epi.start = k;
        break;
    case BBG_IFFT_WITH_CONSTANT:
        inverse = true;
        epi.present = true;
        epi.start = hf::mul(n_inv, k);
        break;
    case BBG_COSET_FFT_WITH_CONSTANT:
        pro.present = true;
        pro.start = k;
        pro.has_shift = true;
        pro.shift = g;
        pro.size = generator_size;
        break;
    case BBG_COSET_FFT_WITH_GENERATOR_SHIFT:
        pro.present = true;
        pro.start = hf::one();
        pro.has_shift = true;
        pro.shift = hf::mul(g, k);
        pro.size = generator_size;
        break;
    default:
        set_last_error("ntt: unknown kind");
        return BBG_ERR_ARG;
    }
    return BBG_OK;
}

static int log2_exact(size_t n, unsigned& lg)
{
    lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    if (((size_t)1 << lg) != n || n == 0) {
        set_last_error("ntt: size must be a power of two");
        return BBG_ERR_ARG;
    }
    return BBG_OK;
}

struct DomainObj {
    size_t n;
};

// one transform of the bb/polynomials/polynomial_arithmetic.hpp:23-39 family on a device array, in place (no locking)
int ntt_run_kind(Context* ctx, void* d_coeffs, size_t n, int kind, size_t generator_size, const void* constant, cudaStream_t st)
{
    unsigned lg;
    int rc = log2_exact(n, lg);
    if (rc) return rc;
    bool inverse;
    NttScale pro, epi;
    if ((rc = ntt_kind_params(kind, lg, generator_size, constant, inverse, pro, epi))) return rc;
    return ntt_device(ctx, d_coeffs, d_coeffs, lg, inverse, pro, epi, 0, 0, st);
}

} // namespace bbg

using namespace bbg;


extern "C" {

int bbg_init(int device)
{
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (g_ctx != nullptr) {
        if (device >= 0 && device != g_ctx->device) {
            set_last_error("bbg_init: the library is already bound to device " + std::to_string(g_ctx->device) +
                           " (an earlier call created the context); call bbg_shutdown() first");
            return BBG_ERR_ARG;
        }
        return BBG_OK;
    }
    return create_context(device);
}

void bbg_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (g_ctx == nullptr) return;
    cudaSetDevice(g_ctx->device);
    cudaDeviceSynchronize();
    // Handles stay valid as *names* only: the objects are freed here, and bbg_delete_pippenger / any later use of a stale
    // handle finds it absent from g_pippengers and does nothing (Python's Pippenger.__del__ at interpreter exit, the
    // shim's ~Pippenger during static destruction).
    for (auto* o : g_pippengers) {
        cudaFree(o->d_points);
        delete o;
    }
    g_pippengers.clear();
    resident_clear(g_ctx);
    for (auto& kv : g_ctx->ntt_twiddles) cudaFree(kv.second);
    for (auto& e : g_ctx->ntt_scale_cache) {
        if (e.tab) cudaFree(e.tab);
    }
    for (int d = 0; d < 2; ++d) {
        if (g_ctx->ntt_stage_tw[d]) cudaFree(g_ctx->ntt_stage_tw[d]);
    }
    g_ctx->ntt_twiddles.clear();
    g_ctx->ntt_scale_cache.clear();
    for (auto& ws : g_ctx->msm_ws) ws.release();
    DevBuf* bufs[] = { &g_ctx->msm_points, &g_ctx->ntt_data, &g_ctx->ntt_scratch, &g_ctx->ntt_pro, &g_ctx->ntt_epi, &g_ctx->ntt_small,
                       &g_ctx->poly_tmp, &g_ctx->poly_stage, &g_ctx->poly_out };
    for (auto* b : bufs) b->release();
    if (g_ctx->inv_fix_fq) cudaFree(g_ctx->inv_fix_fq);
    for (auto& e : g_ctx->prof.ev) {
        if (e) cudaEventDestroy(e);
    }
    g_staging.release();
    if (g_ctx->pinned) cudaFreeHost(g_ctx->pinned);
    cudaEventDestroy(g_ctx->ev_a);
    cudaEventDestroy(g_ctx->ev_b);
    cudaEventDestroy(g_ctx->ev_fork);
    for (auto& e : g_ctx->ev_join) cudaEventDestroy(e);
    cudaEventDestroy(g_ctx->ev_part_fork);
    for (auto& e : g_ctx->ev_part_join) cudaEventDestroy(e);
    for (auto& a : g_ctx->part_stream) cudaStreamDestroy(a);
    cudaEventDestroy(g_ctx->last_use);
    cudaStreamDestroy(g_ctx->stream);
    cudaStreamDestroy(g_ctx->copy_stream);
    for (auto& a : g_ctx->aux_stream) cudaStreamDestroy(a);
    for (auto& e : g_ctx->ev_piece) cudaEventDestroy(e);
    delete g_ctx;
    g_ctx = nullptr;
}

const char* bbg_last_error(void) { return g_last_error.empty() ? g_last_error_global.c_str() : g_last_error.c_str(); }

int bbg_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}
uint64_t bbg_kernel_launches(void) { return g_ctx ? g_ctx->launches : 0; }
double bbg_last_device_ms(void) { return g_ctx ? g_ctx->last_kernel_ms : 0.0; }

static_assert(BBG_NUM_PHASES == PH_COUNT, "include/bbg.h BBG_NUM_PHASES must cover every Phase");
// totals of the host-pointer entry points since the library was loaded: for each of msm, ntt, srs, poly:
// calls, H2D bytes, D2H bytes (12 values)
int bbg_stats_totals(uint64_t* out12)
{
    if (!out12) return BBG_ERR_ARG;
    for (int i = 0; i < 4; ++i) {
        out12[3 * i] = g_stats.rows[i].calls;
        out12[3 * i + 1] = g_stats.rows[i].bytes_h2d;
        out12[3 * i + 2] = g_stats.rows[i].bytes_d2h;
    }
    return BBG_OK;
}

int bbg_profile(int enable)
{
    GET_CTX();
    ctx->prof.on = enable != 0;
    ctx->prof.n = 0;
    return BBG_OK;
}
int bbg_profile_read(double* ms, int n)
{
    GET_CTX();
    for (int i = 0; i < n; ++i) ms[i] = 0.0;
    Profiler& pr = ctx->prof;
    if (pr.n < 2) return BBG_OK;
    BBG_CUDA(cudaEventSynchronize(pr.ev[pr.n - 1]));
    for (int i = 0; i + 1 < pr.n; ++i) {
        float t = 0.f;
        BBG_CUDA(cudaEventElapsedTime(&t, pr.ev[i], pr.ev[i + 1]));
        if (pr.ids[i] >= 0 && pr.ids[i] < n) ms[pr.ids[i]] += t;
    }
    return BBG_OK;
}

void* bbg_malloc(size_t size)
{
    Context* ctx = nullptr;
    if (get_context(&ctx)) return nullptr;
    DeviceGuard dg(ctx->device);
    void* p = nullptr;
    if (cudaHostAlloc(&p, size ? size : 1, cudaHostAllocDefault) != cudaSuccess) {
        set_last_error("cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}
void bbg_free(void* ptr)
{
    if (ptr) cudaFreeHost(ptr);
}

// ---- Pippenger objects
static void* finish_obj(Context* ctx, PippengerObj* o, int rc)
{
    if (rc == BBG_OK) rc = msm_precompute_device(ctx, o->d_points, o->n, o->lv, ctx->stream);
    if (rc == BBG_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        set_last_error("stream sync failed while building the SRS");
        rc = BBG_ERR_CUDA;
    }
    if (rc != BBG_OK) {
        // ctx->mu is held by the caller: drop the object here instead of going through bbg_delete_pippenger
        auto it = std::find(g_pippengers.begin(), g_pippengers.end(), o);
        if (it != g_pippengers.end()) g_pippengers.erase(it);
        if (o->d_points) cudaFree(o->d_points);
        delete o;
        return nullptr;
    }
    return o;
}

void* bbg_new_pippenger(const uint8_t* points, size_t num_points)
{
    GET_CTX_PTR();
    StreamScope order(ctx, ctx->stream);
    PippengerObj* o = new_obj(ctx, num_points);
    if (!o) return nullptr;
    int rc = BBG_OK;
    if (num_points > 0) {
        rc = set_generator(ctx, o->d_points);
        if (!rc) rc = decode_raw_points(ctx, points, num_points - 1, o->d_points + 1);
    }
    return finish_obj(ctx, o, rc);
}

void* bbg_new_pippenger_from_path(const char* srs_dir, size_t num_points)
{
    {
        Context* ctx = nullptr;
        if (get_context(&ctx)) return nullptr;
    }
    std::vector<uint8_t> raw;
    if (read_transcript_raw(srs_dir, num_points, raw)) return nullptr;
    return bbg_new_pippenger(raw.data(), num_points);
}

void* bbg_new_pippenger_from_table(const void* table2n, size_t num_points)
{
    GET_CTX_PTR();
    StreamScope order(ctx, ctx->stream);
    PippengerObj* o = new_obj(ctx, num_points);
    if (!o) return nullptr;
    o->host_table = table2n;
    int rc = num_points ? upload_even_entries(ctx, table2n, num_points, o->d_points) : BBG_OK;
    return finish_obj(ctx, o, rc);
}

void* bbg_new_pippenger_from_points(const void* points, size_t num_points)
{
    GET_CTX_PTR();
    StreamScope order(ctx, ctx->stream);
    PippengerObj* o = new_obj(ctx, num_points);
    if (!o) return nullptr;
    int rc = BBG_OK;
    if (num_points && cudaMemcpyAsync(o->d_points, points, num_points * 64, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
        set_last_error("H2D copy of points failed");
        rc = BBG_ERR_CUDA;
    }
    return finish_obj(ctx, o, rc);
}

void* bbg_new_pippenger_from_device_points(const void* d_points, size_t num_points)
{
    GET_CTX_PTR();
    StreamScope order(ctx, ctx->stream);
    PippengerObj* o = new_obj(ctx, num_points);
    if (!o) return nullptr;
    int rc = BBG_OK;
    if (num_points && cudaMemcpyAsync(o->d_points, d_points, num_points * 64, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) {
        set_last_error("D2D copy of points failed");
        rc = BBG_ERR_CUDA;
    }
    return finish_obj(ctx, o, rc);
}

int bbg_pippenger_bind_host_table(void* pippenger, const void* table2n)
{
    GET_CTX();
    PippengerObj* o = reinterpret_cast<PippengerObj*>(pippenger);
    if (!o) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    o->host_table = table2n;
    return BBG_OK;
}

int bbg_set_auto_adopt(int enable)
{
    GET_CTX();
    g_auto_adopt = enable ? 1 : 0;
    return BBG_OK;
}

void bbg_delete_pippenger(void* pippenger)
{
    PippengerObj* o = reinterpret_cast<PippengerObj*>(pippenger);
    if (!o) return;
    std::lock_guard<std::mutex> glk(g_ctx_mu);
    if (g_ctx == nullptr) return; // after bbg_shutdown every handle is already gone
    std::lock_guard<std::mutex> lk(g_ctx->mu);
    auto it = std::find(g_pippengers.begin(), g_pippengers.end(), o);
    if (it == g_pippengers.end()) return; // stale handle (freed by bbg_shutdown or deleted twice): never dereferenced
    g_pippengers.erase(it);
    DeviceGuard dg(g_ctx->device);
    if (o->d_points) cudaFree(o->d_points);
    delete o;
}

size_t bbg_pippenger_num_points(void* pippenger) { return pippenger ? reinterpret_cast<PippengerObj*>(pippenger)->n : 0; }
const void* bbg_pippenger_device_points(void* pippenger) { return pippenger ? reinterpret_cast<PippengerObj*>(pippenger)->d_points : nullptr; }
unsigned bbg_pippenger_window_bits(void* pippenger) { return pippenger ? reinterpret_cast<PippengerObj*>(pippenger)->lv.c : 0; }
unsigned bbg_pippenger_levels(void* pippenger) { return pippenger ? reinterpret_cast<PippengerObj*>(pippenger)->lv.L : 0; }

int bbg_pippenger_get_point_table(void* pippenger, void* table2n_out)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    PippengerObj* o = reinterpret_cast<PippengerObj*>(pippenger);
    if (!o || !table2n_out) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    if (o->n == 0) return BBG_OK;
    void* d_table = nullptr;
    BBG_CUDA(cudaMalloc(&d_table, o->n * 128));
    int rc = point_table_device(ctx, o->d_points, o->n, d_table, ctx->stream);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(table2n_out, d_table, o->n * 128, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            set_last_error(cudaGetErrorString(e));
            rc = BBG_ERR_CUDA;
        }
    }
    cudaFree(d_table);
    return rc;
}

static int msm_host_scalars(Context* ctx, const void* scalars, size_t n, const affine_t* d_points, size_t stride, const MsmLevels& lv,
                            size_t base, void* result)
{
    int rc;
    MsmWorkspace& ws = ctx->msm_ws[0];
    StreamScope order(ctx, ctx->stream);
    if ((rc = ws.result.reserve(96))) return rc;
    // resident mirror of the scalar array (an ifft result, a slice of the quotient polynomial)?
    void* d_res = nullptr;
    bool hit = false;
    if (n && (rc = resident_acquire(ctx, scalars, n * 32, true, &d_res, &hit, ctx->stream))) return rc;
    StatScope stat(STAT_MSM, ctx, hit ? 0 : n * 32, 96);
    if (d_res != nullptr) {
        DeviceTimer tm(ctx);
        if ((rc = msm_device(ctx, ws, d_res, n, d_points, stride, lv, base, ws.result.p, ctx->stream))) return rc;
        tm.stop();
        BBG_CUDA(cudaMemcpyAsync(result, ws.result.p, 96, cudaMemcpyDeviceToHost, ctx->stream));
        return tm.finish();
    }
    if ((rc = ws.scalars.reserve(std::max<size_t>(n, 1) * 32))) return rc;
    // Large pinned scalar arrays go up in pieces on the copy stream; the histogram pass of msm_device chases them piece by
    // piece (it is the only phase that can start before every scalar is on the device).
    MsmArrival arrival;
    const size_t pieces = 4;
    const bool piecewise = n >= (1u << 18) && !Staging::pageable(scalars);
    if (piecewise) {
        // order the copies after whatever last used the scalar buffer on the work stream
        BBG_CUDA(cudaEventRecord(ctx->ev_piece[pieces], ctx->stream));
        BBG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_piece[pieces], 0));
        arrival.ready = ctx->ev_piece;
        arrival.count = pieces;
        arrival.piece = ((n + pieces - 1) / pieces + 255) & ~(size_t)255;
        for (size_t k = 0; k < pieces; ++k) {
            const size_t lo = k * arrival.piece;
            const size_t len = lo < n ? std::min(arrival.piece, n - lo) : 0;
            if (len) {
                BBG_CUDA(cudaMemcpyAsync((char*)ws.scalars.p + lo * 32, (const char*)scalars + lo * 32, len * 32,
                                         cudaMemcpyHostToDevice, ctx->copy_stream));
            }
            BBG_CUDA(cudaEventRecord(ctx->ev_piece[k], ctx->copy_stream));
        }
    } else if (n && (rc = g_staging.h2d(ws.scalars.p, scalars, n * 32, ctx->stream))) {
        return rc;
    }
    DeviceTimer tm(ctx);
    if ((rc = msm_device(ctx, ws, ws.scalars.p, n, d_points, stride, lv, base, ws.result.p, ctx->stream,
                         piecewise ? &arrival : nullptr))) return rc;
    tm.stop();
    BBG_CUDA(cudaMemcpyAsync(result, ws.result.p, 96, cudaMemcpyDeviceToHost, ctx->stream));
    return tm.finish();
}

int bbg_pippenger_unsafe(void* pippenger, const void* scalars, size_t from, size_t range, void* result)
{
    GET_CTX();
    PippengerObj* o = reinterpret_cast<PippengerObj*>(pippenger);
    if (!o || !result || (range && !scalars)) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    if (from + range > o->n) {
        set_last_error("pippenger_unsafe: [from, from+range) exceeds the SRS");
        return BBG_ERR_SRS;
    }
    return msm_host_scalars(ctx, scalars, range, o->d_points, 1, o->lv, from, result);
}

int bbg_pippenger_unsafe_dev(void* pippenger, const void* d_scalars, size_t from, size_t range, void* d_result, void* stream)
{
    GET_CTX();
    PippengerObj* o = reinterpret_cast<PippengerObj*>(pippenger);
    if (!o || !d_result) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    if (from + range > o->n) {
        set_last_error("pippenger_unsafe: [from, from+range) exceeds the SRS");
        return BBG_ERR_SRS;
    }
    StreamScope order(ctx, (cudaStream_t)stream);
    return msm_device(ctx, ctx->msm_ws[0], d_scalars, range, o->d_points, 1, o->lv, from, d_result, (cudaStream_t)stream);
}

// Several MSMs over the same bases in one call (the prover commits to its four wire polynomials, then to the four
// slices of the quotient polynomial, back to back: prover.cpp:66-82, 84-135).  MSM i runs on stream i mod 4 with workspace
// i mod 4, so the latency-bound tail of one MSM (slot merge, bucket reduction: a few hundred warps) overlaps the
// throughput-bound bucket accumulation of the next ones.
static int msm_batch(Context* ctx, PippengerObj* o, const void* const* scalars, bool device_scalars, size_t count, size_t from,
                     size_t range, void* results, bool device_results, cudaStream_t st0)
{
    int rc;
    StreamScope order(ctx, st0);
    constexpr int WAYS = Context::BATCH_WAYS;
    cudaStream_t st[WAYS];
    st[0] = st0;
    for (int k = 1; k < WAYS; ++k) st[k] = ctx->aux_stream[k - 1];
    // Small MSMs are FUSED, up to four per pass of the kernels (msm_device with an MsmBatch: the digits of vector m go to
    // their own bucket sets, one sort / accumulate / merge / reduce chain for all of them): the latency-bound tail, which is
    // most of a 2^16-point MSM, is paid once per group instead of once per MSM.  Large MSMs are throughput bound and keep
    // one chain each; either way consecutive chains go to different streams and workspaces.
    static const size_t fused_max = [] {
        const char* v = getenv("BBG_MSM_FUSED_BATCH_MAX_LOG2"); // 0 disables fusing
        const unsigned lg = v && *v ? (unsigned)atoi(v) : 17u;
        return lg == 0 ? (size_t)0 : (size_t)1 << std::min(lg, 30u);
    }();
    const size_t group = (count > 1 && range > 0 && range <= fused_max) ? 4 : 1;
    const size_t chains = (count + group - 1) / group;
    const int used = (int)std::min<size_t>(chains, WAYS);
    if (used > 1) {
        BBG_CUDA(cudaEventRecord(ctx->ev_fork, st0));
        for (int k = 1; k < used; ++k) BBG_CUDA(cudaStreamWaitEvent(st[k], ctx->ev_fork, 0));
    }
    uint64_t h2d = 0;
    StatScope stat(STAT_MSM, ctx, 0, device_results ? 0 : 96 * count);
    // Host results go through a pinned staging buffer: a device-to-host copy into pageable memory blocks the calling
    // thread until the MSM before it has finished, which would serialise the batch on the host (the next MSM would
    // not even be queued yet) and throw the overlap away.
    if (!device_results && count * 96 > ctx->pinned_cap) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_cap = 0;
        const size_t cap = std::max<size_t>(count * 96, 4096);
        BBG_CUDA(cudaHostAlloc(&ctx->pinned, cap, cudaHostAllocDefault));
        ctx->pinned_cap = cap;
    }
    for (size_t ch = 0; ch < chains; ++ch) {
        MsmWorkspace& ws = ctx->msm_ws[ch % WAYS];
        cudaStream_t s = st[ch % WAYS];
        const size_t first = ch * group;
        const size_t members = std::min(group, count - first);
        MsmBatch mb;
        mb.count = (unsigned)members;
        for (auto& q : mb.scalars) q = nullptr;
        if (!device_scalars && (rc = ws.scalars.reserve(std::max<size_t>(range, 1) * 32 * members))) return rc;
        for (size_t m = 0; m < members; ++m) {
            const size_t i = first + m;
            const void* d_sc = scalars[i];
            if (!device_scalars) {
                void* d_res = nullptr;
                bool hit = false;
                if (range && (rc = resident_acquire(ctx, scalars[i], range * 32, true, &d_res, &hit, s))) return rc;
                if (d_res != nullptr) {
                    d_sc = d_res;
                    if (!hit) h2d += range * 32;
                } else {
                    void* slot = (char*)ws.scalars.p + m * range * 32;
                    if (range && (rc = g_staging.h2d(slot, scalars[i], range * 32, s))) return rc;
                    d_sc = slot;
                    h2d += range * 32;
                }
            }
            mb.scalars[m] = d_sc;
        }
        void* d_out = device_results ? (char*)results + first * 96 : nullptr;
        if (!device_results) {
            if ((rc = ws.result.reserve(96 * 4))) return rc;
            d_out = ws.result.p;
        }
        if ((rc = msm_device(ctx, ws, mb.scalars[0], range, o->d_points, 1, o->lv, from, d_out, s, nullptr, /*allow_parts=*/chains == 1,
                             members > 1 ? &mb : nullptr)))
            return rc;
        if (!device_results) BBG_CUDA(cudaMemcpyAsync((char*)ctx->pinned + first * 96, d_out, 96 * members, cudaMemcpyDeviceToHost, s));
    }
    if (stat.row) stat.row->bytes_h2d += h2d;
    stat.h2d = h2d;
    for (int k = 1; k < used; ++k) {
        BBG_CUDA(cudaEventRecord(ctx->ev_join[k - 1], st[k]));
        BBG_CUDA(cudaStreamWaitEvent(st0, ctx->ev_join[k - 1], 0));
    }
    if (!device_results) {
        BBG_CUDA(cudaStreamSynchronize(st0));
        memcpy(results, ctx->pinned, count * 96);
    }
    return BBG_OK;
}

int bbg_pippenger_unsafe_batch(void* pippenger, const void* const* scalars, size_t count, size_t from, size_t range, void* results)
{
    GET_CTX();
    PippengerObj* o = reinterpret_cast<PippengerObj*>(pippenger);
    if (!o || (count && (!scalars || !results))) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    if (from + range > o->n) {
        set_last_error("pippenger_unsafe_batch: [from, from+range) exceeds the SRS");
        return BBG_ERR_SRS;
    }
    return msm_batch(ctx, o, scalars, false, count, from, range, results, false, ctx->stream);
}

int bbg_pippenger_unsafe_batch_dev(void* pippenger, const void* const* d_scalars, size_t count, size_t from, size_t range,
                                   void* d_results, void* stream)
{
    GET_CTX();
    PippengerObj* o = reinterpret_cast<PippengerObj*>(pippenger);
    if (!o || (count && (!d_scalars || !d_results))) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    if (from + range > o->n) {
        set_last_error("pippenger_unsafe_batch: [from, from+range) exceeds the SRS");
        return BBG_ERR_SRS;
    }
    return msm_batch(ctx, o, d_scalars, true, count, from, range, d_results, true, (cudaStream_t)stream);
}

// the Pippenger object whose adopted host table contains `points_table2n` (ProverReferenceString::get_monomials() or
// monomials + 2 * from); *from receives the offset in points
static PippengerObj* find_adopted(const void* points_table2n, size_t num_points, size_t* from)
{
    const uint8_t* p = reinterpret_cast<const uint8_t*>(points_table2n);
    for (auto* o : g_pippengers) {
        const uint8_t* base = reinterpret_cast<const uint8_t*>(o->host_table);
        if (base && p >= base && p < base + o->n * 128 && (size_t)(p - base) % 128 == 0) {
            const size_t f = (size_t)(p - base) / 128;
            if (f + num_points <= o->n) {
                *from = f;
                return o;
            }
        }
    }
    return nullptr;
}

// batch flavour of bbg_pippenger(): `count` scalar arrays over the SAME interleaved table, which must belong to an
// adopted Pippenger object (bbg_pippenger_bind_host_table / bbg_new_pippenger_from_table); BBG_ERR_ARG otherwise.
int bbg_pippenger_batch(const void* const* scalars, size_t count, const void* points_table2n, size_t num_points, void* results)
{
    GET_CTX();
    if (count && (!scalars || !results || !points_table2n)) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    size_t from = 0;
    PippengerObj* o = find_adopted(points_table2n, num_points, &from);
    if (o == nullptr) {
        set_last_error("pippenger_batch: the point table is not resident (adopt it with bbg_pippenger_bind_host_table)");
        return BBG_ERR_ARG;
    }
    return msm_batch(ctx, o, scalars, false, count, from, num_points, results, false, ctx->stream);
}

int bbg_pippenger(const void* scalars, const void* points_table2n, size_t num_points, int handle_edge_cases, void* result)
{
    (void)handle_edge_cases; // the device path always handles doubling / infinity
    GET_CTX();
    if (!result || (num_points && (!scalars || !points_table2n))) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    // resident copy?  (points may be monomials + 2*from, bb/.../pippenger.cpp:27-31)
    {
        size_t from = 0;
        if (PippengerObj* o = find_adopted(points_table2n, num_points, &from)) {
            return msm_host_scalars(ctx, scalars, num_points, o->d_points, 1, o->lv, from, result);
        }
    }
    int rc;
    if (g_auto_adopt < 0) {
        const char* e = getenv("BBG_AUTO_ADOPT");
        g_auto_adopt = (e && *e && *e != '0') ? 1 : 0;
    }
    if (g_auto_adopt == 1 && num_points >= 4096) {
        // first sight of an (immutable) SRS table: make it resident, with its fixed-base levels
        PippengerObj* o = new_obj(ctx, num_points);
        if (!o) return BBG_ERR_CUDA;
        o->host_table = points_table2n;
        rc = upload_even_entries(ctx, points_table2n, num_points, o->d_points);
        if (finish_obj(ctx, o, rc) == nullptr) return BBG_ERR_CUDA;
        return msm_host_scalars(ctx, scalars, num_points, o->d_points, 1, o->lv, 0, result);
    }
    if ((rc = ctx->msm_points.reserve(std::max<size_t>(num_points, 1) * 64))) return rc;
    if (num_points && (rc = upload_even_entries(ctx, points_table2n, num_points, (affine_t*)ctx->msm_points.p))) return rc;
    return msm_host_scalars(ctx, scalars, num_points, (const affine_t*)ctx->msm_points.p, 1, MsmLevels(), 0, result);
}

int bbg_msm_points(const void* scalars, const void* points, size_t num_points, void* result)
{
    GET_CTX();
    if (!result || (num_points && (!scalars || !points))) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    int rc;
    if ((rc = ctx->msm_points.reserve(std::max<size_t>(num_points, 1) * 64))) return rc;
    if (num_points) BBG_CUDA(cudaMemcpyAsync(ctx->msm_points.p, points, num_points * 64, cudaMemcpyHostToDevice, ctx->stream));
    return msm_host_scalars(ctx, scalars, num_points, (const affine_t*)ctx->msm_points.p, 1, MsmLevels(), 0, result);
}

int bbg_msm_points_dev(const void* d_scalars, const void* d_points, size_t point_stride, size_t num_points, void* d_result, void* stream)
{
    GET_CTX();
    StreamScope order(ctx, (cudaStream_t)stream);
    return msm_device(ctx, ctx->msm_ws[0], d_scalars, num_points, d_points, point_stride ? point_stride : 1, MsmLevels(), 0, d_result,
                      (cudaStream_t)stream);
}

int bbg_generate_pippenger_point_table(const void* points, void* table, size_t num_points)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (num_points == 0) return BBG_OK;
    void *d_pts = nullptr, *d_table = nullptr;
    BBG_CUDA(cudaMalloc(&d_pts, num_points * 64));
    cudaError_t e = cudaMalloc(&d_table, num_points * 128);
    if (e != cudaSuccess) {
        cudaFree(d_pts);
        set_last_error(cudaGetErrorString(e));
        return BBG_ERR_CUDA;
    }
    int rc = BBG_OK;
    e = cudaMemcpyAsync(d_pts, points, num_points * 64, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = point_table_device(ctx, d_pts, num_points, d_table, ctx->stream);
    if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(table, d_table, num_points * 128, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_pts);
    cudaFree(d_table);
    if (e != cudaSuccess) {
        set_last_error(cudaGetErrorString(e));
        return BBG_ERR_CUDA;
    }
    return rc;
}

int bbg_g1_sum(const void* elements, size_t num_points, void* result)
{
    GET_CTX();
    int rc;
    MsmWorkspace& ws = ctx->msm_ws[0];
    StreamScope order(ctx, ctx->stream);
    if ((rc = ctx->msm_points.reserve(std::max<size_t>(num_points, 1) * 96))) return rc;
    if ((rc = ws.result.reserve(96))) return rc;
    if (num_points) BBG_CUDA(cudaMemcpyAsync(ctx->msm_points.p, elements, num_points * 96, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = g1_sum_device(ctx, ctx->msm_points.p, num_points, ws.result.p, ctx->stream))) return rc;
    BBG_CUDA(cudaMemcpyAsync(result, ws.result.p, 96, cudaMemcpyDeviceToHost, ctx->stream));
    BBG_CUDA(cudaStreamSynchronize(ctx->stream));
    return BBG_OK;
}
int bbg_g1_sum_dev(const void* d_elements, size_t num_points, void* d_result, void* stream)
{
    GET_CTX();
    StreamScope order(ctx, (cudaStream_t)stream);
    return g1_sum_device(ctx, d_elements, num_points, d_result, (cudaStream_t)stream);
}

int bbg_read_g1_elements_from_buffer(void* elements, const char* buffer, size_t buffer_size)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    size_t n = buffer_size / 64;
    if (n == 0) return BBG_OK;
    void* d_out = nullptr;
    BBG_CUDA(cudaMalloc(&d_out, n * 64));
    int rc = decode_raw_points(ctx, reinterpret_cast<const uint8_t*>(buffer), n, (affine_t*)d_out);
    cudaError_t e = cudaSuccess;
    if (!rc) e = cudaMemcpyAsync(elements, d_out, n * 64, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) {
        set_last_error(cudaGetErrorString(e));
        return BBG_ERR_CUDA;
    }
    return rc;
}

int bbg_read_transcript_g1(void* monomials, size_t degree, const char* srs_dir)
{
    if (degree == 0) return BBG_OK;
    std::vector<uint8_t> raw;
    int rc = read_transcript_raw(srs_dir, degree, raw);
    if (rc) return rc;
    memcpy(monomials, G1_ONE_X, 32);
    memcpy(reinterpret_cast<uint8_t*>(monomials) + 32, G1_ONE_Y, 32);
    return bbg_read_g1_elements_from_buffer(reinterpret_cast<uint8_t*>(monomials) + 64, (const char*)raw.data(), raw.size());
}

// ---- NTT
int bbg_ntt_dev(void* d_coeffs, size_t n, int kind, size_t generator_size, const void* constant, void* stream)
{
    GET_CTX();
    StreamScope order(ctx, (cudaStream_t)stream);
    unsigned lg;
    int rc = log2_exact(n, lg);
    if (rc) return rc;
    bool inverse;
    NttScale pro, epi;
    if ((rc = ntt_kind_params(kind, lg, generator_size, constant, inverse, pro, epi))) return rc;
    return ntt_device(ctx, d_coeffs, d_coeffs, lg, inverse, pro, epi, 0, 0, (cudaStream_t)stream);
}

// host-pointer transform; keep_on_device: leave the result in the array's device mirror (host copy stale) when it has one
static int ntt_host(Context* ctx, void* coeffs, size_t n, int kind, size_t generator_size, const void* constant, bool keep_on_device)
{
    StreamScope order(ctx, ctx->stream);
    unsigned lg;
    int rc = log2_exact(n, lg);
    if (rc) return rc;
    bool inverse;
    NttScale pro, epi;
    if ((rc = ntt_kind_params(kind, lg, generator_size, constant, inverse, pro, epi))) return rc;
    void* d_res = nullptr;
    bool hit = false;
    if ((rc = resident_acquire(ctx, coeffs, n * 32, true, &d_res, &hit, ctx->stream))) return rc;
    const bool keep = keep_on_device && d_res != nullptr;
    StatScope stat(STAT_NTT, ctx, hit ? 0 : n * 32, keep ? 0 : n * 32);
    if (d_res != nullptr) {
        // resident mirror: transform it in place and bring the result home; the mirror stays valid for the next call
        DeviceTimer tm(ctx);
        if ((rc = ntt_device(ctx, d_res, d_res, lg, inverse, pro, epi, 0, 0, ctx->stream))) return rc;
        tm.stop();
        if ((rc = resident_commit(ctx, coeffs, n * 32, !keep, ctx->stream))) return rc;
        return tm.finish();
    }
    if ((rc = ctx->ntt_data.reserve(n * 32))) return rc;
    if ((rc = g_staging.h2d(ctx->ntt_data.p, coeffs, n * 32, ctx->stream))) return rc;
    DeviceTimer tm(ctx);
    if ((rc = ntt_device(ctx, ctx->ntt_data.p, ctx->ntt_data.p, lg, inverse, pro, epi, 0, 0, ctx->stream))) return rc;
    tm.stop();
    if ((rc = g_staging.d2h(coeffs, ctx->ntt_data.p, n * 32, ctx->stream))) return rc;
    return tm.finish();
}

static void ntt_factor(unsigned log_n, unsigned& first_bits, unsigned& last_bits)
{
    // must mirror ntt_device's factorisation
    const unsigned num_passes = log_n <= 16 ? 2 : (log_n <= 24 ? 3 : 4);
    first_bits = log_n / num_passes + (0 < log_n % num_passes ? 1 : 0);
    last_bits = log_n / num_passes + ((num_passes - 1) < log_n % num_passes ? 1 : 0);
}

int bbg_ntt(void* coeffs, size_t n, int kind, size_t generator_size, const void* constant)
{
    GET_CTX();
    return ntt_host(ctx, coeffs, n, kind, generator_size, constant, false);
}
int bbg_ntt_ex(void* coeffs, size_t n, int kind, size_t generator_size, const void* constant, unsigned flags)
{
    GET_CTX();
    const bool keep = (flags & BBG_KEEP_ON_DEVICE) != 0 || ((flags & BBG_KEEP_IF_AHEAD) != 0 && resident_is_ahead(ctx, coeffs, n * 32));
    return ntt_host(ctx, coeffs, n, kind, generator_size, constant, keep);
}

int bbg_ntt_dist_layout(size_t n, int world, unsigned* in_pos, unsigned* out_pos)
{
    unsigned lg, rb = 0;
    int rc = log2_exact(n, lg);
    if (rc) return rc;
    while ((1 << rb) < world) ++rb;
    unsigned first, last;
    ntt_factor(lg, first, last);
    if ((1 << rb) != world || lg < 12 || last < rb + 3 || first < rb + 3) {
        set_last_error("ntt_dist: world must be a power of two with n >= 2^12 and at most 2^(digit - 3) ranks");
        return BBG_ERR_ARG;
    }
    if (in_pos) *in_pos = last - rb;
    if (out_pos) *out_pos = first - rb;
    return BBG_OK;
}

int bbg_ntt_dist_dev(const void* d_src, void* d_dst, size_t n, int kind, size_t generator_size, const void* constant, int rank,
                     int world, int phase, void* stream)
{
    GET_CTX();
    StreamScope order(ctx, (cudaStream_t)stream);
    unsigned lg;
    int rc = log2_exact(n, lg);
    if (rc) return rc;
    unsigned ip, op;
    if ((rc = bbg_ntt_dist_layout(n, world, &ip, &op))) return rc;
    if (rank < 0 || rank >= world || (phase != 0 && phase != 1) || (phase == 1 && d_src == d_dst)) {
        set_last_error("ntt_dist: bad rank / phase / aliasing");
        return BBG_ERR_ARG;
    }
    bool inverse;
    NttScale pro, epi;
    if ((rc = ntt_kind_params(kind, lg, generator_size, constant, inverse, pro, epi))) return rc;
    NttDist dist;
    dist.rank = (unsigned)rank;
    dist.phase = phase;
    while ((1 << dist.rank_bits) < world) ++dist.rank_bits;
    if (dist.rank_bits == 0) {
        // a single rank: phase 0 does the whole transform, phase 1 is a copy
        if (phase == 0) return ntt_device(ctx, d_src, d_dst, lg, inverse, pro, epi, 0, 0, (cudaStream_t)stream);
        BBG_CUDA(cudaMemcpyAsync(d_dst, d_src, n * 32, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return BBG_OK;
    }
    return ntt_device(ctx, d_src, d_dst, lg, inverse, pro, epi, 0, 0, (cudaStream_t)stream, dist);
}

// Phase 0 of the multi-GPU NTT with the exchange FUSED into its last pass: the intermediate is written straight into the
// receive buffers of the owning ranks over NVLink peer memory (peer_recv[r] = rank r's buffer of n / world elements, opened
// with bbg_peer_buffer_open; peer_recv[rank] = this rank's own).  d_work: n / world elements of local scratch for the
// earlier passes.  world <= 8.  The caller must make sure (a stream-ordered all-reduce is enough) that every rank has finished
// this call before any rank runs phase 1 on its receive buffer.
int bbg_ntt_dist_fused_dev(const void* d_src, void* d_work, void* const* peer_recv, size_t n, int kind, size_t generator_size,
                           const void* constant, int rank, int world, void* stream)
{
    GET_CTX();
    StreamScope order(ctx, (cudaStream_t)stream);
    unsigned lg;
    int rc = log2_exact(n, lg);
    if (rc) return rc;
    unsigned ip, op;
    if ((rc = bbg_ntt_dist_layout(n, world, &ip, &op))) return rc;
    if (rank < 0 || rank >= world || world < 2 || world > 8 || peer_recv == nullptr) {
        set_last_error("ntt_dist_fused: 2..8 ranks and a table of peer receive buffers");
        return BBG_ERR_ARG;
    }
    bool inverse;
    NttScale pro, epi;
    if ((rc = ntt_kind_params(kind, lg, generator_size, constant, inverse, pro, epi))) return rc;
    NttDist dist;
    dist.rank = (unsigned)rank;
    dist.phase = 0;
    dist.peer_recv = peer_recv;
    while ((1 << dist.rank_bits) < world) ++dist.rank_bits;
    return ntt_device(ctx, d_src, d_work, lg, inverse, pro, epi, 0, 0, (cudaStream_t)stream, dist);
}

// The multi-GPU transform on NATURAL contiguous blocks (rank q holds x[q n / W, (q + 1) n / W) and ends with
// X[q n / W, (q + 1) n / W)), entirely over peer memory: phase 0 loads its sliced sub-array from the owners' input blocks
// in its first pass and stores into the owners' receive buffers in its last; phase 1 transforms the own receive buffer and
// stores every output to the owner of its natural index.  No all-to-all, no re-distribution.  The caller orders the phases
// across ranks (a stream-ordered one-word all-reduce each): inputs written -> phase 0 -> phase 1 -> outputs readable.
int bbg_ntt_dist_natural_dev(void* const* peer_in, void* d_work, void* const* peer_recv, void* const* peer_out, size_t n, int kind,
                             size_t generator_size, const void* constant, int rank, int world, int phase, void* stream)
{
    GET_CTX();
    StreamScope order(ctx, (cudaStream_t)stream);
    unsigned lg;
    int rc = log2_exact(n, lg);
    if (rc) return rc;
    unsigned ip, op;
    if ((rc = bbg_ntt_dist_layout(n, world, &ip, &op))) return rc;
    if (rank < 0 || rank >= world || world < 2 || world > 8 || (phase != 0 && phase != 1) || peer_recv == nullptr ||
        (phase == 0 && (peer_in == nullptr || d_work == nullptr)) || (phase == 1 && peer_out == nullptr)) {
        set_last_error("ntt_dist_natural: 2..8 ranks, phase 0 (peer_in, d_work, peer_recv) or 1 (peer_recv, peer_out)");
        return BBG_ERR_ARG;
    }
    bool inverse;
    NttScale pro, epi;
    if ((rc = ntt_kind_params(kind, lg, generator_size, constant, inverse, pro, epi))) return rc;
    NttDist dist;
    dist.rank = (unsigned)rank;
    dist.phase = phase;
    while ((1 << dist.rank_bits) < world) ++dist.rank_bits;
    if (phase == 0) {
        dist.peer_src = peer_in;
        dist.peer_recv = peer_recv;
        return ntt_device(ctx, peer_in[rank], d_work, lg, inverse, pro, epi, 0, 0, (cudaStream_t)stream, dist);
    }
    dist.peer_out = peer_out;
    return ntt_device(ctx, peer_recv[rank], peer_out[rank], lg, inverse, pro, epi, 0, 0, (cudaStream_t)stream, dist);
}

// ---- peer-mapped buffers (CUDA IPC): memory another rank's kernels can store into over NVLink
int bbg_peer_buffer_alloc(size_t bytes, void** d_ptr, void* ipc_handle64)
{
    GET_CTX();
    if (!d_ptr || !ipc_handle64 || bytes == 0) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C-ABI passes IPC handles as 64 opaque bytes");
    void* p = nullptr;
    BBG_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_last_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
        return BBG_ERR_CUDA;
    }
    memcpy(ipc_handle64, &h, 64);
    *d_ptr = p;
    return BBG_OK;
}
int bbg_peer_buffer_open(const void* ipc_handle64, void** d_ptr)
{
    GET_CTX();
    if (!d_ptr || !ipc_handle64) {
        set_last_error("null argument");
        return BBG_ERR_ARG;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, 64);
    BBG_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BBG_OK;
}
int bbg_peer_buffer_close(void* d_ptr)
{
    GET_CTX();
    if (d_ptr) BBG_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return BBG_OK;
}
int bbg_peer_buffer_free(void* d_ptr)
{
    GET_CTX();
    if (d_ptr) BBG_CUDA(cudaFree(d_ptr));
    return BBG_OK;
}

static int coset_fft_ext_device(Context* ctx, void* d_coeffs, unsigned lg, size_t ext, cudaStream_t st)
{
    unsigned lg_ext = 0;
    while (((size_t)1 << lg_ext) < ext) ++lg_ext;
    if (((size_t)1 << lg_ext) != ext || lg + lg_ext > 28) {
        set_last_error("coset_fft: domain_extension must be a power of two with n*ext <= 2^28");
        return BBG_ERR_ARG;
    }
    const size_t n = (size_t)1 << lg;
    int rc;
    // the interleaved output overwrites the input, so keep a copy of the n input coefficients
    if ((rc = ctx->ntt_small.reserve(n * 32))) return rc;
    BBG_CUDA(cudaMemcpyAsync(ctx->ntt_small.p, d_coeffs, n * 32, cudaMemcpyDeviceToDevice, st));
    // coset k uses generator g * w_{ext n}^k (polynomial_arithmetic.cpp:414-421)
    const hf::Fr prim = ntt_root_of_unity(lg + lg_ext);
    hf::Fr gk = coset_generator();
    for (size_t k = 0; k < ext; ++k) {
        NttScale pro, epi;
        pro.present = true;
        pro.start = hf::one();
        pro.has_shift = true;
        pro.shift = gk;
        pro.size = n;
        if ((rc = ntt_device(ctx, ctx->ntt_small.p, d_coeffs, lg, false, pro, epi, lg_ext, (unsigned)k, st))) return rc;
        gk = hf::mul(gk, prim);
    }
    return BBG_OK;
}

int bbg_coset_fft_ext_dev(void* d_coeffs, size_t n, size_t domain_extension, void* stream)
{
    GET_CTX();
    StreamScope order(ctx, (cudaStream_t)stream);
    unsigned lg;
    int rc = log2_exact(n, lg);
    if (rc) return rc;
    return coset_fft_ext_device(ctx, d_coeffs, lg, domain_extension, (cudaStream_t)stream);
}

int bbg_coset_fft_ext(void* coeffs, size_t n, size_t domain_extension)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    unsigned lg;
    int rc = log2_exact(n, lg);
    if (rc) return rc;
    StatScope stat(STAT_NTT, ctx, n * 32, n * domain_extension * 32);
    if ((rc = ctx->ntt_data.reserve(n * domain_extension * 32))) return rc;
    if ((rc = g_staging.h2d(ctx->ntt_data.p, coeffs, n * 32, ctx->stream))) return rc;
    DeviceTimer tm(ctx);
    if ((rc = coset_fft_ext_device(ctx, ctx->ntt_data.p, lg, domain_extension, ctx->stream))) return rc;
    tm.stop();
    if ((rc = g_staging.d2h(coeffs, ctx->ntt_data.p, n * domain_extension * 32, ctx->stream))) return rc;
    return tm.finish();
}

void* bbg_new_evaluation_domain(size_t circuit_size)
{
    DomainObj* d = new DomainObj();
    d->n = circuit_size;
    return d;
}
void bbg_delete_evaluation_domain(void* domain) { delete reinterpret_cast<DomainObj*>(domain); }
int bbg_ifft(void* coeffs, void* domain) { return bbg_ntt(coeffs, reinterpret_cast<DomainObj*>(domain)->n, BBG_IFFT, 0, nullptr); }
int bbg_coset_fft_with_generator_shift(void* coeffs, const void* constant, void* domain)
{
    return bbg_ntt(coeffs, reinterpret_cast<DomainObj*>(domain)->n, BBG_COSET_FFT_WITH_GENERATOR_SHIFT, 0, constant);
}

int bbg_domain_constants(size_t n, void* out6)
{
    unsigned lg;
    int rc = log2_exact(n, lg);
    if (rc) return rc;
    hf::Fr c[6];
    c[0] = hf::reduce(ntt_root_of_unity(lg));
    c[1] = hf::invert(c[0]);
    c[2] = hf::from_u64(n);
    c[3] = hf::invert(c[2]);
    c[4] = coset_generator();
    c[5] = hf::invert(c[4]);
    memcpy(out6, c, sizeof(c));
    return BBG_OK;
}

// ---- measurement / synthetic-input utilities
int bbg_bench_field_mul(int field, int iters, double* muls_per_second)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (!muls_per_second || iters <= 0) {
        set_last_error("bad argument");
        return BBG_ERR_ARG;
    }
    void* d_out = nullptr;
    BBG_CUDA(cudaMalloc(&d_out, 256 * 32));
    const int blocks = ctx->num_sms * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) { // first repetition warms up
        cudaEventRecord(ctx->ev_a, ctx->stream);
        if (field == 0) {
            k_bench_mul<FqParams><<<blocks, 256, 0, ctx->stream>>>((fq_t*)d_out, iters);
        } else {
            k_bench_mul<FrParams><<<blocks, 256, 0, ctx->stream>>>((fr_t*)d_out, iters);
        }
        ctx->launches += 1;
        cudaEventRecord(ctx->ev_b, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            cudaFree(d_out);
            set_last_error(cudaGetErrorString(e));
            return BBG_ERR_CUDA;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(d_out);
    *muls_per_second = (double)blocks * 256.0 * 2.0 * iters / (best * 1e-3);
    return BBG_OK;
}

int bbg_g1_add_affine_dev(const void* d_in, void* d_out, size_t n, const void* q_affine, void* stream)
{
    GET_CTX();
    StreamScope order(ctx, (cudaStream_t)stream);
    if (n == 0) return BBG_OK;
    affine_t q;
    memcpy(&q, q_affine, 64);
    k_g1_add_affine<<<div_up(n, 128), 128, 0, (cudaStream_t)stream>>>((const affine_t*)d_in, (affine_t*)d_out, n, q);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

// ---- probes
int bbg_field_op(int field, int op, const void* a, const void* b, void* out, size_t n)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (n == 0) return BBG_OK;
    void *da = nullptr, *db = nullptr, *dout = nullptr;
    BBG_CUDA(cudaMalloc(&da, n * 32));
    BBG_CUDA(cudaMalloc(&dout, n * 32));
    if (b) BBG_CUDA(cudaMalloc(&db, n * 32));
    BBG_CUDA(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (b) BBG_CUDA(cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (field == 0) {
        k_field_op<FqParams><<<div_up(n, 128), 128, 0, ctx->stream>>>(op, (const fq_t*)da, (const fq_t*)db, (fq_t*)dout, n);
    } else {
        k_field_op<FrParams><<<div_up(n, 128), 128, 0, ctx->stream>>>(op, (const fr_t*)da, (const fr_t*)db, (fr_t*)dout, n);
    }
    ctx->launches += 1;
    BBG_CUDA(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    BBG_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(da);
    cudaFree(dout);
    if (db) cudaFree(db);
    return BBG_OK;
}

int bbg_field_op_dev(int field, int op, const void* d_a, const void* d_b, void* d_out, size_t n, void* stream)
{
    GET_CTX();
    StreamScope order(ctx, (cudaStream_t)stream);
    if (n == 0) return BBG_OK;
    if (field == 0) {
        k_field_op<FqParams><<<div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(op, (const fq_t*)d_a, (const fq_t*)d_b, (fq_t*)d_out, n);
    } else {
        k_field_op<FrParams><<<div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(op, (const fr_t*)d_a, (const fr_t*)d_b, (fr_t*)d_out, n);
    }
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

int bbg_g1_normalize(const void* elements, size_t n, void* affine_out)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (n == 0) return BBG_OK;
    void *d_in = nullptr, *d_out = nullptr;
    BBG_CUDA(cudaMalloc(&d_in, n * 96));
    if (cudaMalloc(&d_out, n * 64) != cudaSuccess) {
        cudaFree(d_in);
        set_last_error("cudaMalloc failed");
        return BBG_ERR_CUDA;
    }
    cudaError_t e = cudaMemcpyAsync(d_in, elements, n * 96, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        k_g1_normalize<<<div_up(n, 64), 64, 0, ctx->stream>>>((const jac_t*)d_in, (affine_t*)d_out, n);
        ctx->launches += 1;
        e = cudaMemcpyAsync(affine_out, d_out, n * 64, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) {
        set_last_error(cudaGetErrorString(e));
        return BBG_ERR_CUDA;
    }
    return BBG_OK;
}

int bbg_g1_op(int op, const void* a, const void* b, void* out, size_t n)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    if (n == 0) return BBG_OK;
    const size_t bsz = op == 0 ? 64 : 96;
    void *da = nullptr, *db = nullptr, *dout = nullptr;
    BBG_CUDA(cudaMalloc(&da, n * 96));
    BBG_CUDA(cudaMalloc(&dout, n * 96));
    if (b) BBG_CUDA(cudaMalloc(&db, n * bsz));
    BBG_CUDA(cudaMemcpyAsync(da, a, n * 96, cudaMemcpyHostToDevice, ctx->stream));
    if (b) BBG_CUDA(cudaMemcpyAsync(db, b, n * bsz, cudaMemcpyHostToDevice, ctx->stream));
    k_g1_op<<<div_up(n, 64), 64, 0, ctx->stream>>>(op, (const jac_t*)da, db, (jac_t*)dout, n);
    ctx->launches += 1;
    BBG_CUDA(cudaMemcpyAsync(out, dout, n * 96, cudaMemcpyDeviceToHost, ctx->stream));
    BBG_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(da);
    cudaFree(dout);
    if (db) cudaFree(db);
    return BBG_OK;
}

} // extern "C"
