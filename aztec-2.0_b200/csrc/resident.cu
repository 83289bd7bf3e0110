// resident.cu -- device mirrors of caller-owned host arrays ("resident polynomials").
//
// barretenberg's prover hands the hot path the SAME host array several times in a row: a wire polynomial is
// ifft'd (work_queue.hpp:272-276), committed to (:213-243) and coset-FFT'd (:260-270); the quotient polynomial is
// divided, inverse-transformed and committed to in four slices (prover.cpp:326-343, 84-135).  Through a plain
// host-pointer C-ABI each of those calls pays a pageable H2D (and often a D2H) of the whole array, which is
// 95 % of the wall time of the NTT calls of a join-split proof.  With residency ON (bbg_resident_mode(1) or
// BBG_RESIDENT=1; the shim's process_queue replacement turns it on) the library keeps the device copy it already
// has, keyed by host address, and the next call on that array -- or on a slice of it -- skips the upload.
//
// Safety: the caller may rewrite its array between calls without telling us.  Every mirror therefore carries a
// fingerprint of the host content it was made from (72 stratified 8-byte words incl. the first and last elements, where
// the prover puts its blinding scalars); a lookup re-reads those words and treats any difference as a miss.  A
// mirror whose write-back was deferred (host_stale) is the newer copy by construction and is not checked.
#include <algorithm>
#include <cstring>

#include "api_common.hpp"

namespace bbg {

static constexpr size_t RESIDENT_MIN_BYTES = 4096;

bool resident_enabled(Context* ctx)
{
    if (ctx->resident_mode < 0) {
        const char* v = getenv("BBG_RESIDENT");
        ctx->resident_mode = (v && *v && atoi(v) != 0) ? 1 : 0;
    }
    return ctx->resident_mode == 1;
}

static void plan_samples(Context::Resident& e)
{
    // word offsets: the first element, the last four elements, and stratified words in between
    const size_t words = e.bytes / 8;
    uint32_t k = 0;
    auto add = [&](size_t w) {
        if (k < Context::Resident::SAMPLES && w < words) e.sample_off[k++] = w * 8;
    };
    for (size_t w = 0; w < 4; ++w) add(w);
    for (size_t w = 0; w < 16; ++w) add(words >= 16 ? words - 16 + w : w);
    const uint32_t strata = Context::Resident::SAMPLES - k;
    for (uint32_t s = 0; s < strata; ++s) {
        // a different word of the 4-word element in every stratum
        const size_t el = (words / 4) * (2 * (size_t)s + 1) / (2 * (size_t)strata);
        add(el * 4 + (s & 3));
    }
    e.n_samples = k;
}
static void take_samples(Context::Resident& e, size_t lo, size_t hi)
{
    for (uint32_t k = 0; k < e.n_samples; ++k) {
        const uint64_t off = e.sample_off[k];
        if (off >= lo && off + 8 <= hi) memcpy(&e.sample_val[k], e.host + off, 8);
    }
}
// true when every fingerprint word inside [lo, hi) still matches host memory and there are enough of them to tell
static bool samples_match(const Context::Resident& e, size_t lo, size_t hi)
{
    uint32_t seen = 0;
    for (uint32_t k = 0; k < e.n_samples; ++k) {
        const uint64_t off = e.sample_off[k];
        if (off >= lo && off + 8 <= hi) {
            uint64_t v;
            memcpy(&v, e.host + off, 8);
            if (v != e.sample_val[k]) return false;
            ++seen;
        }
    }
    return seen >= 8;
}

static size_t resident_budget(Context* ctx)
{
    if (ctx->resident_budget == 0) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) total_b = (size_t)16 << 30;
        ctx->resident_budget = total_b / 8;
        const char* v = getenv("BBG_RESIDENT_MAX_MB");
        if (v && *v) ctx->resident_budget = (size_t)atoll(v) << 20;
    }
    return ctx->resident_budget;
}

static int write_back(Context* ctx, Context::Resident& e, cudaStream_t st)
{
    int rc = g_staging.d2h((void*)e.host, e.d, e.bytes, st); // synchronises st
    if (rc) return rc;
    e.host_stale = false;
    take_samples(e, 0, e.bytes);
    return BBG_OK;
}

// A retired mirror's device block goes to a small free list instead of back to the driver: a prover allocates fresh
// host arrays for every proof (witness wires, z), so mirrors of the same few sizes are created and orphaned all the time,
// and cudaMalloc / cudaFree cost 0.1-1 ms apiece (they synchronise).  Blocks are only ever reused by work queued on the
// library's streams AFTER the work that last touched them, so no synchronisation is needed here.
static constexpr size_t RESIDENT_MAX_ENTRIES = 192;
static constexpr size_t RESIDENT_FREE_BLOCKS = 32;
static void drop(Context* ctx, size_t idx)
{
    Context::Resident& e = ctx->resident[idx];
    if (e.d) {
        if (ctx->resident_free.size() < RESIDENT_FREE_BLOCKS) {
            ctx->resident_free.push_back({ e.d, e.cap });
        } else {
            cudaFree(e.d);
            ctx->resident_bytes -= e.cap;
        }
    }
    ctx->resident.erase(ctx->resident.begin() + (long)idx);
}
static void* take_block(Context* ctx, size_t bytes, size_t* cap)
{
    auto& fl = ctx->resident_free;
    size_t best = fl.size();
    for (size_t i = 0; i < fl.size(); ++i) {
        if (fl[i].second >= bytes && fl[i].second <= 2 * bytes && (best == fl.size() || fl[i].second < fl[best].second)) best = i;
    }
    if (best != fl.size()) {
        void* p = fl[best].first;
        *cap = fl[best].second;
        fl.erase(fl.begin() + (long)best);
        return p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        // make room by returning the free list to the driver, then try once more
        for (auto& b : fl) {
            cudaFree(b.first);
            ctx->resident_bytes -= b.second;
        }
        fl.clear();
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    *cap = bytes;
    ctx->resident_bytes += bytes;
    return p;
}

int resident_acquire(Context* ctx, const void* host_v, size_t bytes, bool need_data, void** d_out, bool* hit, cudaStream_t st)
{
    *d_out = nullptr;
    if (hit) *hit = false;
    if (!resident_enabled(ctx) || bytes < RESIDENT_MIN_BYTES || host_v == nullptr) return BBG_OK;
    const char* host = (const char*)host_v;
    auto& tab = ctx->resident;
    const uint64_t now = ++ctx->resident_clock;
    int rc;
    for (size_t i = 0; i < tab.size(); ++i) {
        Context::Resident& e = tab[i];
        if (host >= e.host && host + bytes <= e.host + e.bytes) {
            const size_t lo = (size_t)(host - e.host);
            e.last_use = now;
            *d_out = (char*)e.d + lo;
            if (!need_data) return BBG_OK; // about to be overwritten: content irrelevant
            if (e.host_stale || samples_match(e, lo, lo + bytes)) {
                if (hit) *hit = true;
                ctx->resident_hits += 1;
                ctx->resident_h2d_saved += bytes;
                return BBG_OK;
            }
            // the caller rewrote (this part of) its array: refresh the mirror
            ctx->resident_misses += 1;
            if ((rc = g_staging.h2d(*d_out, host, bytes, st))) return rc;
            take_samples(e, lo, lo + bytes);
            return BBG_OK;
        }
    }
    // No mirror contains the range: mirrors that merely overlap it describe an array that no longer exists in that shape
    // (freed and reallocated) and are retired.  The library NEVER writes into host memory on its own initiative: a
    // pending write-back of such a mirror is abandoned -- its array is gone -- rather than scribbled over whatever lives
    // there now.  Host memory is written only by a call that names the array (an entry point with that host pointer, or
    // bbg_resident_flush), i.e. by a caller that vouches for it.
    for (size_t i = tab.size(); i-- > 0;) {
        Context::Resident& e = tab[i];
        if (host < e.host + e.bytes && e.host < host + bytes) drop(ctx, i);
    }
    // room: evict least-recently-used mirrors; mirrors that are ahead of host memory hold the only copy and stay
    const size_t budget = resident_budget(ctx);
    while (ctx->resident_bytes + bytes > budget || tab.size() >= RESIDENT_MAX_ENTRIES) {
        size_t victim = tab.size();
        for (size_t i = 0; i < tab.size(); ++i) {
            if (tab[i].host_stale) continue;
            if (victim == tab.size() || tab[i].last_use < tab[victim].last_use) victim = i;
        }
        if (victim == tab.size()) break; // nothing evictable: this array is simply not mirrored
        drop(ctx, victim);
        if (ctx->resident_bytes + bytes > budget && !ctx->resident_free.empty()) {
            // over the byte budget: the block really goes back to the driver
            for (auto& b : ctx->resident_free) {
                cudaFree(b.first);
                ctx->resident_bytes -= b.second;
            }
            ctx->resident_free.clear();
        }
    }
    if (ctx->resident_bytes + bytes > budget || tab.size() >= RESIDENT_MAX_ENTRIES) return BBG_OK;
    if (bytes > budget) return BBG_OK; // too large to mirror: caller stages as before
    Context::Resident e;
    e.host = host;
    e.bytes = bytes;
    e.d = take_block(ctx, bytes, &e.cap);
    if (e.d == nullptr) return BBG_OK; // out of memory: not resident
    e.last_use = now;
    plan_samples(e);
    if (need_data) {
        ctx->resident_misses += 1;
        if ((rc = g_staging.h2d(e.d, host, bytes, st))) {
            ctx->resident_free.push_back({ e.d, e.cap });
            return rc;
        }
        take_samples(e, 0, bytes);
    } else {
        // nothing valid yet: poison the fingerprint so that only a commit makes it match
        for (uint32_t k = 0; k < e.n_samples; ++k) e.sample_val[k] = 0x9e3779b97f4a7c15ull * (k + 1);
        e.host_stale = false;
    }
    *d_out = e.d;
    tab.push_back(e);
    return BBG_OK;
}

int resident_commit(Context* ctx, const void* host_v, size_t bytes, bool wb, cudaStream_t st)
{
    const char* host = (const char*)host_v;
    for (Context::Resident& e : ctx->resident) {
        if (host >= e.host && host + bytes <= e.host + e.bytes) {
            const size_t lo = (size_t)(host - e.host);
            if (wb) {
                int rc = g_staging.d2h((void*)host, (char*)e.d + lo, bytes, st); // synchronises st
                if (rc) return rc;
                take_samples(e, lo, lo + bytes);
                if (lo == 0 && bytes == e.bytes) e.host_stale = false;
            } else {
                e.host_stale = true;
            }
            return BBG_OK;
        }
    }
    set_last_error("resident_commit: no mirror for this array");
    return BBG_ERR_ARG;
}

// The mirrors keep their memory (the same arrays come back every proof); only their content is declared unknown:
// the fingerprint is poisoned so the next use uploads, and a deferred write-back is abandoned.
void resident_invalidate(Context* ctx, const void* host_v, size_t bytes)
{
    const char* host = (const char*)host_v;
    for (Context::Resident& e : ctx->resident) {
        if (bytes == 0 || (host < e.host + e.bytes && e.host < host + bytes)) {
            for (uint32_t k = 0; k < e.n_samples; ++k) e.sample_val[k] = 0x9e3779b97f4a7c15ull * (k + 1) + e.sample_off[k];
            e.host_stale = false;
        }
    }
}

int resident_flush(Context* ctx, const void* host_v, size_t bytes, cudaStream_t st)
{
    const char* host = (const char*)host_v;
    for (Context::Resident& e : ctx->resident) {
        if (!e.host_stale) continue;
        if (bytes == 0 || (host < e.host + e.bytes && e.host < host + bytes)) {
            int rc = write_back(ctx, e, st);
            if (rc) return rc;
        }
    }
    return BBG_OK;
}

// The caller vouches that host bytes [host, host + bytes) currently equal what `st` has just put into the mirror of that
// range (a device-to-device copy of data uploaded for another array): fingerprint the range from host memory so that
// later lookups find it valid.
void resident_adopt(Context* ctx, const void* host_v, size_t bytes)
{
    const char* host = (const char*)host_v;
    for (Context::Resident& e : ctx->resident) {
        if (host >= e.host && host + bytes <= e.host + e.bytes) {
            const size_t lo = (size_t)(host - e.host);
            take_samples(e, lo, lo + bytes);
            return;
        }
    }
}

bool resident_is_ahead(Context* ctx, const void* host_v, size_t bytes)
{
    const char* host = (const char*)host_v;
    for (const Context::Resident& e : ctx->resident) {
        if (host >= e.host && host + bytes <= e.host + e.bytes) return e.host_stale;
    }
    return false;
}

void resident_clear(Context* ctx)
{
    for (Context::Resident& e : ctx->resident) {
        if (e.d) cudaFree(e.d);
    }
    for (auto& b : ctx->resident_free) cudaFree(b.first);
    ctx->resident.clear();
    ctx->resident_free.clear();
    ctx->resident_bytes = 0;
}

} // namespace bbg

using namespace bbg;

extern "C" {

int bbg_resident_mode(int enable)
{
    GET_CTX();
    if (enable < 0) return resident_enabled(ctx) ? 1 : 0;
    if (!enable && resident_enabled(ctx)) {
        // mirrors are dropped, including ones whose write-back was deferred: flush first (bbg_resident_flush) if the host
        // copies are still wanted -- the library does not write into arrays nobody named
        cudaDeviceSynchronize();
        resident_clear(ctx);
    }
    ctx->resident_mode = enable ? 1 : 0;
    return BBG_OK;
}

int bbg_resident_invalidate(const void* host, size_t bytes)
{
    GET_CTX();
    resident_invalidate(ctx, host, bytes);
    return BBG_OK;
}

int bbg_resident_flush(const void* host, size_t bytes)
{
    GET_CTX();
    StreamScope order(ctx, ctx->stream);
    int rc = resident_flush(ctx, host, bytes, ctx->stream);
    if (rc) return rc;
    BBG_CUDA(cudaStreamSynchronize(ctx->stream));
    return BBG_OK;
}

int bbg_resident_stats(uint64_t* out4)
{
    GET_CTX();
    out4[0] = ctx->resident_hits;
    out4[1] = ctx->resident_misses;
    out4[2] = ctx->resident_h2d_saved;
    out4[3] = ctx->resident_bytes;
    return BBG_OK;
}

} // extern "C"
