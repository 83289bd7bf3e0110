// api_common.hpp -- plumbing shared by the extern "C" translation units (api.cu, poly_api.cu): context access with
// device binding, staging copies, BBG_STATS accounting, the Pippenger object.  Not part of the C-ABI.
#pragma once
#include <chrono>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/bbg.h"
#include "ctx.cuh"
#include "g1.cuh"
#include "internal.hpp"
#include "staging.hpp"

namespace bbg {

// binds the context's device for the duration of an entry point and restores the caller's current device afterwards
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device);
    ~DeviceGuard();
};

#define GET_CTX()                                \
    ::bbg::Context* ctx = nullptr;               \
    {                                            \
        int _rc = ::bbg::get_context(&ctx);      \
        if (_rc) return _rc;                     \
    }                                            \
    ::bbg::DeviceGuard _dg(ctx->device);         \
    std::lock_guard<std::mutex> _lk(ctx->mu)
#define GET_CTX_PTR()                            \
    ::bbg::Context* ctx = nullptr;               \
    if (::bbg::get_context(&ctx)) return nullptr; \
    ::bbg::DeviceGuard _dg(ctx->device);         \
    std::lock_guard<std::mutex> _lk(ctx->mu)

int ntt_run_kind(Context* ctx, void* d_coeffs, size_t n, int kind, size_t generator_size, const void* constant, cudaStream_t st);

extern Staging g_staging; // pinned staging buffers + copy threads for pageable host memory (staging.hpp)


// RAII device timer around the kernels of one host-pointer call: construct after the H2D copies are queued, stop()
// before the D2H copies are queued, finish() after them (synchronises the stream).
struct DeviceTimer {
    Context* c;
    bool stopped = false;
    explicit DeviceTimer(Context* ctx) : c(ctx) { cudaEventRecord(c->ev_a, c->stream); }
    void stop()
    {
        cudaEventRecord(c->ev_b, c->stream);
        stopped = true;
    }
    int finish()
    {
        if (!stopped) stop();
        BBG_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        BBG_CUDA(cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
        c->last_kernel_ms = ms;
        return BBG_OK;
    }
};

// ---- BBG_STATS=1: wall time / device time / PCIe bytes spent inside the host-pointer entry points, printed to stderr at
// exit.  Lets a drop-in user see how much of (say) a proof is the hot path and how much of THAT is pageable copies.
struct HostStats {
    struct Row {
        const char* name;
        uint64_t calls = 0, bytes_h2d = 0, bytes_d2h = 0;
        double wall_s = 0, device_ms = 0;
    };
    Row rows[4] = { { "msm" }, { "ntt" }, { "srs" }, { "poly" } };
    bool enabled = false;
    bool per_call = false; // BBG_STATS=2: one stderr line per call as well
    HostStats()
    {
        const char* v = getenv("BBG_STATS");
        enabled = v && *v && atoi(v) != 0;
        per_call = enabled && atoi(v) >= 2;
    }
    ~HostStats()
    {
        if (!enabled) return;
        for (const Row& r : rows) {
            if (r.calls == 0) continue;
            fprintf(stderr, "{\"bbg_stats\": \"%s\", \"calls\": %llu, \"wall_s\": %.6f, \"device_kernel_s\": %.6f, \"h2d_bytes\": %llu, \"d2h_bytes\": %llu}\n",
                    r.name, (unsigned long long)r.calls, r.wall_s, r.device_ms * 1e-3, (unsigned long long)r.bytes_h2d,
                    (unsigned long long)r.bytes_d2h);
        }
    }
};
extern HostStats g_stats;
struct StatScope {
    HostStats::Row* row;
    Context* ctx;
    std::chrono::steady_clock::time_point t0;
    uint64_t h2d, d2h, launches0 = 0;
    StatScope(int which, Context* c, uint64_t h2d_, uint64_t d2h_)
        : row(&g_stats.rows[which]), ctx(c), h2d(h2d_), d2h(d2h_) // always accounted (a few adds); printing is what BBG_STATS gates
    {
        if (!row) return;
        launches0 = ctx->launches;
        t0 = std::chrono::steady_clock::now();
        row->calls += 1;
        row->bytes_h2d += h2d;
        row->bytes_d2h += d2h;
        ctx->last_kernel_ms = 0.0;
    }
    ~StatScope()
    {
        if (!row) return;
        const double w = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        row->wall_s += w;
        row->device_ms += ctx->last_kernel_ms;
        if (g_stats.per_call) {
            fprintf(stderr, "{\"bbg_call\": \"%s\", \"h2d_bytes\": %llu, \"d2h_bytes\": %llu, \"wall_ms\": %.3f, \"device_kernel_ms\": %.3f, \"kernels\": %llu}\n",
                    row->name, (unsigned long long)h2d, (unsigned long long)d2h, w * 1e3, ctx->last_kernel_ms,
                    (unsigned long long)(ctx->launches - launches0));
        }
    }
};
enum { STAT_MSM = 0, STAT_NTT = 1, STAT_SRS = 2, STAT_POLY = 3 };

// ---- Pippenger object: the SRS resident in HBM as n contiguous affine points
struct PippengerObj {
    affine_t* d_points = nullptr;     // level 0 = the n SRS points; levels 1..L-1 follow (msm.cu k_msm_precompute)
    size_t n = 0;
    MsmLevels lv;
    const void* host_table = nullptr; // adopted 2n host table (for pointer recognition), may be null
};
extern std::vector<PippengerObj*> g_pippengers;

} // namespace bbg
