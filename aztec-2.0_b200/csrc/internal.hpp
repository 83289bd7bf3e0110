// internal.hpp -- declarations shared by api.cu, msm.cu and ntt.cu (not part of the C-ABI).
#pragma once
#include "ctx.cuh"
#include "host_field.hpp"

namespace bbg {

// Orders a call that uses the context's shared workspaces / cached tables on stream `st` after the previous call if that
// one ran on a different stream (see Context::last_use).  Construct before queueing any work, destroy after the last launch.
struct StreamScope {
    Context* c;
    cudaStream_t st;
    StreamScope(Context* ctx, cudaStream_t s) : c(ctx), st(s)
    {
        if (c->last_valid && c->last_stream != st) cudaStreamWaitEvent(st, c->last_use, 0);
    }
    ~StreamScope()
    {
        cudaEventRecord(c->last_use, st);
        c->last_stream = st;
        c->last_valid = true;
    }
};

// msm.cu
// Fixed-base levels of a Pippenger object: entry l * stride + i holds 2^(D l) * P_i.  L == 1: plain points.
struct MsmLevels {
    unsigned L = 1;    // levels held in HBM
    unsigned D = 0;    // bits between consecutive levels (a multiple of c)
    unsigned c = 0;    // window bits the levels were built for
    size_t stride = 0; // table entries per level
};
MsmLevels msm_levels_plan(size_t n, size_t max_table_bytes);
int msm_precompute_device(Context* ctx, void* d_table, size_t n, const MsmLevels& lv, cudaStream_t st);
// Scalars still in flight from the host: `count` equal pieces of `piece` scalars each (the last may be short); piece k is
// valid on the device once ready[k] has fired.  msm_device then runs the histogram pass piece by piece behind the copies.
struct MsmArrival {
    const cudaEvent_t* ready = nullptr;
    size_t count = 0;
    size_t piece = 0;
};
// fused batch: up to four scalar vectors (device pointers, n elements each) over the same bases and range in ONE pass of the
// MSM kernels; d_out then receives `count` Jacobian results
struct MsmBatch {
    const void* scalars[4];
    unsigned count;
};
int msm_device(Context* ctx, MsmWorkspace& ws, const void* d_scalars, size_t n, const void* d_points, size_t point_stride,
               const MsmLevels& lv, size_t base, void* d_out, cudaStream_t st, const MsmArrival* arrival = nullptr,
               bool allow_parts = true, const struct MsmBatch* batch = nullptr);
int g1_sum_device(Context* ctx, const void* d_jacs, size_t n, void* d_out, cudaStream_t st);
int srs_decode_device(Context* ctx, const void* d_raw, size_t n, void* d_points, cudaStream_t st);
int point_table_device(Context* ctx, const void* d_points, size_t n, void* d_table, cudaStream_t st);
int compact_even_device(Context* ctx, const void* d_table, size_t n, void* d_points, cudaStream_t st);

// resident.cu -- device mirrors of host arrays (see Context::Resident)
bool resident_enabled(Context* ctx);
// Device address mirroring host bytes [host, host + bytes).  need_data: the device copy must hold the host content on return
// (queued on `st`): served from a valid mirror, else uploaded.  *hit tells which.  With residency off this returns
// BBG_OK and *d_out = nullptr: the caller uses its own staging buffer.
int resident_acquire(Context* ctx, const void* host, size_t bytes, bool need_data, void** d_out, bool* hit, cudaStream_t st);
// The device mirror of [host, host + bytes) was just (re)written on `st`.  write_back: copy it to the host array now
// (synchronises `st`) and re-fingerprint; otherwise the host copy is marked stale until resident_flush.
int resident_commit(Context* ctx, const void* host, size_t bytes, bool write_back, cudaStream_t st);
// The host array [host, host + bytes) was written by the caller (or will be): drop / refresh mirrors overlapping it.
void resident_invalidate(Context* ctx, const void* host, size_t bytes);
// Write every stale mirror overlapping [host, host + bytes) back to host memory (bytes == 0: all of them).
int resident_flush(Context* ctx, const void* host, size_t bytes, cudaStream_t st);
// true when a mirror containing the range holds data newer than host memory (a deferred write-back is pending)
bool resident_is_ahead(Context* ctx, const void* host, size_t bytes);
void resident_adopt(Context* ctx, const void* host, size_t bytes);
void resident_clear(Context* ctx);

// ntt.cu
struct NttScale {
    bool present = false;
    hf::Fr start;       // constant factor
    bool has_shift = false;
    hf::Fr shift;       // geometric factor shift^i
    uint64_t size = 0;  // prologue only: applies to i < size
};
struct NttDist {
    unsigned rank_bits = 0; // log2(number of ranks); 0 = single GPU
    unsigned rank = 0;
    int phase = 0;          // 0: every pass but the last (before the all-to-all); 1: the last pass (after it)
    void* const* peer_recv = nullptr; // phase 0, fused exchange: receive buffer of every rank (peer-mapped); null = local store + NCCL
    // natural block distribution over peer memory: phase 0's first pass loads from the owners' input blocks, phase 1's pass
    // stores to the owners' output blocks (each block n / world elements); null = the packed sliced layouts
    void* const* peer_src = nullptr;
    void* const* peer_out = nullptr;
};
int ntt_device(Context* ctx, const void* d_src, void* d_dst, unsigned log_n, bool inverse, const NttScale& pro, const NttScale& epi,
               unsigned out_shift, unsigned out_off, cudaStream_t st, const NttDist& dist = NttDist());
hf::Fr ntt_root_of_unity(unsigned log_n);

} // namespace bbg
