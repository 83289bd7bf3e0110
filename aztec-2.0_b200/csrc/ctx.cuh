// ctx.cuh -- process-wide device context shared by the MSM and NTT front-ends.
#pragma once
#include <cuda_runtime.h>

#include "../../include/bbg.h"

#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

namespace bbg {

void set_last_error(const std::string& s);

#define BBG_CUDA(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            ::bbg::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " __FILE__ + \
                                  ":" + std::to_string(__LINE__));                                      \
            return BBG_ERR_CUDA;                                                                 \
        }                                                                                               \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) {
            return BBG_OK;
        }
        if (p) {
            BBG_CUDA(cudaFree(p));
            p = nullptr;
            cap = 0;
        }
        size_t want = bytes + (bytes >> 3);
        BBG_CUDA(cudaMalloc(&p, want));
        cap = want;
        return BBG_OK;
    }
    void release()
    {
        if (p) {
            cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
};

// Optional per-phase device timing (bbg_profile): CUDA events recorded on the launching stream at
// phase boundaries; bench.py reads the per-phase milliseconds for the roofline of the dominant kernel.
enum Phase {
    PH_MSM_DIGITS = 0, PH_MSM_SCAN, PH_MSM_SCATTER, PH_MSM_PAIRS, PH_MSM_ACCUMULATE, PH_MSM_FIXUP, PH_MSM_REDUCE, PH_MSM_COMBINE,
    PH_NTT_TABLES, PH_NTT_PASS0, PH_NTT_PASS1, PH_NTT_PASS2, PH_NTT_PASS3, PH_COUNT
};
struct Profiler {
    static constexpr int MAX_MARKS = 64;
    bool on = false;
    int n = 0;
    int ids[MAX_MARKS];
    cudaEvent_t ev[MAX_MARKS] = {};
    void begin() { n = 0; }
    // marks the START of phase `id` (id < 0: end of the last phase)
    void mark(cudaStream_t st, int id)
    {
        if (!on || n >= MAX_MARKS) return;
        if (ev[n] == nullptr) cudaEventCreate(&ev[n]);
        cudaEventRecord(ev[n], st);
        ids[n++] = id;
    }
};

// Everything one MSM in flight needs.  The context owns four, so that a batch (bbg_pippenger*_batch) can run MSM i + 1 ..
// i + 3 on other streams while the latency-bound tail of MSM i (slot merge, bucket reduction) is still draining.
struct MsmWorkspace {
    DevBuf scalars, counts, offsets, cursors, sorted, buckets, partials, reduce, scan_tmp, result, lvl_offsets, pairs_a, pairs_b,
        pair_pre, pair_meta, pts0;
    void release()
    {
        DevBuf* all[] = { &scalars, &counts, &offsets, &cursors, &sorted, &buckets, &partials, &reduce, &scan_tmp, &result,
                          &lvl_offsets, &pairs_a, &pairs_b, &pair_pre, &pair_meta, &pts0 };
        for (DevBuf* b : all) b->release();
    }
};

struct Context {
    int device = -1;
    int num_sms = 148;
    cudaStream_t stream = nullptr; // default work stream for the host-pointer entry points
    cudaStream_t copy_stream = nullptr; // H2D pieces that overlap with kernels on `stream`
    static constexpr int BATCH_WAYS = 4;
    cudaStream_t aux_stream[BATCH_WAYS - 1] = {}; // extra work streams of the batched entry points
    cudaEvent_t ev_piece[8] = {};
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join[BATCH_WAYS - 1] = {};
    // a single large MSM is cut into bucket ranges ("parts") whose accumulate -> merge -> reduce chains run on these
    // high-priority streams, so that the latency-bound tail of one part runs under the accumulation of the next (msm.cu)
    static constexpr int MSM_MAX_PARTS = 4;
    cudaStream_t part_stream[MSM_MAX_PARTS - 1] = {};
    cudaEvent_t ev_part_fork = nullptr, ev_part_join[MSM_MAX_PARTS - 1] = {};
    // Cross-stream ordering of the shared workspaces and cached tables: every call records `last_use` on its stream when
    // it has queued its work, and a call on a DIFFERENT stream first waits for it (StreamScope in internal.hpp).  Calls on
    // different streams therefore serialise on the device; they never race on the workspaces.
    cudaEvent_t last_use = nullptr;
    cudaStream_t last_stream = nullptr;
    bool last_valid = false;
    // MSM workspaces
    MsmWorkspace msm_ws[BATCH_WAYS];
    DevBuf msm_points;
    void* inv_fix_fq = nullptr; // inv.cuh fix-up constants (fq)
    // NTT workspaces
    DevBuf ntt_data, ntt_scratch, ntt_pro, ntt_epi, ntt_small;
    std::map<unsigned, void*> ntt_twiddles; // log2n -> w_N^e table (N entries)
    // full geometric tables T[i] = start * shift^i cached across calls (coset pre/post scalings, 1/n-scaled twiddles):
    // a key is promoted to a table the second time it is seen, so one-off constants never pay for a build
    struct ScaleTab {
        uint64_t count = 0;
        uint64_t start[4] = {}, shift[4] = {};
        void* tab = nullptr; // nullptr: key seen once, not built yet
        uint64_t last_use = 0;
    };
    std::vector<ScaleTab> ntt_scale_cache;
    uint64_t ntt_scale_clock = 0;
    void* ntt_stage_tw[2] = { nullptr, nullptr }; // per-direction small stage-twiddle tables
    bool ntt_attr_set = false;                    // cudaFuncSetAttribute done for this context's device
    size_t ntt_table_bytes = 0;                   // bytes held by ntt_twiddles + ntt_scale_cache (budget: ntt.cu)
    std::map<unsigned, uint64_t> ntt_twiddle_use; // log2n -> last-use clock (LRU eviction together with the scale tables)
    // small pinned staging buffer for results of batched calls
    void* pinned = nullptr;
    size_t pinned_cap = 0;
    // pointwise / scan workspaces (poly.cu)
    DevBuf poly_tmp, poly_stage, poly_out;
    // Resident polynomials (resident.cu): device mirrors of caller-owned host arrays, keyed by host address, so that a
    // chain of calls on the same array (ifft -> commitment MSM -> coset FFT) crosses PCIe once.  Off unless enabled.
    struct Resident {
        const char* host = nullptr;
        size_t bytes = 0;    // extent of the host array mirrored
        void* d = nullptr;
        size_t cap = 0;
        bool host_stale = false; // the device copy is newer than host memory (write-back deferred)
        uint64_t last_use = 0;
        static constexpr int SAMPLES = 72;
        uint32_t n_samples = 0;
        uint64_t sample_off[SAMPLES]; // byte offsets of the fingerprint words
        uint64_t sample_val[SAMPLES];
    };
    std::vector<Resident> resident;
    std::vector<std::pair<void*, size_t>> resident_free; // retired device blocks (pointer, capacity) awaiting reuse
    int resident_mode = -1; // -1: read BBG_RESIDENT on first use; 0 off; 1 on
    size_t resident_bytes = 0, resident_budget = 0;
    uint64_t resident_clock = 0;
    uint64_t resident_hits = 0, resident_misses = 0, resident_h2d_saved = 0;
    Profiler prof;
    uint64_t launches = 0; // kernels launched by this library (bench.py reports it)
    double last_kernel_ms = 0.0;
    std::mutex mu;
};

int get_context(Context** out);

inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

} // namespace bbg
