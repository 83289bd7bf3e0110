// ctx.cuh -- process-wide device context shared by the MSM and NTT front-ends.
#pragma once
#include <cuda_runtime.h>

#include "../../include/bbg.h"

#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
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

struct NttTables; // ntt.cu

struct Context {
    int device = -1;
    int num_sms = 148;
    cudaStream_t stream = nullptr; // default work stream for the host-pointer entry points
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    // MSM workspaces
    DevBuf msm_scalars, msm_counts, msm_offsets, msm_cursors, msm_sorted, msm_buckets, msm_partials, msm_reduce,
        msm_scan_tmp, msm_result, msm_points;
    // NTT workspaces
    DevBuf ntt_data, ntt_scratch, ntt_pro, ntt_epi, ntt_small;
    std::map<unsigned, void*> ntt_twiddles; // log2n -> w_N^e table (N entries)
    void* ntt_stage_tw[2] = { nullptr, nullptr }; // per-direction small stage-twiddle tables
    // pinned staging for small results
    void* pinned = nullptr;
    size_t pinned_cap = 0;
    uint64_t launches = 0; // kernels launched by this library (bench.py reports it)
    double last_kernel_ms = 0.0;
    std::mutex mu;
};

int get_context(Context** out);

inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

} // namespace bbg
