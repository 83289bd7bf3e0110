// internal.hpp -- declarations shared by api.cu, msm.cu and ntt.cu (not part of the C-ABI).
#pragma once
#include "ctx.cuh"
#include "host_field.hpp"

namespace bbg {

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
int msm_device(Context* ctx, const void* d_scalars, size_t n, const void* d_points, size_t point_stride, const MsmLevels& lv,
               size_t base, void* d_out, cudaStream_t st, const MsmArrival* arrival = nullptr);
int g1_sum_device(Context* ctx, const void* d_jacs, size_t n, void* d_out, cudaStream_t st);
int srs_decode_device(Context* ctx, const void* d_raw, size_t n, void* d_points, cudaStream_t st);
int point_table_device(Context* ctx, const void* d_points, size_t n, void* d_table, cudaStream_t st);
int compact_even_device(Context* ctx, const void* d_table, size_t n, void* d_points, cudaStream_t st);

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
};
int ntt_device(Context* ctx, const void* d_src, void* d_dst, unsigned log_n, bool inverse, const NttScale& pro, const NttScale& epi,
               unsigned out_shift, unsigned out_off, cudaStream_t st, const NttDist& dist = NttDist());
hf::Fr ntt_root_of_unity(unsigned log_n);

} // namespace bbg
