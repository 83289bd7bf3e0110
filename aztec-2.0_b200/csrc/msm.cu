// msm.cu -- BN254 G1 multi-scalar multiplication (Pippenger bucket method) for sm_100a.
//
// Replaces barretenberg's scalar_multiplication::pippenger / pippenger_unsafe
// (bb/ecc/curves/bn254/scalar_multiplication/scalar_multiplication.cpp:853-929).  Same inputs
// (Montgomery fr scalars, Montgomery affine points), same group element out (96-byte Jacobian).
// The CPU reference's pipeline (compute_wnaf_states :188-252 -> organize_buckets :260-271 ->
// reduce_buckets/add_affine_points :305-521 -> evaluate_pippenger_rounds :720-838) is re-thought
// for the GPU rather than translated:
//
//   0. k_msm_precompute      once per SRS (Pippenger object): level l holds 2^(D*l) * P_i in affine form, so
//                            window w = l*S + r of every scalar can use the SAME S bucket sets (S = 1 when
//                            all W levels fit in HBM): the per-window bucket reductions and the c*W
//                            Horner doublings of the textbook method disappear.  [The reference makes the
//                            same kind of trade at SRS-load time: it doubles the table with the
//                            endomorphism images, scalar_multiplication.cpp:104-112.]
//   1. k_msm_digits<false>   scalar -> canonical -> signed c-bit windows, histogram of bucket sizes
//   2. scan                  bucket sizes -> bucket start offsets (flat key space: set*B + |digit|-1)
//   3. k_msm_digits<true>    counting-sort scatter of (level-table index, sign) by bucket   [order inside a
//                            bucket is irrelevant: the group is commutative]
//   4. k_msm_accumulate      the sorted list is cut into equal-length chunks, one per thread, so
//                            every thread does the same number of mixed additions no matter how
//                            skewed the scalar distribution is; a bucket cut by a chunk boundary leaves
//                            partial sums ("slots") that
//      k_msm_merge           folds, 8 slots per thread, level after level (log depth even when every
//                            scalar is equal and one bucket holds everything)
//   5. k_msm_segments        sum_b (b+1)*bucket[b]: running sums over segments of `ell` buckets (the
//                            reference's running-sum trick :773-783), each segment then shifted to its
//                            place by a short double-and-add, and
//      k_msm_tree_sum        a warp-shuffle / shared-memory tree of g1 additions per bucket set
//   6. k_msm_finish          Horner over the S bucket sets (S = 1: nothing), XYZZ -> Jacobian
//
// Arithmetic is integer-pipe bound (IMAD.WIDE at half rate: 70 G fq-mul/s measured on B200); HBM
// traffic is ~(32 + 68*W) B per point.  No kernel calls a non-inlined device function (see g1.cuh).
#include <algorithm>

#include "g1.cuh"
#include "internal.hpp"

namespace bbg {

// ------------------------------------------------------------------------------------------------
// 1/3. digits: histogram (SCATTER = false) or counting-sort scatter (SCATTER = true)
// ------------------------------------------------------------------------------------------------
struct DigitParams {
    uint32_t n;
    uint32_t c;            // window bits
    uint32_t W;            // windows
    uint32_t S;            // bucket sets (windows per level)
    uint32_t B;            // buckets per set = 2^(c-1)
    uint32_t level_stride; // table entries between consecutive levels
    uint32_t base;         // table entry of scalar 0 (Pippenger `from`)
};

template <bool SCATTER>
__global__ void __launch_bounds__(256) k_msm_digits(const fr_t* __restrict__ scalars,
                                                    const DigitParams P,
                                                    uint32_t* __restrict__ counters,
                                                    uint32_t* __restrict__ sorted)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) {
        return;
    }
    // from_montgomery_form => canonical integer in [0, r)  (scalar_multiplication.cpp:224)
    fr_t s = fe_from_mont(fe_load_nc<FrParams>(scalars + i));
    uint32_t k[9];
#pragma unroll
    for (int j = 0; j < 8; ++j) k[j] = s.l[j];
    k[8] = 0;
    const uint32_t c = P.c;
    const uint32_t half = 1u << (c - 1);
    const uint32_t mask = (1u << c) - 1;
    uint32_t carry = 0;
    uint32_t level = 0, set = 0;
    for (unsigned w = 0; w < P.W; ++w) {
        unsigned pos = w * c;
        unsigned limb = pos >> 5, off = pos & 31;
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            lo = limb == (unsigned)j ? k[j] : lo;
            hi = limb == (unsigned)j ? k[j + 1] : hi;
        }
        uint64_t two = (uint64_t)lo | ((uint64_t)hi << 32);
        uint32_t v = ((uint32_t)(two >> off) & mask) + carry;
        // signed digit in (-2^(c-1), 2^(c-1)]: v > half  =>  digit v - 2^c, borrow one from the next window
        uint32_t neg = v > half ? 1u : 0u;
        uint32_t mag = neg ? (mask + 1 - v) : v;
        carry = neg;
        if (mag) {
            uint32_t g = set * P.B + (mag - 1);
            if (SCATTER) {
                uint32_t dst = atomicAdd(&counters[g], 1u);
                sorted[dst] = ((level * P.level_stride + P.base + i) << 1) | neg;
            } else {
                atomicAdd(&counters[g], 1u);
            }
        }
        if (++set == P.S) {
            set = 0;
            ++level;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 2. exclusive scan of G bucket sizes -> offsets[0..G] (offsets[G] = total), three small kernels
// ------------------------------------------------------------------------------------------------
static constexpr int SCAN_THREADS = 1024;
static constexpr int SCAN_ITEMS = 4;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem, uint32_t& total)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = smem[lane];
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= (unsigned)d) winc += t;
        }
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    total = smem[32];
    uint32_t r = smem[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ tile_sums)
{
    __shared__ uint32_t smem[33];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < n) s += in[base + j];
    }
    uint32_t total;
    block_exclusive_scan(s, smem, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_top(uint32_t* __restrict__ tile_sums, unsigned num_tiles)
{
    __shared__ uint32_t smem[33];
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        unsigned idx = threadIdx.x * SCAN_ITEMS + j;
        v[j] = idx < num_tiles ? tile_sums[idx] : 0;
        s += v[j];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, smem, total);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        unsigned idx = threadIdx.x * SCAN_ITEMS + j;
        if (idx < num_tiles) tile_sums[idx] = ex;
        ex += v[j];
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* __restrict__ in,
                                                             size_t n,
                                                             const uint32_t* __restrict__ tile_prefix,
                                                             uint32_t* __restrict__ out,
                                                             uint32_t* __restrict__ out_copy)
{
    __shared__ uint32_t smem[33];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        v[j] = (base + j < n) ? in[base + j] : 0;
        s += v[j];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, smem, total) + tile_prefix[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < n) {
            out[base + j] = ex;
            out_copy[base + j] = ex;
        }
        ex += v[j];
        if (base + j + 1 == n) {
            out[n] = ex; // grand total
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 4. bucket accumulation over equal-length chunks of the bucket-sorted list
// ------------------------------------------------------------------------------------------------
// A "slot" is the partial sum of a bucket whose entries straddle a chunk boundary.  Every worker
// emits exactly two: slot[2t] (the run that continues from the previous worker) and slot[2t+1] (the run that
// continues into the next).  Consecutive slots of one bucket are therefore adjacent, and a worker
// whose whole range lies inside one bucket emits (sum, identity) so the chain is never broken.
struct alignas(16) Slot {
    xyzz_t acc;
    uint32_t bucket; // SLOT_NONE: empty
    uint32_t pad[3];
};
static constexpr uint32_t SLOT_NONE = 0xffffffffu;
static constexpr int ACC_THREADS = 128;
static constexpr uint32_t MERGE_K = 8;

__device__ __forceinline__ void slot_store(Slot* s, const xyzz_t& v, uint32_t bucket)
{
    xyzz_store(&s->acc, v);
    s->bucket = bucket;
}

__device__ __forceinline__ uint32_t upper_bound_u32(const uint32_t* __restrict__ a, uint32_t n, uint32_t key)
{
    // first index with a[idx] > key, over a[0..n)
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= key) {
            lo = mid + 1;
        } else {
            hi = mid;
        }
    }
    return lo;
}

__global__ void __launch_bounds__(ACC_THREADS, 4) k_msm_accumulate(const uint32_t* __restrict__ sorted,
                                                                   const uint32_t* __restrict__ offsets, // G+1
                                                                   uint32_t G,
                                                                   uint32_t chunk,
                                                                   uint32_t num_chunks,
                                                                   const affine_t* __restrict__ points,
                                                                   uint32_t point_stride,
                                                                   xyzz_t* __restrict__ buckets,
                                                                   Slot* __restrict__ slots,
                                                                   uint32_t* __restrict__ pending)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_chunks) {
        return;
    }
    Slot* slot_a = slots + 2 * (size_t)t;
    Slot* slot_b = slot_a + 1;
    const uint32_t total = __ldg(offsets + G); // number of non-zero digits, produced by the scan
    const uint64_t start64 = (uint64_t)t * chunk;
    if (start64 >= total) {
        slot_a->bucket = SLOT_NONE;
        slot_b->bucket = SLOT_NONE;
        return;
    }
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (start64 + chunk < total) ? start + chunk : total;

    // the (non-empty) bucket that contains position `start`
    uint32_t g = upper_bound_u32(offsets, G + 1, start) - 1;
    uint32_t g_begin = __ldg(offsets + g);
    uint32_t g_end = __ldg(offsets + g + 1);

    bool a_set = false, b_set = false;
    xyzz_t acc = xyzz_infinity();
    uint32_t pos = start;
    uint32_t v = __ldg(sorted + pos);
    affine_t pt = affine_load(points + (size_t)(v >> 1) * point_stride);
    while (true) {
        // prefetch the next entry's point while this one is being added
        uint32_t vn = 0;
        affine_t ptn;
        const bool more = pos + 1 < end;
        if (more) {
            vn = __ldg(sorted + pos + 1);
            ptn = affine_load(points + (size_t)(vn >> 1) * point_stride);
        }
        if (!affine_is_inf(pt)) {
            if (v & 1) {
                pt.y = fe_neg(pt.y);
            }
            xyzz_madd(acc, pt);
        }
        ++pos;
        if (pos == g_end || pos == end) {
            const bool cont_l = g_begin < start; // bucket began in an earlier chunk
            const bool cont_r = g_end > end;     // bucket continues into a later chunk
            if (!cont_l && !cont_r) {
                xyzz_store(buckets + g, acc);
            } else {
                if (cont_l) {
                    slot_store(slot_a, acc, g);
                    a_set = true;
                }
                if (cont_r) {
                    slot_store(slot_b, cont_l ? xyzz_infinity() : acc, g);
                    b_set = true;
                }
            }
            if (pos == end) {
                break;
            }
            acc = xyzz_infinity();
            // next non-empty bucket
            do {
                ++g;
                g_begin = g_end;
                g_end = __ldg(offsets + g + 1);
            } while (g_end == g_begin);
        }
        v = vn;
        pt = ptn;
    }
    if (!a_set) slot_a->bucket = SLOT_NONE;
    if (!b_set) slot_b->bucket = SLOT_NONE;
    if (a_set || b_set) {
        atomicAdd(pending, 1u); // something for the merge levels to do
    }
}

// One merge level: worker u folds slots [u*K + 1, (u+1)*K + 1) (worker 0 also takes slot 0), i.e. the
// boundary falls between the two slots of one chunk so the typical (tail, next head) pair is never cut.
__global__ void __launch_bounds__(128) k_msm_merge(const Slot* __restrict__ in,
                                                    uint32_t n_in,
                                                    uint32_t workers,
                                                    const uint32_t* __restrict__ pending_in,
                                                    xyzz_t* __restrict__ buckets,
                                                    Slot* __restrict__ out,
                                                    uint32_t* __restrict__ pending_out)
{
    if (__ldg(pending_in) == 0) {
        return; // the previous level left nothing open (the usual case after the first merge)
    }
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= workers) {
        return;
    }
    const uint32_t lo = u == 0 ? 0 : u * MERGE_K + 1;
    uint32_t hi = (u + 1) * MERGE_K + 1;
    if (hi > n_in || u + 1 == workers) hi = n_in;
    Slot* out_a = out + 2 * (size_t)u;
    Slot* out_b = out_a + 1;
    bool a_set = false, b_set = false;

    xyzz_t acc = xyzz_infinity();
    uint32_t run_bucket = SLOT_NONE;
    uint32_t run_start = lo;
    for (uint32_t i = lo; i <= hi; ++i) {
        const uint32_t b = i < hi ? in[i].bucket : SLOT_NONE;
        const bool close = run_bucket != SLOT_NONE && (i == hi || b != run_bucket);
        if (close) {
            const bool cont_l = run_start == lo && lo > 0 && in[lo - 1].bucket == run_bucket;
            const bool cont_r = i == hi && hi < n_in && in[hi].bucket == run_bucket;
            if (!cont_l && !cont_r) {
                xyzz_store(buckets + run_bucket, acc);
            } else {
                if (cont_l) {
                    slot_store(out_a, acc, run_bucket);
                    a_set = true;
                }
                if (cont_r) {
                    slot_store(out_b, cont_l ? xyzz_infinity() : acc, run_bucket);
                    b_set = true;
                }
            }
            run_bucket = SLOT_NONE;
        }
        if (i < hi && b != SLOT_NONE) {
            xyzz_t x = xyzz_load(&in[i].acc);
            if (run_bucket == SLOT_NONE) {
                run_bucket = b;
                run_start = i;
                acc = x;
            } else {
                xyzz_add(acc, x); // the one inlined add site of this kernel
            }
        }
    }
    if (!a_set) out_a->bucket = SLOT_NONE;
    if (!b_set) out_b->bucket = SLOT_NONE;
    if (a_set || b_set) {
        atomicAdd(pending_out, 1u);
    }
}

// ------------------------------------------------------------------------------------------------
// 5. bucket reduction: sum over b of (b + 1) * bucket[b] per set
// ------------------------------------------------------------------------------------------------
// Worker (set, seg): running sums over buckets [seg*ell, (seg+1)*ell):  R = sum x_j,  A = sum (j+1) x_j,
// then V = A + (seg*ell) * R by double-and-add.  One add site: the loop alternates run += x / acc += run.
__global__ void __launch_bounds__(128) k_msm_segments(const xyzz_t* __restrict__ buckets,
                                                       uint32_t B,   // buckets per set
                                                       uint32_t ell, // segment length (divides B)
                                                       uint32_t num_workers,
                                                       xyzz_t* __restrict__ out)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_workers) {
        return;
    }
    const uint32_t segs = B / ell;
    const uint32_t set = t / segs, seg = t % segs;
    const xyzz_t* src = buckets + (size_t)set * B + (size_t)seg * ell;
    xyzz_t run = xyzz_infinity(), acc = xyzz_infinity();
    for (uint32_t step = 0; step < 2 * ell; ++step) {
        const bool second = step & 1;
        xyzz_t rhs;
        if (!second) {
            rhs = xyzz_load(src + (ell - 1 - (step >> 1)));
        } else {
            rhs = run;
        }
        xyzz_t lhs = xyzz_select(second, acc, run);
        xyzz_add(lhs, rhs);
        if (second) {
            acc = lhs;
        } else {
            run = lhs;
        }
    }
    // V = acc + (seg * ell) * run
    const uint32_t k = seg * ell;
    if (k != 0 && !xyzz_is_inf(run)) {
        const int msb = 31 - __clz(k);
        xyzz_t m = run;
        for (int bit = msb - 1; bit >= -1; --bit) {
            xyzz_t rhs;
            bool do_add;
            if (bit >= 0) {
                m = xyzz_dbl(m);
                rhs = run;
                do_add = (k >> bit) & 1;
            } else {
                rhs = acc; // last step: fold in A
                do_add = true;
            }
            if (do_add) {
                xyzz_add(m, rhs);
            }
        }
        acc = m;
    }
    xyzz_store(out + t, acc);
}

// Sum of m consecutive XYZZ points per row: grid (parts, rows); each CTA strides over its share and
// finishes with a warp-shuffle tree + one shared-memory hop.  out[row * parts + part].
static constexpr int TREE_THREADS = 256;
__global__ void __launch_bounds__(TREE_THREADS) k_msm_tree_sum(const xyzz_t* __restrict__ in, uint32_t m, xyzz_t* __restrict__ out)
{
    __shared__ xyzz_t sm[TREE_THREADS / 32];
    const uint32_t parts = gridDim.x, part = blockIdx.x, row = blockIdx.y;
    const xyzz_t* src = in + (size_t)row * m;
    const uint32_t per = (m + parts - 1) / parts;
    const uint32_t lo = part * per;
    const uint32_t hi = min(lo + per, m);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    xyzz_t acc = xyzz_infinity();
    const uint32_t iters = (per + TREE_THREADS - 1) / TREE_THREADS;
    // steps [0, iters): strided loads; next 5: shuffle tree inside each warp; last 5: warp 0 folds the warp sums
    for (uint32_t step = 0; step < iters + 10; ++step) {
        xyzz_t rhs = xyzz_infinity();
        if (step < iters) {
            const uint32_t i = lo + step * TREE_THREADS + threadIdx.x;
            if (i < hi) rhs = xyzz_load(src + i);
        } else {
            const uint32_t s = (step - iters) % 5;
            if (step == iters + 5) {
                if (lane == 0) sm[warp] = acc;
                __syncthreads();
                acc = (warp == 0 && lane < TREE_THREADS / 32) ? sm[lane] : xyzz_infinity();
            }
            rhs = xyzz_shfl_down(acc, 16 >> s);
            if (lane + (16 >> s) >= 32) rhs = xyzz_infinity();
        }
        xyzz_add(acc, rhs);
    }
    if (threadIdx.x == 0) {
        xyzz_store(out + (size_t)row * parts + part, acc);
    }
}

// 6. result = sum_r 2^(c r) * set_sum[r], XYZZ -> Jacobian
__global__ void __launch_bounds__(32) k_msm_finish(const xyzz_t* __restrict__ set_sums, uint32_t S, uint32_t c, jac_t* __restrict__ out)
{
    if (threadIdx.x != 0) {
        return;
    }
    xyzz_t acc = xyzz_infinity();
    for (int r = (int)S - 1; r >= 0; --r) {
        if (r != (int)S - 1) {
            for (uint32_t d = 0; d < c; ++d) {
                acc = xyzz_dbl(acc);
            }
        }
        xyzz_t s = xyzz_load(set_sums + r);
        xyzz_add(acc, s);
    }
    jac_t j = xyzz_to_jacobian(acc);
    fe_store(&out->x, j.x);
    fe_store(&out->y, j.y);
    fe_store(&out->z, j.z);
}

__global__ void k_set_infinity(jac_t* out)
{
    jac_t j = xyzz_to_jacobian(xyzz_infinity());
    fe_store(&out->x, j.x);
    fe_store(&out->y, j.y);
    fe_store(&out->z, j.z);
}

// sum of Jacobian elements (bb/ecc/curves/bn254/scalar_multiplication/c_bind.cpp:40-45 g1_sum);
// one warp: lanes stride over the inputs, then a shuffle tree of g1 additions.
__global__ void __launch_bounds__(32) k_g1_sum(const jac_t* __restrict__ in, uint32_t n, jac_t* __restrict__ out)
{
    const unsigned lane = threadIdx.x;
    xyzz_t acc = xyzz_infinity();
    const uint32_t iters = (n + 31) / 32;
    for (uint32_t step = 0; step < iters + 5; ++step) {
        xyzz_t rhs = xyzz_infinity();
        if (step < iters) {
            const uint32_t i = step * 32 + lane;
            if (i < n) {
                jac_t j;
                j.x = fe_load<FqParams>(&in[i].x);
                j.y = fe_load<FqParams>(&in[i].y);
                j.z = fe_load<FqParams>(&in[i].z);
                rhs = xyzz_from_jacobian(j);
            }
        } else {
            const int d = 16 >> (step - iters);
            rhs = xyzz_shfl_down(acc, d);
            if (lane + d >= 32) rhs = xyzz_infinity();
        }
        xyzz_add(acc, rhs);
    }
    if (lane == 0) {
        jac_t j = xyzz_to_jacobian(acc);
        fe_store(&out->x, j.x);
        fe_store(&out->y, j.y);
        fe_store(&out->z, j.z);
    }
}

// ------------------------------------------------------------------------------------------------
// 0. fixed-base levels: table[l * stride + i] = 2^(D l) * P_i (affine), l = 1 .. L-1 (level 0 = the SRS itself)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_msm_precompute(affine_t* __restrict__ table, uint32_t n, uint32_t stride, uint32_t L, uint32_t D)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    affine_t p = affine_load(table + i);
    xyzz_t cur = xyzz_from_affine(p);
    for (uint32_t l = 1; l < L; ++l) {
        for (uint32_t d = 0; d < D; ++d) {
            cur = xyzz_dbl(cur);
        }
        affine_t a = xyzz_to_affine(cur);
        fe_store(&table[(size_t)l * stride + i].x, a.x);
        fe_store(&table[(size_t)l * stride + i].y, a.y);
        cur = xyzz_from_affine(a); // ZZ = ZZZ = 1 again
    }
}

// ------------------------------------------------------------------------------------------------
// SRS helpers: on-device transcript decode and Pippenger point table
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// raw transcript point (64 B: x then y, each 4 u64 limbs least-significant first, each limb big-endian,
// non-Montgomery) -> Montgomery affine  (bb/srs/io.cpp:47-67: bswap64 every limb + to_montgomery_form)
__global__ void __launch_bounds__(256) k_srs_decode(const uint32_t* __restrict__ raw, uint32_t n, affine_t* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const uint32_t* p = raw + (size_t)i * 16;
    affine_t r;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        // u64 big-endian limb = bytes b0..b7; little-endian u32 view: w0 = b0..b3, w1 = b4..b7
        uint32_t w0 = p[2 * l], w1 = p[2 * l + 1];
        r.x.l[2 * l] = bswap32(w1);
        r.x.l[2 * l + 1] = bswap32(w0);
        uint32_t y0 = p[8 + 2 * l], y1 = p[8 + 2 * l + 1];
        r.y.l[2 * l] = bswap32(y1);
        r.y.l[2 * l + 1] = bswap32(y0);
    }
    r.x = fe_to_mont(r.x);
    r.y = fe_to_mont(r.y);
    fe_store(&out[i].x, r.x);
    fe_store(&out[i].y, r.y);
}

// table[2i] = P_i ; table[2i+1] = (beta * x_i, -y_i)   (scalar_multiplication.cpp:104-112)
__global__ void __launch_bounds__(256) k_point_table(const affine_t* __restrict__ points, uint32_t n, affine_t* __restrict__ table)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    // beta: cube root of unity in fq, Montgomery form (bb/ecc/curves/bn254/fq.hpp:21-24)
    fq beta;
    beta.l[0] = 0xd782e155; beta.l[1] = 0x71930c11; beta.l[2] = 0xffbe3323; beta.l[3] = 0xa6bb947c;
    beta.l[4] = 0xd4741444; beta.l[5] = 0xaa303344; beta.l[6] = 0x26594943; beta.l[7] = 0x2c3b3f0d;
    affine_t p = affine_load(points + i);
    fe_store(&table[2 * (size_t)i].x, p.x);
    fe_store(&table[2 * (size_t)i].y, p.y);
    // the reference negates with 2p - y and multiplies x blindly (the infinity flag is not special-cased there either)
    Fe<FqParams> bx = fe_mul(p.x, beta);
    Fe<FqParams> p2, ny;
#pragma unroll
    for (int k = 0; k < 8; ++k) p2.l[k] = FqParams::P2(k);
    sub8(ny.l, p2.l, p.y.l);
    fe_store(&table[2 * (size_t)i + 1].x, bx);
    fe_store(&table[2 * (size_t)i + 1].y, ny);
}

// compact the even entries of a 2n interleaved table into n contiguous points
__global__ void __launch_bounds__(256) k_compact_even(const uint4* __restrict__ table, size_t n, uint4* __restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; // one uint4 (16 B) per thread, 4 per point
    if (i >= n * 4) {
        return;
    }
    size_t pt = i >> 2, q = i & 3;
    out[i] = table[pt * 8 + q];
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
static int exclusive_scan(Context* ctx, const uint32_t* d_in, size_t n, uint32_t* d_out, uint32_t* d_out_copy, cudaStream_t st)
{
    unsigned tiles = div_up(n, SCAN_TILE);
    if (tiles > SCAN_TILE) {
        set_last_error("scan too large");
        return BBG_ERR_ARG;
    }
    int rc = ctx->msm_scan_tmp.reserve((size_t)tiles * 4 + 16);
    if (rc) return rc;
    uint32_t* tmp = (uint32_t*)ctx->msm_scan_tmp.p;
    k_scan_tile_sums<<<tiles, SCAN_THREADS, 0, st>>>(d_in, n, tmp);
    k_scan_top<<<1, SCAN_THREADS, 0, st>>>(tmp, tiles);
    k_scan_apply<<<tiles, SCAN_THREADS, 0, st>>>(d_in, n, tmp, d_out, d_out_copy);
    ctx->launches += 3;
    return BBG_OK;
}

static unsigned env_uint(const char* name, unsigned dflt)
{
    const char* v = getenv(name);
    return v && *v ? (unsigned)atoi(v) : dflt;
}

static unsigned floor_log2(size_t n)
{
    unsigned lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) ++lg;
    return lg;
}

// window width for an SRS of `n` points when every level is precomputed (bucket reduction is cheap then)
static unsigned msm_window_bits(size_t n)
{
    int c = (int)floor_log2(n ? n : 1) - 4;
    if (c < 4) c = 4;
    if (c > 20) c = 20;
    return env_uint("BBG_MSM_C", (unsigned)c);
}

MsmLevels msm_levels_plan(size_t n, size_t max_table_bytes)
{
    MsmLevels lv;
    lv.c = msm_window_bits(n);
    const unsigned W = (255 + lv.c - 1) / lv.c;
    unsigned L = env_uint("BBG_MSM_LEVELS", W);
    if (L > W) L = W;
    if (L < 1) L = 1;
    while (L > 1 && (size_t)L * n * 64 > max_table_bytes) --L;
    const unsigned S = (W + L - 1) / L;
    lv.L = (W + S - 1) / S;
    lv.D = S * lv.c;
    lv.stride = n;
    return lv;
}

int msm_precompute_device(Context* ctx, void* d_table, size_t n, const MsmLevels& lv, cudaStream_t st)
{
    if (lv.L <= 1 || n == 0) return BBG_OK;
    k_msm_precompute<<<div_up(n, 128), 128, 0, st>>>((affine_t*)d_table, (uint32_t)n, (uint32_t)lv.stride, lv.L, lv.D);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

// Device-pointer MSM over table entries [base, base + n) of level 0 (and the same range of every other level).
//   lv.L == 1: plain points (any stride), W bucket sets.   lv.L > 1: fixed-base levels, S = D / c bucket sets.
int msm_device(Context* ctx, const void* d_scalars, size_t n, const void* d_points, size_t point_stride, const MsmLevels& lv_in,
               size_t base, void* d_out, cudaStream_t st)
{
    Profiler& pr = ctx->prof;
    pr.begin();
    if (n == 0) {
        k_set_infinity<<<1, 1, 0, st>>>((jac_t*)d_out);
        ctx->launches += 1;
        BBG_CUDA(cudaGetLastError());
        return BBG_OK;
    }
    MsmLevels lv = lv_in;
    if (lv.L <= 1) {
        lv.L = 1;
        lv.c = env_uint("BBG_MSM_C1", (unsigned)std::max(2, std::min(16, (int)floor_log2(n) - 4)));
        lv.stride = 0;
    }
    const unsigned c = lv.c;
    const unsigned W = (255 + c - 1) / c; // W*c >= 255: the top window absorbs the last carry
    const unsigned S = lv.L == 1 ? W : lv.D / c;
    const uint32_t B = 1u << (c - 1);
    const size_t G = (size_t)S * B;
    const size_t max_entries = (size_t)W * n;
    if (max_entries >= (1ull << 32) || ((size_t)(lv.L - 1) * lv.stride + base + n) >= (1ull << 31)) {
        set_last_error("msm: too many points for 32-bit schedule entries (shard the MSM by point range)");
        return BBG_ERR_ARG;
    }
    int rc;
    if ((rc = ctx->msm_counts.reserve((G + 1) * 4 + 64))) return rc;
    if ((rc = ctx->msm_offsets.reserve((G + 1) * 4))) return rc;
    if ((rc = ctx->msm_cursors.reserve((G + 1) * 4))) return rc;
    if ((rc = ctx->msm_sorted.reserve(max_entries * 4))) return rc;
    if ((rc = ctx->msm_buckets.reserve(G * sizeof(xyzz_t)))) return rc;
    uint32_t* counts = (uint32_t*)ctx->msm_counts.p;
    uint32_t* offsets = (uint32_t*)ctx->msm_offsets.p;
    uint32_t* cursors = (uint32_t*)ctx->msm_cursors.p;
    uint32_t* sorted = (uint32_t*)ctx->msm_sorted.p;
    xyzz_t* buckets = (xyzz_t*)ctx->msm_buckets.p;
    uint32_t* pending = counts + G + 1; // 15 merge-level counters behind the histogram (zeroed with it)

    pr.mark(st, PH_MSM_DIGITS);
    BBG_CUDA(cudaMemsetAsync(counts, 0, (G + 1) * 4 + 64, st));
    BBG_CUDA(cudaMemsetAsync(buckets, 0, G * sizeof(xyzz_t), st)); // all-zero XYZZ = infinity

    DigitParams dp;
    dp.n = (uint32_t)n;
    dp.c = c;
    dp.W = W;
    dp.S = S;
    dp.B = B;
    dp.level_stride = (uint32_t)lv.stride;
    dp.base = (uint32_t)base;
    const unsigned dig_blocks = div_up(n, 256);
    k_msm_digits<false><<<dig_blocks, 256, 0, st>>>((const fr_t*)d_scalars, dp, counts, nullptr);
    ctx->launches += 1;
    pr.mark(st, PH_MSM_SCAN);
    if ((rc = exclusive_scan(ctx, counts, G, offsets, cursors, st))) return rc;
    pr.mark(st, PH_MSM_SCATTER);
    k_msm_digits<true><<<dig_blocks, 256, 0, st>>>((const fr_t*)d_scalars, dp, cursors, sorted);
    ctx->launches += 1;

    // chunking: equal work per thread; about two waves of resident threads, chunk >= 16 entries.
    // The number of non-zero digits is only known on the device; size for the maximum.
    const size_t resident = (size_t)ctx->num_sms * ACC_THREADS * 4;
    size_t chunk = (max_entries + 2 * resident - 1) / (2 * resident);
    if (chunk < 16) chunk = 16;
    const size_t num_chunks = (max_entries + chunk - 1) / chunk;
    // slot arrays of all merge levels, back to back
    size_t slots_total = 2 * num_chunks;
    {
        size_t n_in = 2 * num_chunks;
        for (unsigned level = 0; level < 14; ++level) {
            const size_t workers = n_in > 1 ? (n_in - 1 + MERGE_K - 1) / MERGE_K : 1;
            slots_total += 2 * workers;
            if (workers == 1) break;
            n_in = 2 * workers;
        }
    }
    if ((rc = ctx->msm_partials.reserve(slots_total * sizeof(Slot)))) return rc;
    Slot* slots = (Slot*)ctx->msm_partials.p;

    pr.mark(st, PH_MSM_ACCUMULATE);
    k_msm_accumulate<<<div_up(num_chunks, ACC_THREADS), ACC_THREADS, 0, st>>>(
        sorted, offsets, (uint32_t)G, (uint32_t)chunk, (uint32_t)num_chunks, (const affine_t*)d_points, (uint32_t)point_stride,
        buckets, slots, pending);
    ctx->launches += 1;
    pr.mark(st, PH_MSM_FIXUP);
    {
        size_t n_in = 2 * num_chunks;
        Slot* in = slots;
        unsigned level = 0;
        while (level < 14) {
            const size_t workers = n_in > 1 ? (n_in - 1 + MERGE_K - 1) / MERGE_K : 1;
            Slot* out = in + n_in;
            k_msm_merge<<<div_up(workers, 128), 128, 0, st>>>(in, (uint32_t)n_in, (uint32_t)workers, pending + level, buckets, out,
                                                             pending + level + 1);
            ctx->launches += 1;
            ++level;
            if (workers == 1) break;
            in = out;
            n_in = 2 * workers;
        }
    }

    // bucket reduction
    pr.mark(st, PH_MSM_REDUCE);
    uint32_t ell = 16;
    while (ell > 1 && (G / ell) < 8192) ell >>= 1; // keep at least ~8K workers in flight
    if (ell > B) ell = B;
    const uint32_t segs = B / ell;
    const uint32_t workers = segs * S;
    const uint32_t parts = std::min<uint32_t>(64, std::max<uint32_t>(1, segs / 64));
    if ((rc = ctx->msm_reduce.reserve(((size_t)workers + (size_t)S * parts + S) * sizeof(xyzz_t)))) return rc;
    xyzz_t* seg_out = (xyzz_t*)ctx->msm_reduce.p;
    xyzz_t* part_out = seg_out + workers;
    xyzz_t* set_out = part_out + (size_t)S * parts;
    k_msm_segments<<<div_up(workers, 128), 128, 0, st>>>(buckets, B, ell, workers, seg_out);
    k_msm_tree_sum<<<dim3(parts, S), TREE_THREADS, 0, st>>>(seg_out, segs, part_out);
    ctx->launches += 2;
    const xyzz_t* sums = part_out;
    if (parts > 1) {
        k_msm_tree_sum<<<dim3(1, S), TREE_THREADS, 0, st>>>(part_out, parts, set_out);
        ctx->launches += 1;
        sums = set_out;
    }
    pr.mark(st, PH_MSM_COMBINE);
    k_msm_finish<<<1, 32, 0, st>>>(sums, S, c, (jac_t*)d_out);
    ctx->launches += 1;
    pr.mark(st, -1);
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

int g1_sum_device(Context* ctx, const void* d_jacs, size_t n, void* d_out, cudaStream_t st)
{
    k_g1_sum<<<1, 32, 0, st>>>((const jac_t*)d_jacs, (uint32_t)n, (jac_t*)d_out);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

int srs_decode_device(Context* ctx, const void* d_raw, size_t n, void* d_points, cudaStream_t st)
{
    if (n == 0) return BBG_OK;
    k_srs_decode<<<div_up(n, 256), 256, 0, st>>>((const uint32_t*)d_raw, (uint32_t)n, (affine_t*)d_points);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}
int point_table_device(Context* ctx, const void* d_points, size_t n, void* d_table, cudaStream_t st)
{
    if (n == 0) return BBG_OK;
    k_point_table<<<div_up(n, 256), 256, 0, st>>>((const affine_t*)d_points, (uint32_t)n, (affine_t*)d_table);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}
int compact_even_device(Context* ctx, const void* d_table, size_t n, void* d_points, cudaStream_t st)
{
    if (n == 0) return BBG_OK;
    k_compact_even<<<div_up(n * 4, 256), 256, 0, st>>>((const uint4*)d_table, n, (uint4*)d_points);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

} // namespace bbg
