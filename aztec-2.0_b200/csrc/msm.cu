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
#include "g1_team.cuh"
#include "internal.hpp"
#include "inv.cuh"

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
    uint32_t agg_from;     // windows >= agg_from are narrow (few distinct digits): warp-aggregated counter updates
    // fused batch: blockIdx.y = member; member m > 0 reads more[m - 1]; its windows go to bucket sets [m S, (m + 1) S)
    const void* more[3];
};

// MODE 0: histogram of bucket sizes.  MODE 1: counting-sort scatter of schedule words (table index << 1 | negate).
// MODE 2: counting-sort scatter of the POINTS themselves (sign applied): thread i reads entry i of every level --
// coalesced -- and the pairwise passes / the accumulation then stream the bucket-ordered copy instead of gathering.
enum { DIG_HISTOGRAM = 0, DIG_SCATTER_INDEX = 1, DIG_SCATTER_POINTS = 2 };
template <int MODE>
__global__ void __launch_bounds__(256) k_msm_digits(const fr_t* __restrict__ scalars,
                                                    const DigitParams P,
                                                    uint32_t* __restrict__ counters,
                                                    uint32_t* __restrict__ sorted,
                                                    const affine_t* __restrict__ points,
                                                    uint32_t point_stride,
                                                    affine_t* __restrict__ pts_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // no early exit: every lane of a warp walks the (uniform) window loop so that the aggregated path below can name
    // the full warp in its *_sync primitives; lanes past the end carry an all-zero scalar and never touch memory
    const bool valid = i < P.n;
    const uint32_t member = blockIdx.y;
    if (member != 0) scalars = (const fr_t*)P.more[member - 1];
    const uint32_t set_base = member * P.S;
    // from_montgomery_form => canonical integer in [0, r)  (scalar_multiplication.cpp:224)
    fr_t s = fe_zero<FrParams>();
    if (valid) s = fe_from_mont(fe_load_nc<FrParams>(scalars + i));
    uint32_t k[9];
#pragma unroll
    for (int j = 0; j < 8; ++j) k[j] = s.l[j];
    k[8] = 0;
    const uint32_t c = P.c;
    const uint32_t half = 1u << (c - 1);
    const uint32_t mask = (1u << c) - 1;
    const unsigned lane = threadIdx.x & 31;
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
        const bool live = valid && mag != 0;
        const uint32_t g = (set_base + set) * P.B + (mag - 1);
        uint32_t dst = 0;
        if (w >= P.agg_from) {
            // The top window of a 254-bit scalar can be a few bits wide (c = 18: two bits): a million atomics on
            // three addresses serialise in L2.  One atomic per distinct bucket per warp instead.
            const unsigned peers = __match_any_sync(0xffffffffu, live ? g : 0xffffffffu);
            const int leader = __ffs(peers) - 1;
            uint32_t first = 0;
            if (live && (int)lane == leader) first = atomicAdd(&counters[g], (uint32_t)__popc(peers));
            first = __shfl_sync(0xffffffffu, first, leader);
            dst = first + __popc(peers & ((1u << lane) - 1u));
        } else if (live) {
            dst = atomicAdd(&counters[g], 1u);
        }
        if (live) {
            if (MODE == DIG_SCATTER_INDEX) {
                sorted[dst] = ((level * P.level_stride + P.base + i) << 1) | neg;
            } else if (MODE == DIG_SCATTER_POINTS) {
                // (re)load per window: with fixed-base levels every window has its own point; with L == 1 the
                // reload hits L1
                affine_t pt = affine_load(points + (size_t)(level * P.level_stride + P.base + i) * point_stride);
                if (neg) pt.y = fe_neg(pt.y);
                fe_store(&pts_out[dst].x, pt.x);
                fe_store(&pts_out[dst].y, pt.y);
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

// The scans are batched over blockIdx.y: row y scans ceil(in[i] / 2^(shift_base + y)) -- row 0 with shift 0 is the
// plain bucket histogram, the other rows are the bucket sizes after y rounds of pairwise additions (k_msm_pair_pass).
__device__ __forceinline__ uint32_t ceil_shift(uint32_t v, uint32_t sh) { return (v + ((1u << sh) - 1u)) >> sh; }

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ tile_sums,
                                                                 uint32_t shift_base)
{
    __shared__ uint32_t smem[33];
    const uint32_t sh = shift_base + blockIdx.y;
    tile_sums += (size_t)blockIdx.y * SCAN_TILE;
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < n) s += ceil_shift(in[base + j], sh);
    }
    uint32_t total;
    block_exclusive_scan(s, smem, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_top(uint32_t* __restrict__ tile_sums, unsigned num_tiles)
{
    __shared__ uint32_t smem[33];
    tile_sums += (size_t)blockIdx.y * SCAN_TILE;
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
                                                             uint32_t* __restrict__ out_copy, // may be null
                                                             uint32_t shift_base)
{
    __shared__ uint32_t smem[33];
    const uint32_t sh = shift_base + blockIdx.y;
    tile_prefix += (size_t)blockIdx.y * SCAN_TILE;
    out += (size_t)blockIdx.y * (n + 1);
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        v[j] = (base + j < n) ? ceil_shift(in[base + j], sh) : 0;
        s += v[j];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, smem, total) + tile_prefix[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < n) {
            out[base + j] = ex;
            if (out_copy) out_copy[base + j] = ex;
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

// DIRECT: the entry at position pos IS points[pos] (the output of the pairwise passes); otherwise it is the schedule
// word sorted[pos] = (table index << 1) | negate.
// Launch bounds (128, 4): 128 registers with 12 spilled bytes and 16 warps per SM.  (128, 3) gives 148 registers, no spills
// and 12 warps: measured slower, 2.142 -> 2.177 ms at 2^20 (profiles/r2_msm_acc_ctas_ab.txt).
template <bool DIRECT>
__global__ void __launch_bounds__(ACC_THREADS, 4) k_msm_accumulate(const uint32_t* __restrict__ sorted,
                                                                   const uint32_t* __restrict__ offsets, // G+1
                                                                   uint32_t G,
                                                                   uint32_t g_lo, // this launch takes the entries of buckets [g_lo, g_hi)
                                                                   uint32_t g_hi,
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
    const uint32_t total = __ldg(offsets + g_hi); // end of the range's entries (g_hi = G: the number of non-zero digits)
    const uint64_t start64 = (uint64_t)__ldg(offsets + g_lo) + (uint64_t)t * chunk;
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
    uint32_t v = DIRECT ? 0u : __ldg(sorted + pos);
    affine_t pt = affine_load(DIRECT ? points + pos : points + (size_t)(v >> 1) * point_stride);
    while (true) {
        // prefetch the next entry's point while this one is being added
        uint32_t vn = 0;
        affine_t ptn;
        const bool more = pos + 1 < end;
        if (more) {
            vn = DIRECT ? 0u : __ldg(sorted + pos + 1);
            ptn = affine_load(DIRECT ? points + pos + 1 : points + (size_t)(vn >> 1) * point_stride);
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

// ------------------------------------------------------------------------------------------------
// 4a. pairwise affine additions with one shared inversion per CTA batch ("batched affine"; optional path)
// ------------------------------------------------------------------------------------------------
// The reference sums a bucket's points in affine coordinates, pairing them level by level and sharing one field
// inversion among all the additions of a level (scalar_multiplication.cpp:273-401 add_affine_points, :523-718
// evaluate_addition_chains): 5M + 1S per addition instead of 8M + 2S for a projective mixed addition.  Same idea
// here, laid out for the GPU: one launch per level; level j holds, bucket after bucket, the ceil(m / 2^j) partial
// sums of every bucket (offsets from the batched scan), output slot q of a bucket is input slots 2q and 2q+1 (or a
// copy of 2q when the count is odd).  A thread owns K consecutive output slots, a CTA batch is PAIR_THREADS * K
// slots:
//   forward : d_i = x2 - x1 per slot, running product stored per slot (Montgomery's trick)
//   combine : warp 0 multiplies the 8 thread products of each lane column, inverts the 32 column products -- one
//             inversion per lane, all lanes at once, on the ALU pipe (inv.cuh) -- and hands every thread the inverse
//             of its own product
//   backward: 1/d_i from the running inverse and the stored prefix, then lambda, x3, y3.
// Exceptional cases keep the batch intact by substituting the denominator: equal points double (d = 2y, numerator
// 3x^2), opposite points give infinity (d = 1), an infinite operand or a missing partner copies (d = 1).
static constexpr int PAIR_THREADS = 256;
static constexpr int PAIR_K_MAX = 64;
static constexpr int PAIR_WARPS = PAIR_THREADS / 32;
enum : uint32_t { PK_NONE = 0, PK_ADD = 1, PK_DBL = 2, PK_COPY1 = 3, PK_COPY2 = 4, PK_INF = 5 };

// The running products and the per-slot bookkeeping of a batch live in global scratch indexed by output slot (they
// are written in the forward sweep and read back once in the backward sweep, mostly out of L2): a thread can then
// own tens of slots, which is what amortises the inversion.
struct PairSmem {
    uint4 tp[2][PAIR_THREADS];   // thread products, then their inverses
    uint4 wp[2][PAIR_WARPS][32]; // per lane column: products over warps 0..w
};

__device__ __forceinline__ void sm_put(uint4* lo, uint4* hi, const fq& v)
{
    *lo = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    *hi = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ fq sm_get(const uint4* lo, const uint4* hi)
{
    const uint4 a = *lo, b = *hi;
    fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// Every lane of a full warp inverts its own Montgomery-form value (non-zero mod p): result is the Montgomery form
// of the inverse.  fix = the C_k table of inv.cuh.
__device__ __forceinline__ fq fq_inv_warp(const fq& a, const uint32_t* __restrict__ fix)
{
    const fq ar = fe_reduce_once(a);
    uint32_t pl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pl[i] = FqParams::P(i);
    inv::Kaliski st;
    st.init(ar.l, pl);
#pragma unroll 1
    while (__any_sync(0xffffffffu, st.alive())) {
        st.step();
    }
    fq y;
    st.finish(y.l, pl);
    uint32_t k = st.k < inv::K_MIN ? inv::K_MIN : (st.k > inv::K_MAX ? inv::K_MAX : st.k);
    const fq c = fe_load_nc<FqParams>(fix + (size_t)(k - inv::K_MIN) * 8);
    return fe_mul(y, c);
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <bool PASS0> struct PairSrc {
    const uint32_t* sorted;
    const affine_t* pts;
    uint32_t stride;
    __device__ __forceinline__ fq x(uint32_t pos) const
    {
        if (PASS0) {
            const uint32_t v = __ldg(sorted + pos);
            return fe_load_nc<FqParams>(&pts[(size_t)(v >> 1) * stride].x);
        }
        return fe_load_nc<FqParams>(&pts[pos].x);
    }
    __device__ __forceinline__ affine_t point(uint32_t pos) const
    {
        if (PASS0) {
            const uint32_t v = __ldg(sorted + pos);
            affine_t p = affine_load(pts + (size_t)(v >> 1) * stride);
            if (v & 1) p.y = fe_neg(p.y);
            return p;
        }
        return affine_load(pts + pos);
    }
};

// PASS0: the inputs are schedule words (gather from the level tables); otherwise the previous level's points.
template <bool PASS0>
__global__ void __launch_bounds__(PAIR_THREADS, 2) k_msm_pair_pass(const uint32_t* __restrict__ sorted,
                                                                   const affine_t* __restrict__ in,
                                                                   uint32_t point_stride,
                                                                   const uint32_t* __restrict__ off_in,  // G+1
                                                                   const uint32_t* __restrict__ off_out, // G+1
                                                                   uint32_t G,
                                                                   affine_t* __restrict__ out,
                                                                   const uint32_t* __restrict__ inv_fix,
                                                                   uint32_t K,                 // slots per thread
                                                                   uint4* __restrict__ pre,    // 2 x uint4 per output slot
                                                                   uint2* __restrict__ meta)   // (input position, kind) per slot
{
    __shared__ PairSmem sm;
    const uint32_t tid = threadIdx.x;
    const uint32_t total = __ldg(off_out + G);
    const uint64_t cta_base = (uint64_t)blockIdx.x * ((uint64_t)PAIR_THREADS * K);
    if (cta_base >= total) {
        return; // whole CTA past the end (the grid is sized for the worst case)
    }
    const uint64_t o0_64 = cta_base + (uint64_t)tid * K;
    const uint32_t o0 = (uint32_t)(o0_64 < total ? o0_64 : total);
    const uint32_t n_slots = min(K, total - o0);
    const PairSrc<PASS0> src{ sorted, in, point_stride };
    const uint32_t in_total = __ldg(off_in + G);

    // ---- forward
    fq run = fe_one<FqParams>();
    {
        uint32_t b = 0, ob = 0, oe = 0, ib = 0, ie = 0;
        if (n_slots) {
            b = upper_bound_u32(off_out, G + 1, o0) - 1;
            ob = __ldg(off_out + b);
            oe = __ldg(off_out + b + 1);
            ib = __ldg(off_in + b);
            ie = __ldg(off_in + b + 1);
        }
#pragma unroll 1
        for (uint32_t i = 0; i < n_slots; ++i) {
            uint32_t kind = PK_NONE, p1 = 0;
            {
                const uint32_t o = o0 + i;
                if (o >= oe) {
                    do { // next non-empty bucket
                        ++b;
                        ob = oe;
                        oe = __ldg(off_out + b + 1);
                    } while (oe == ob);
                    ib = __ldg(off_in + b);
                    ie = __ldg(off_in + b + 1);
                }
                p1 = ib + 2 * (o - ob);
                if (!PASS0 && p1 + 5 < in_total) { // the next slots' operands follow in memory: pull them into L1 early
                    prefetch_l1(in + p1 + 4);
                    prefetch_l1(in + p1 + 5);
                }
                const fq x1 = src.x(p1);
                if (p1 + 1 >= ie) {
                    kind = PK_COPY1;
                } else {
                    const fq x2 = src.x(p1 + 1);
                    if (x1.l[7] & INF_BIT) {
                        kind = PK_COPY2;
                    } else if (x2.l[7] & INF_BIT) {
                        kind = PK_COPY1;
                    } else {
                        fq d = fe_sub(x2, x1);
                        kind = PK_ADD;
                        if (__builtin_expect(fe_is_zero(d), 0)) {
                            const affine_t a = src.point(p1), c = src.point(p1 + 1);
                            if (fe_is_zero(fe_sub(a.y, c.y))) {
                                kind = PK_DBL;
                                d = fe_dbl(a.y); // y != 0 on a curve of odd order
                            } else {
                                kind = PK_INF;
                            }
                        }
                        if (kind != PK_INF) {
                            run = fe_mul(run, d);
                        }
                    }
                }
            }
            sm_put(&pre[2 * (size_t)(o0 + i)], &pre[2 * (size_t)(o0 + i) + 1], run);
            meta[o0 + i] = make_uint2(p1, kind);
        }
    }

    // ---- combine: inverse of every thread's product
    sm_put(&sm.tp[0][tid], &sm.tp[1][tid], run);
    __syncthreads();
    if (tid < 32) {
        fq acc = sm_get(&sm.tp[0][tid], &sm.tp[1][tid]);
        sm_put(&sm.wp[0][0][tid], &sm.wp[1][0][tid], acc);
#pragma unroll 1
        for (uint32_t w = 1; w < PAIR_WARPS; ++w) {
            acc = fe_mul(acc, sm_get(&sm.tp[0][w * 32 + tid], &sm.tp[1][w * 32 + tid]));
            sm_put(&sm.wp[0][w][tid], &sm.wp[1][w][tid], acc);
        }
        fq inv_acc = fq_inv_warp(acc, inv_fix);
#pragma unroll 1
        for (uint32_t w = PAIR_WARPS - 1; w >= 1; --w) {
            const fq t = sm_get(&sm.tp[0][w * 32 + tid], &sm.tp[1][w * 32 + tid]);
            const fq below = sm_get(&sm.wp[0][w - 1][tid], &sm.wp[1][w - 1][tid]);
            sm_put(&sm.tp[0][w * 32 + tid], &sm.tp[1][w * 32 + tid], fe_mul(inv_acc, below));
            inv_acc = fe_mul(inv_acc, t);
        }
        sm_put(&sm.tp[0][tid], &sm.tp[1][tid], inv_acc);
    }
    __syncthreads();

    // ---- backward
    fq inv_run = sm_get(&sm.tp[0][tid], &sm.tp[1][tid]);
#pragma unroll 1
    for (int i = (int)n_slots - 1; i >= 0; --i) {
        const uint2 mt = meta[o0 + (uint32_t)i];
        const uint32_t kind = mt.y;
        const uint32_t p1 = mt.x;
        if (!PASS0 && i >= 2) { // descending sweep: the slots before this one
            if (p1 >= 4) {
                prefetch_l1(in + p1 - 4);
                prefetch_l1(in + p1 - 3);
            }
            prefetch_l1(&pre[2 * (size_t)(o0 + i - 2)]);
        }
        affine_t* dst = out + (o0 + (uint32_t)i);
        if (kind >= PK_COPY1) {
            affine_t r;
            if (kind == PK_INF) {
                r.x = fe_zero<FqParams>();
                r.y = fe_zero<FqParams>();
                r.x.l[7] |= INF_BIT;
            } else {
                r = src.point(kind == PK_COPY1 ? p1 : p1 + 1);
            }
            fe_store(&dst->x, r.x);
            fe_store(&dst->y, r.y);
            continue;
        }
        const affine_t a = src.point(p1);
        fq x2, d, num;
        if (kind == PK_ADD) {
            const affine_t c = src.point(p1 + 1);
            x2 = c.x;
            d = fe_sub(c.x, a.x);
            num = fe_sub(c.y, a.y);
        } else {
            x2 = a.x;
            d = fe_dbl(a.y);
            const fq xx = fe_sqr(a.x);
            num = fe_add(fe_dbl(xx), xx);
        }
        fq inv_d = inv_run;
        if (i > 0) {
            inv_d = fe_mul(inv_run, sm_get(&pre[2 * (size_t)(o0 + i - 1)], &pre[2 * (size_t)(o0 + i - 1) + 1]));
            inv_run = fe_mul(inv_run, d);
        }
        const fq lam = fe_mul(num, inv_d);
        const fq x3 = fe_sub(fe_sub(fe_sqr(lam), a.x), x2);
        const fq y3 = fe_sub(fe_mul(lam, fe_sub(a.x, x3)), a.y);
        fe_store(&dst->x, x3);
        fe_store(&dst->y, y3);
    }
}

// One merge level: worker u folds slots [u*K + 1, (u+1)*K + 1) (worker 0 also takes slot 0), i.e. the
// boundary falls between the two slots of one chunk so the typical (tail, next head) pair is never cut.
// A worker is a TEAM of four lanes (g1_team.cuh): the chain of up to MERGE_K - 1 additions per level is what this phase
// costs, and there are far fewer workers than lanes.  All lanes of a team read the same slots and take the same branches;
// lane 0 writes.
// TEAM = false: the worker is ONE thread with plain additions -- level 0 of a large MSM has ~57 k workers with ~7 additions
// each, i.e. it is throughput bound, and the cooperative form spends 16 multiplies of four lanes where 14 of one do.
template <bool TEAM = true>
__device__ __forceinline__ void merge_worker(const Team& tm, const Slot* __restrict__ in, uint32_t n_in, uint32_t workers, uint32_t u,
                                             xyzz_t* __restrict__ buckets, Slot* __restrict__ out, uint32_t* __restrict__ pending_out)
{
    const bool writer = !TEAM || tm.r == 0;
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
                if (writer) xyzz_store(buckets + run_bucket, acc);
            } else {
                if (cont_l) {
                    if (writer) slot_store(out_a, acc, run_bucket);
                    a_set = true;
                }
                if (cont_r) {
                    if (writer) slot_store(out_b, cont_l ? xyzz_infinity() : acc, run_bucket);
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
                // the one inlined add site of this kernel
                if constexpr (TEAM) {
                    xyzz_add_team(tm, acc, x);
                } else {
                    xyzz_add(acc, x);
                }
            }
        }
    }
    if (writer) {
        if (!a_set) out_a->bucket = SLOT_NONE;
        if (!b_set) out_b->bucket = SLOT_NONE;
        if (a_set || b_set) {
            atomicAdd(pending_out, 1u);
        }
    }
}

// levels 0-2: one team per worker over the slots the level before left
template <bool TEAM>
__global__ void __launch_bounds__(128) k_msm_merge(const Slot* __restrict__ in,
                                                    uint32_t n_in,
                                                    uint32_t workers,
                                                    const uint32_t* __restrict__ pending_in,
                                                    xyzz_t* __restrict__ buckets,
                                                    Slot* __restrict__ out,
                                                    uint32_t* __restrict__ pending_out)
{
    if (__ldg(pending_in) == 0) {
        return; // nothing was left open
    }
    const Team tm = team_of_lane();
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t u = TEAM ? gtid >> 2 : gtid;
    if (u >= workers) {
        return; // team-uniform
    }
    merge_worker<TEAM>(tm, in, n_in, workers, u, buckets, out, pending_out);
}

// Levels 0-2 are ordinary grid-wide launches of k_msm_merge (each exits at once when the level before left nothing
// open); whatever is still open after them -- only when a handful of buckets hold most of the digits, e.g. all scalars
// equal -- is finished by ONE single-CTA launch that loops over the remaining levels with a block barrier in between
// (level 3 has at most num_chunks / 256 workers).  Round 1 launched up to 14 levels one by one.
static constexpr int MERGE_GRID_LEVELS = 3;
static constexpr int MERGE_REST_THREADS = 256; // 64 teams
__global__ void __launch_bounds__(MERGE_REST_THREADS) k_msm_merge_rest(Slot* __restrict__ slots, uint32_t n_in, uint32_t first_level,
                                                                        uint32_t* __restrict__ pending, xyzz_t* __restrict__ buckets)
{
    const Team tm = team_of_lane();
    Slot* in = slots;
    for (uint32_t level = first_level; level < 15; ++level) {
        if (*(volatile uint32_t*)(pending + level) == 0) {
            break; // uniform: every thread reads the same counter after the barrier
        }
        const uint32_t workers = n_in > 1 ? (n_in - 1 + MERGE_K - 1) / MERGE_K : 1;
        Slot* out = in + n_in;
        for (uint32_t u = threadIdx.x >> 2; u < workers; u += MERGE_REST_THREADS / 4) {
            merge_worker(tm, in, n_in, workers, u, buckets, out, pending + level + 1);
        }
        __threadfence();
        __syncthreads();
        if (workers == 1) {
            break;
        }
        in = out;
        n_in = 2 * workers;
    }
}

// ------------------------------------------------------------------------------------------------
// 5. bucket reduction: sum over b of (b + 1) * bucket[b] per set
// ------------------------------------------------------------------------------------------------
// Two levels of running sums.  Cut the B items of a set into segments of `ell0`:  with T_t = sum of segment t and
// R_t = sum_j (j + 1) x_{t ell0 + j} (both fall out of one running-sum sweep, 2 ell0 additions),
//     sum_b (b + 1) x_b  =  sum_t R_t  +  ell0 * sum_{t >= 1} t T_t,
// and the last term is the SAME problem on the array T[1..] (weights 1, 2, ...), a factor ell0 smaller.  Level 0
// (k_msm_segments<false>) is the throughput-bound sweep over all buckets; level 1 (k_msm_segments<true>) works on
// T[1..] and finishes each segment with the double-and-add by its offset, which is now paid by B / (ell0 ell1)
// workers instead of B / ell.  Everything here is latency bound (one g1 addition by a lone thread is ~6 us), so the
// design minimises the serial chain: 2 ell0 + 2 ell1 + log2(B / ell0) additions/doublings, then one batched tree
// sum over both levels and a short Horner step in k_msm_finish.
// Worker (set, seg): items [seg*ell, min((seg+1)*ell, count)) of in + set*in_stride.  One add site: the loop
// alternates run += x / acc += run.
// Every worker below is a TEAM of four adjacent lanes (g1_team.cuh): the chains of this phase are far shorter than the
// machine is wide, so each addition is spread over four lanes (4 multiply latencies instead of 14).
static constexpr int SEG_THREADS = 128; // 32 teams per CTA
// TEAM: a worker is a team of four lanes (cooperative additions: shortest chain) -- the right shape while there are
// fewer workers than the GPU has lanes.  !TEAM: one thread per worker (plain additions: ~14 % fewer multiplies, no
// shuffles) -- the right shape for level 0 of a large bucket set, which is throughput bound.
template <bool OFFSET, bool TEAM>
__global__ void __launch_bounds__(SEG_THREADS, TEAM ? 1 : 3) k_msm_segments(const xyzz_t* __restrict__ in,
                                                               uint32_t in_stride, // items between consecutive sets
                                                               uint32_t count,     // items per set
                                                               uint32_t ell,       // segment length
                                                               uint32_t segs,      // segments per set = ceil(count / ell)
                                                               uint32_t num_workers,
                                                               xyzz_t* __restrict__ out_r,
                                                               xyzz_t* __restrict__ out_t)
{
    const Team tm = team_of_lane();
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t t = TEAM ? gtid >> 2 : gtid;
    if (t >= num_workers) {
        return; // team-uniform
    }
    const bool writer = !TEAM || tm.r == 0;
    const uint32_t set = t / segs, seg = t % segs;
    const xyzz_t* src = in + (size_t)set * in_stride;
    const uint32_t first = seg * ell;
    xyzz_t run = xyzz_infinity(), acc = xyzz_infinity();
    for (uint32_t step = 0; step < 2 * ell; ++step) {
        const bool second = step & 1;
        xyzz_t rhs;
        if (!second) {
            const uint32_t idx = first + (ell - 1 - (step >> 1));
            rhs = xyzz_infinity();
            if (idx < count) rhs = xyzz_load(src + idx);
        } else {
            rhs = run;
        }
        xyzz_t lhs = xyzz_select(second, acc, run);
        if (TEAM) {
            xyzz_add_team(tm, lhs, rhs);
        } else {
            xyzz_add(lhs, rhs);
        }
        if (second) {
            acc = lhs;
        } else {
            run = lhs;
        }
    }
    if (OFFSET) {
        // V = acc + first * run by double-and-add
        const uint32_t k = first;
        if (k != 0 && !xyzz_is_inf(run)) {
            const int msb = 31 - __clz(k);
            xyzz_t m = run;
            for (int bit = msb - 1; bit >= -1; --bit) {
                xyzz_t rhs;
                bool do_add;
                if (bit >= 0) {
                    if (TEAM) {
                        xyzz_dbl_team(tm, m);
                    } else {
                        m = xyzz_dbl(m);
                    }
                    rhs = run;
                    do_add = (k >> bit) & 1;
                } else {
                    rhs = acc; // last step: fold in A
                    do_add = true;
                }
                if (do_add) {
                    if (TEAM) {
                        xyzz_add_team(tm, m, rhs);
                    } else {
                        xyzz_add(m, rhs);
                    }
                }
            }
            acc = m;
        }
        if (writer) xyzz_store(out_r + t, acc);
    } else if (writer) {
        xyzz_store(out_r + t, acc);
        xyzz_store(out_t + t, run);
    }
}

// Rows of the batched tree sum: row = level * S + set; level `l` has m[l] entries per set starting at off[l].
static constexpr int REDUCE_MAX_LEVELS = 24;
struct ReduceRows {
    uint32_t S;
    uint32_t levels;
    uint32_t off[REDUCE_MAX_LEVELS]; // entry offset of level l's R array (set-major, m[l] per set)
    uint32_t m[REDUCE_MAX_LEVELS];
    uint32_t mult[REDUCE_MAX_LEVELS];  // ell of level l (segment length: the factor between level l + 1 and level l); any integer >= 1
};

// lane-to-lane move of a whole point between TEAMS of one warp (delta in lanes, a multiple of 4)
__device__ __forceinline__ xyzz_t xyzz_shfl_down_warp(const xyzz_t& p, int delta) { return xyzz_shfl_down(p, delta); }

// Sum of the XYZZ points of every row: grid (parts, rows); the 64 teams of a CTA stride over its share, then a tree:
// 3 steps across the 8 teams of each warp, one shared-memory hop, 3 steps across the warps.  out[row * parts + part].
static constexpr int TREE_THREADS = 256;
static constexpr int TREE_TEAMS = TREE_THREADS / 4;
__global__ void __launch_bounds__(TREE_THREADS) k_msm_tree_sum(const xyzz_t* __restrict__ in, const ReduceRows rows, xyzz_t* __restrict__ out)
{
    __shared__ xyzz_t sm[TREE_THREADS / 32];
    const Team tm = team_of_lane();
    const uint32_t parts = gridDim.x, part = blockIdx.x, row = blockIdx.y;
    const uint32_t level = row / rows.S, set = row % rows.S;
    const uint32_t m = rows.m[level];
    const xyzz_t* src = in + rows.off[level] + (size_t)set * m;
    const uint32_t per = (m + parts - 1) / parts;
    const uint32_t lo = part * per;
    const uint32_t hi = min(lo + per, m);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned team = threadIdx.x >> 2; // 0..63
    if (lo >= hi) { // uniform per CTA
        if (threadIdx.x == 0) xyzz_store(out + (size_t)row * parts + part, xyzz_infinity());
        return;
    }
    xyzz_t acc = xyzz_infinity();
    const uint32_t iters = (per + TREE_TEAMS - 1) / TREE_TEAMS;
    // steps [0, iters): strided loads; next 3: tree across the teams of a warp; last 3: warp 0 folds the 8 warp sums.
    // The step count is uniform over the CTA, so the full-warp shuffles that move points between teams are legal.
    for (uint32_t step = 0; step < iters + 6; ++step) {
        xyzz_t rhs = xyzz_infinity();
        if (step < iters) {
            const uint32_t i = lo + step * TREE_TEAMS + team;
            if (i < hi) rhs = xyzz_load(src + i);
        } else {
            uint32_t delta = 16u >> (step - iters); // lanes: 16, 8, 4
            if (step >= iters + 3) {
                if (step == iters + 3) {
                    if (lane == 0) sm[warp] = acc;
                    __syncthreads();
                    acc = (warp == 0 && (lane >> 2) < TREE_THREADS / 32) ? sm[lane >> 2] : xyzz_infinity();
                }
                delta = 16u >> (step - iters - 3);
            }
            rhs = xyzz_shfl_down_warp(acc, (int)delta);
            if (lane + delta >= 32) rhs = xyzz_infinity();
        }
        xyzz_add_team(tm, acc, rhs);
    }
    if (threadIdx.x == 0) {
        xyzz_store(out + (size_t)row * parts + part, acc);
    }
}

// 6. per set: V = R_0 + ell_0 (R_1 + ell_1 (R_2 + ...)) (one team per set), then
//    result = sum_r 2^(c r) * V[r] (team 0), XYZZ -> Jacobian.  level_sums[level * S + set].
static constexpr int FINISH_THREADS = 256; // 64 teams; more bucket sets than that are taken in turns
//    Part of a bucket set (weight_offset != 0): the buckets handed in are [weight_offset, weight_offset + B') of their
//    set, so weight_offset * (plain sum of the buckets, row `plain_level`) is added; out_xyzz != nullptr: the result
//    stays in XYZZ for k_msm_parts_sum.
__global__ void __launch_bounds__(FINISH_THREADS) k_msm_finish(const xyzz_t* __restrict__ level_sums, const ReduceRows rows, uint32_t c,
                                                               jac_t* __restrict__ out, uint32_t plain_level, uint32_t weight_offset,
                                                               xyzz_t* __restrict__ out_xyzz, uint32_t members)
{
    extern __shared__ xyzz_t sm_sets[]; // S entries
    const Team tm = team_of_lane();
    const uint32_t S = rows.S;
    const uint32_t team = threadIdx.x >> 2;
    for (uint32_t set = team; set < S; set += FINISH_THREADS / 4) {
        // Horner from the deepest level
        xyzz_t v = xyzz_infinity();
        for (int level = (int)rows.levels - 1; level >= 0; --level) {
            if (level != (int)rows.levels - 1 && rows.mult[level] > 1) {
                // v *= ell (double-and-add; a power of two is doublings only)
                const uint32_t k = rows.mult[level];
                const xyzz_t base = v;
                for (int bit = 30 - __clz(k); bit >= 0; --bit) {
                    xyzz_dbl_team(tm, v);
                    if ((k >> bit) & 1) xyzz_add_team(tm, v, base);
                }
            }
            xyzz_t r = xyzz_load(level_sums + (size_t)level * S + set);
            xyzz_add_team(tm, v, r);
        }
        if (weight_offset != 0) {
            const xyzz_t plain = xyzz_load(level_sums + (size_t)plain_level * S + set);
            xyzz_t m = plain;
            for (int bit = 30 - __clz(weight_offset); bit >= 0; --bit) {
                xyzz_dbl_team(tm, m);
                if ((weight_offset >> bit) & 1) xyzz_add_team(tm, m, plain);
            }
            xyzz_add_team(tm, v, m);
        }
        if (tm.r == 0) sm_sets[set] = v;
    }
    __syncthreads();
    // fused batch: the S sets are `members` groups of S / members; team m combines group m into out[m]
    if (team >= members) {
        return;
    }
    const uint32_t per = S / members;
    xyzz_t acc = xyzz_infinity();
    for (int r = (int)per - 1; r >= 0; --r) {
        if (r != (int)per - 1) {
            for (uint32_t d = 0; d < c; ++d) {
                xyzz_dbl_team(tm, acc);
            }
        }
        xyzz_t s = sm_sets[team * per + r];
        xyzz_add_team(tm, acc, s);
    }
    if (tm.r == 0) {
        if (out_xyzz != nullptr) {
            xyzz_store(out_xyzz + team, acc);
            return;
        }
        jac_t j = xyzz_to_jacobian(acc);
        fe_store(&out[team].x, j.x);
        fe_store(&out[team].y, j.y);
        fe_store(&out[team].z, j.z);
    }
}

// sum of the per-part results of one MSM (msm_device: parts), XYZZ -> Jacobian; one team
__global__ void __launch_bounds__(32) k_msm_parts_sum(const xyzz_t* __restrict__ parts, uint32_t count, jac_t* __restrict__ out)
{
    const Team tm = team_of_lane();
    if (threadIdx.x >= 4) {
        return;
    }
    xyzz_t acc = xyzz_load(parts);
    for (uint32_t i = 1; i < count; ++i) {
        const xyzz_t x = xyzz_load(parts + i);
        xyzz_add_team(tm, acc, x);
    }
    if (tm.r == 0) {
        jac_t j = xyzz_to_jacobian(acc);
        fe_store(&out->x, j.x);
        fe_store(&out->y, j.y);
        fe_store(&out->z, j.z);
    }
}

__global__ void k_set_infinity(jac_t* out)
{
    jac_t j = xyzz_to_jacobian(xyzz_infinity());
    out += blockIdx.x;
    fe_store(&out->x, j.x);
    fe_store(&out->y, j.y);
    fe_store(&out->z, j.z);
}

// sum of Jacobian elements (bb/ecc/curves/bn254/scalar_multiplication/c_bind.cpp:40-45 g1_sum);
// one warp = 8 teams: teams stride over the inputs, then a 3-step tree of cooperative g1 additions.
__global__ void __launch_bounds__(32) k_g1_sum(const jac_t* __restrict__ in, uint32_t n, jac_t* __restrict__ out)
{
    const Team tm = team_of_lane();
    const unsigned lane = threadIdx.x, team = lane >> 2;
    xyzz_t acc = xyzz_infinity();
    const uint32_t iters = (n + 7) / 8;
    for (uint32_t step = 0; step < iters + 3; ++step) {
        xyzz_t rhs = xyzz_infinity();
        if (step < iters) {
            const uint32_t i = step * 8 + team;
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
        xyzz_add_team(tm, acc, rhs);
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
// rows >= 1: row y of d_out (n + 1 entries each) = exclusive scan of ceil(d_in / 2^(shift_base + y))
static int exclusive_scan(Context* ctx, MsmWorkspace& ws, const uint32_t* d_in, size_t n, uint32_t* d_out, uint32_t* d_out_copy,
                          cudaStream_t st, unsigned rows = 1, unsigned shift_base = 0)
{
    unsigned tiles = div_up(n, SCAN_TILE);
    if (tiles > SCAN_TILE) {
        set_last_error("scan too large");
        return BBG_ERR_ARG;
    }
    int rc = ws.scan_tmp.reserve((size_t)rows * SCAN_TILE * 4 + 16);
    if (rc) return rc;
    uint32_t* tmp = (uint32_t*)ws.scan_tmp.p;
    k_scan_tile_sums<<<dim3(tiles, rows), SCAN_THREADS, 0, st>>>(d_in, n, tmp, shift_base);
    k_scan_top<<<dim3(1, rows), SCAN_THREADS, 0, st>>>(tmp, tiles);
    k_scan_apply<<<dim3(tiles, rows), SCAN_THREADS, 0, st>>>(d_in, n, tmp, d_out, d_out_copy, shift_base);
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

// C_k table of inv.cuh for fq, built once per context
static int ensure_inv_fix(Context* ctx)
{
    if (ctx->inv_fix_fq != nullptr) return BBG_OK;
    uint32_t p[8], r2[8];
    for (int i = 0; i < 8; ++i) {
        p[i] = FqParams::P(i);
        r2[i] = FqParams::R2(i);
    }
    std::vector<uint32_t> tab((size_t)inv::FIX_ENTRIES * 8);
    inv::inv_fix_table(p, r2, tab.data());
    void* d = nullptr;
    BBG_CUDA(cudaMalloc(&d, tab.size() * 4));
    BBG_CUDA(cudaMemcpy(d, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    ctx->inv_fix_fq = d;
    return BBG_OK;
}

// window width for an SRS of `n` points when every level is precomputed (one bucket set of 2^(c-1) buckets).
// Cost model: W(c) n mixed additions (10 mul each, throughput bound) + a bucket reduction whose cost is ~60 serial g1
// additions of latency plus 2 full additions per bucket.  Widths whose TOP window holds only a few scalar bits
// (c = 14, 18, 19: 254 - (W - 1) c = 2, 2, 7) are skipped: that window's n digits fall into a handful of buckets,
// which the equal-chunk accumulation handles correctly but pays for in merge levels.
// Measured on B200 (ms per MSM, uniform scalars):  n = 2^16: c 14 1.03, 15 0.82, 16 0.78;  2^18: c 16 1.40, 17 1.30,
// 18 1.54;  2^20: c 16 3.65, 17 3.37, 18 3.61, 19 3.71, 20 3.27;  2^23: c 20 20.2.
static unsigned msm_window_bits(size_t n)
{
    const int lg = (int)floor_log2(n ? n : 1);
    int c;
    if (lg >= 20) c = 20;
    else if (lg >= 17) c = 17;
    else if (lg >= 15) c = 16;
    else if (lg >= 13) c = 15;
    else c = 13;
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
int msm_device(Context* ctx, MsmWorkspace& ws, const void* d_scalars, size_t n, const void* d_points, size_t point_stride,
               const MsmLevels& lv_in, size_t base, void* d_out, cudaStream_t st, const MsmArrival* arrival, bool allow_parts,
               const MsmBatch* batch)
{
    Profiler& pr = ctx->prof;
    pr.begin();
    const unsigned members = batch != nullptr ? batch->count : 1u;
    if (members < 1 || members > 4) {
        set_last_error("msm: a fused batch holds 1..4 scalar vectors");
        return BBG_ERR_ARG;
    }
    if (batch != nullptr) d_scalars = batch->scalars[0];
    if (n == 0) {
        k_set_infinity<<<members, 1, 0, st>>>((jac_t*)d_out);
        ctx->launches += 1;
        BBG_CUDA(cudaGetLastError());
        return BBG_OK;
    }
    MsmLevels lv = lv_in;
    if (lv.L <= 1) {
        lv.L = 1;
        // plain points: W bucket sets of 2^(c-1) buckets and a 255-doubling Horner at the end whatever c is; windows narrower
        // than 8 bits only add Horner additions (the verifier's 27-point MSMs ran with c = 2: 128 sets, 128 additions)
        lv.c = env_uint("BBG_MSM_C1", (unsigned)std::max(8, std::min(16, (int)floor_log2(n) - 4)));
        lv.stride = 0;
    }
    if (lv.L > 1 && env_uint("BBG_MSM_CALL_WINDOW", 1) != 0) {
        // A range much shorter than the object was built for (Pippenger::pippenger_unsafe(scalars, from, range) over a
        // piece of a large SRS): the object's window would leave 2^(c-1) mostly empty buckets to reduce.  The levels are
        // 2^(D l) P, so any c' = D / k works with k bucket sets per level (set s carries weight 2^(c' s), applied by
        // k_msm_finish): take the k with the least arithmetic, in fq multiplies: W n x 10 (mixed additions) + 2 x 14 per bucket.
        // Measured on a 2^20-point object (D = 20; ms with c = 20 -> c = 10, two sets): 2^12 points 0.94 -> 0.40, 2^14 0.90 ->
        // 0.51, but 2^16 0.91 -> 1.10: with ~1 700 digits per bucket the histogram / scatter atomics pile up on 1 024 counters
        // and every bucket is cut by dozens of chunks, so the switch is limited to <= 600 digits per bucket.
        auto cost = [&](unsigned cc) {
            const double Wc = (double)((255 + cc - 1) / cc);
            return Wc * (double)n * 10.0 + (double)(lv.D / cc) * (double)(1ull << (cc - 1)) * 28.0;
        };
        unsigned best_c = lv.c;
        double best = cost(lv.c);
        for (unsigned k = 2; k <= 4; ++k) {
            if (lv.D % k != 0) continue;
            const unsigned c2 = lv.D / k;
            if (c2 < 6 || c2 >= lv.c || lv.c % c2 != 0) continue;
            const unsigned W2 = (255 + c2 - 1) / c2, S2 = lv.D / c2;
            if ((W2 + S2 - 1) / S2 > lv.L) continue; // the narrower windows would reach past the last level
            if ((double)W2 * (double)n > 600.0 * (double)S2 * (double)(1ull << (c2 - 1))) continue;
            if (cost(c2) * 1.15 < best) {
                best = cost(c2);
                best_c = c2;
            }
        }
        lv.c = best_c;
    }
    const unsigned c = lv.c;
    const unsigned W = (255 + c - 1) / c; // W*c >= 255: the top window absorbs the last carry
    // bucket sets: Sm per scalar vector; a fused batch of `members` vectors over the same bases is ONE pass of every kernel
    // below over members * Sm sets (the digits of vector m go to sets [m Sm, (m + 1) Sm)), split again in k_msm_finish
    const unsigned Sm = lv.L == 1 ? W : lv.D / c;
    const unsigned S = Sm * members;
    const uint32_t B = 1u << (c - 1);
    const size_t G = (size_t)S * B;
    const size_t max_entries = (size_t)W * n * members;
    if (max_entries >= (1ull << 32) || ((size_t)(lv.L - 1) * lv.stride + base + n) >= (1ull << 31)) {
        set_last_error("msm: too many points for 32-bit schedule entries (shard the MSM by point range)");
        return BBG_ERR_ARG;
    }
    int rc;
    constexpr int MAX_PARTS = Context::MSM_MAX_PARTS;
    if ((rc = ws.counts.reserve((G + 1) * 4 + 64 * MAX_PARTS))) return rc;
    if ((rc = ws.offsets.reserve((G + 1) * 4))) return rc;
    if ((rc = ws.cursors.reserve((G + 1) * 4))) return rc;
    if ((rc = ws.sorted.reserve(max_entries * 4))) return rc;
    if ((rc = ws.buckets.reserve(G * sizeof(xyzz_t)))) return rc;
    uint32_t* counts = (uint32_t*)ws.counts.p;
    uint32_t* offsets = (uint32_t*)ws.offsets.p;
    uint32_t* cursors = (uint32_t*)ws.cursors.p;
    uint32_t* sorted = (uint32_t*)ws.sorted.p;
    xyzz_t* buckets = (xyzz_t*)ws.buckets.p;
    uint32_t* pending = counts + G + 1; // 15 merge-level counters (per part) behind the histogram (zeroed with it)

    pr.mark(st, PH_MSM_DIGITS);
    BBG_CUDA(cudaMemsetAsync(counts, 0, (G + 1) * 4 + 64 * MAX_PARTS, st));
    BBG_CUDA(cudaMemsetAsync(buckets, 0, G * sizeof(xyzz_t), st)); // all-zero XYZZ = infinity

    DigitParams dp;
    dp.n = (uint32_t)n;
    dp.c = c;
    dp.W = W;
    dp.S = Sm;
    for (unsigned m = 1; m < 4; ++m) dp.more[m - 1] = (batch != nullptr && m < members) ? batch->scalars[m] : nullptr;
    dp.B = B;
    dp.level_stride = (uint32_t)lv.stride;
    dp.base = (uint32_t)base;
    // windows that hold fewer than ~10 scalar bits (+1 for the signed-digit carry) collide in a handful of buckets
    dp.agg_from = W;
    for (unsigned w = 0; w < W; ++w) {
        if (w * c + 10 >= 254) {
            dp.agg_from = w;
            break;
        }
    }
    dp.agg_from = env_uint("BBG_MSM_AGG_FROM", dp.agg_from);
    const dim3 dig_blocks(div_up(n, 256), members);
    if (batch != nullptr) arrival = nullptr;
    if (arrival != nullptr && arrival->count > 1) {
        // the histogram does not care which scalar a digit came from: count every piece as soon as its copy has landed
        for (size_t k = 0; k < arrival->count; ++k) {
            const size_t lo = k * arrival->piece;
            if (lo >= n) break;
            DigitParams dk = dp;
            dk.n = (uint32_t)std::min(arrival->piece, n - lo);
            BBG_CUDA(cudaStreamWaitEvent(st, arrival->ready[k], 0));
            k_msm_digits<DIG_HISTOGRAM><<<div_up(dk.n, 256), 256, 0, st>>>((const fr_t*)d_scalars + lo, dk, counts, nullptr, nullptr, 0, nullptr);
            ctx->launches += 1;
        }
    } else {
        if (arrival != nullptr && arrival->count == 1) BBG_CUDA(cudaStreamWaitEvent(st, arrival->ready[0], 0));
        k_msm_digits<DIG_HISTOGRAM><<<dig_blocks, 256, 0, st>>>((const fr_t*)d_scalars, dp, counts, nullptr, nullptr, 0, nullptr);
        ctx->launches += 1;
    }
    pr.mark(st, PH_MSM_SCAN);
    if ((rc = exclusive_scan(ctx, ws, counts, G, offsets, cursors, st))) return rc;
    // ---- optional pairwise affine passes (k_msm_pair_pass): J levels, each halving every bucket's entry count.
    // OFF by default: measured on B200 (2^20 points, c = 17) the passes take 3.6-4.2 ms against 2.47 ms for the
    // projective accumulation below -- 5.8 instead of 10 multiplies per addition, but two dependent sweeps over
    // memory per level and a ~70 us inversion per CTA batch leave the IMAD pipe idle.  Kept as a
    // tested alternative (BBG_MSM_PAIR_PASSES = number of levels) for parts where the balance differs.
    unsigned J = env_uint("BBG_MSM_PAIR_PASSES", 0);
    if (J > 12) J = 12;
    if (members > 1) J = 0;
    // optional: the counting sort moves the points themselves (64 B each) so that every later read streams
    const bool materialise = J > 0 && env_uint("BBG_MSM_MATERIALISE", 0) != 0;

    pr.mark(st, PH_MSM_SCATTER);
    affine_t* pts0 = nullptr;
    if (materialise) {
        if ((rc = ws.pts0.reserve(max_entries * sizeof(affine_t)))) return rc;
        pts0 = (affine_t*)ws.pts0.p;
        k_msm_digits<DIG_SCATTER_POINTS><<<dig_blocks, 256, 0, st>>>((const fr_t*)d_scalars, dp, cursors, nullptr,
                                                                      (const affine_t*)d_points, (uint32_t)point_stride, pts0);
    } else {
        k_msm_digits<DIG_SCATTER_INDEX><<<dig_blocks, 256, 0, st>>>((const fr_t*)d_scalars, dp, cursors, sorted, nullptr, 0, nullptr);
    }
    ctx->launches += 1;

    const uint32_t* acc_offsets = offsets;
    const affine_t* acc_points = (const affine_t*)d_points;
    size_t acc_entries = max_entries;
    if (J > 0) {
        if ((rc = ensure_inv_fix(ctx))) return rc;
        if ((rc = ws.lvl_offsets.reserve((size_t)J * (G + 1) * 4))) return rc;
        uint32_t* lvl = (uint32_t*)ws.lvl_offsets.p; // row j-1: offsets of level j
        if ((rc = exclusive_scan(ctx, ws, counts, G, lvl, nullptr, st, J, 1))) return rc;
        // worst-case entry counts per level: sum_b ceil(m_b / 2) <= (E + G) / 2
        size_t e_max[13];
        e_max[0] = max_entries;
        for (unsigned j = 1; j <= J; ++j) e_max[j] = std::min(e_max[j - 1], (e_max[j - 1] + G + 1) / 2);
        if ((rc = ws.pairs_a.reserve(e_max[1] * sizeof(affine_t)))) return rc;
        if (J > 1 && (rc = ws.pairs_b.reserve(e_max[2] * sizeof(affine_t)))) return rc;
        if ((rc = ws.pair_pre.reserve(e_max[1] * 32))) return rc;
        if ((rc = ws.pair_meta.reserve(e_max[1] * 8))) return rc;
        const unsigned k_force = std::min<unsigned>(PAIR_K_MAX, env_uint("BBG_MSM_PAIR_K", 0)); // 0: choose per level
        pr.mark(st, PH_MSM_PAIRS);
        const affine_t* in = pts0;
        const uint32_t* off_in = offsets;
        for (unsigned j = 1; j <= J; ++j) {
            affine_t* out = (affine_t*)((j & 1) ? ws.pairs_a.p : ws.pairs_b.p);
            const uint32_t* off_out = lvl + (size_t)(j - 1) * (G + 1);
            // slots per thread: as many as amortise the inversion, but keep >= ~4 CTAs per SM in the grid
            unsigned K = (unsigned)(e_max[j] / ((size_t)PAIR_THREADS * ctx->num_sms * 4));
            K = k_force ? k_force : std::max(4u, std::min(32u, K));
            const unsigned blocks = div_up(e_max[j], (size_t)PAIR_THREADS * K);
            if (j == 1 && !materialise) {
                k_msm_pair_pass<true><<<blocks, PAIR_THREADS, 0, st>>>(
                    sorted, (const affine_t*)d_points, (uint32_t)point_stride, off_in, off_out, (uint32_t)G, out,
                    (const uint32_t*)ctx->inv_fix_fq, K, (uint4*)ws.pair_pre.p, (uint2*)ws.pair_meta.p);
            } else {
                k_msm_pair_pass<false><<<blocks, PAIR_THREADS, 0, st>>>(nullptr, in, 1, off_in, off_out, (uint32_t)G, out,
                                                                        (const uint32_t*)ctx->inv_fix_fq, K,
                                                                        (uint4*)ws.pair_pre.p, (uint2*)ws.pair_meta.p);
            }
            ctx->launches += 1;
            in = out;
            off_in = off_out;
        }
        acc_offsets = off_in;
        acc_points = in;
        acc_entries = e_max[J];
    }

    // chunking: equal work per thread; about three waves of resident threads, chunk >= 16 entries.  More waves balance the
    // tail of the accumulation better but every chunk boundary costs a slot merge (2^20, ms: waves 1: 2.28 + 0.09 fixup,
    // 2: 2.20 + 0.13, 3: 2.15 + 0.14, 4: 2.14 + 0.18, 6: 2.10 + 0.22).
    // The number of non-zero digits is only known on the device; size for the maximum.
    const size_t resident = (size_t)ctx->num_sms * ACC_THREADS * 4;
    const size_t waves = std::max(1u, env_uint("BBG_MSM_WAVES", 3));
    size_t chunk = (acc_entries + waves * resident - 1) / (waves * resident);
    if (chunk < 16) chunk = 16;
    const size_t num_chunks = (acc_entries + chunk - 1) / chunk;
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
    // ---- parts (experiment, OFF by default: BBG_MSM_PARTS = 2 | 4): a single large MSM over one bucket set is cut into H
    // contiguous bucket ranges.  Each range runs its own accumulate -> slot merge -> bucket reduction chain; the chains of
    // ranges H-1 .. 1 go to high-priority side streams and the chain of range 0 follows on `st`, the idea being that the
    // latency-bound tail of every range but the last runs underneath the accumulation of the next.  Range h's reduction adds
    // (h B / H) * (plain sum of its buckets) for the weights it does not see; k_msm_parts_sum adds the results.
    // Measured on B200 (2^20 points, c = 20): 3.41 ms with 2 parts, 3.65 with 4, against 3.12 ms uncut (2^18: 1.64 / 1.76 /
    // 1.35).  The accumulation's CTAs fill the register file (4 x 128 threads x 128 registers per SM) and each lives for
    // about a third of the kernel, so the tail kernels of the side streams -- priority or not -- only find SM slots at the
    // wave boundaries and end up running beside the LAST range's tail instead of under its accumulation, while the extra
    // launches and the offset doublings are paid in full.  Kept (parity-tested) for parts with a finer-grained block scheduler.
    unsigned H = 1;
    if (allow_parts && J == 0 && S == 1) {
        const unsigned want = env_uint("BBG_MSM_PARTS", 1u);
        while (H * 2 <= want && H * 2 <= (unsigned)MAX_PARTS && (B / (H * 2)) >= (1u << 12)) H *= 2;
    }
    const uint32_t Bp = B / H; // buckets per part (H = 1: all of them; S sets are only cut when S = 1)
    if ((rc = ws.partials.reserve(H * slots_total * sizeof(Slot)))) return rc;

    // bucket reduction of `sets` sets of `nb` buckets: geometry and scratch size
    struct ReducePlan {
        ReduceRows rows;
        uint32_t segs0, count1, segs1, parts;
        unsigned ell0, sh1;
        bool team0;
        size_t total_segs, n_rows, elems;
    };
    auto plan_reduce = [&](uint32_t nb, bool plain) {
        // Level 0 sweeps all buckets (2 additions each, throughput bound once there are many): short segments there give
        // many workers; level 1 sweeps the nb / ell0 segment sums with cooperative teams.
        // Measured on B200, 2^19 buckets (ms of bucket reduction; round 1's shape ell0 = 16, ell1 = 4 with plain additions: 0.65):
        // teams at both levels 16/4: 0.65, 4/16: 0.77; plain level 0 + team level 1: 4/8 0.61, 4/16 0.53, 4/32 0.58,
        // 8/8 0.51, 8/16 0.53, 2/16 0.73.  Small bucket sets (2^15 buckets, c = 16): teams at both levels, 4/4: 0.18 (0.34).
        // A later sweep with arbitrary (non power-of-two) ell0 -- 6, 10, 12 at 2^19 buckets; 3, 5, 6 at 2^17; 2, 3, 6 at 2^15 --
        // and a 128-register / four-CTA build of level 0 moved the whole MSM by less than 1.5 % either way: level 0 is bound by
        // its multiplies (2 full additions per bucket), not by wave quantisation.
        ReducePlan rp;
        const bool big = nb >= (1u << 17);
        // ell0 need not be a power of two (k_msm_finish multiplies by it with a double-and-add): BBG_MSM_ELL0 = any length
        rp.ell0 = 1u << std::min(6u, env_uint("BBG_MSM_ELL0_LOG2", big ? 3u : 2u));
        rp.ell0 = std::max(1u, std::min(64u, env_uint("BBG_MSM_ELL0", rp.ell0)));
        rp.sh1 = std::min(6u, env_uint("BBG_MSM_ELL1_LOG2", big ? 3u : 2u));
        rp.team0 = env_uint("BBG_MSM_SEG0_TEAM", ((size_t)S * (nb / rp.ell0)) < (1u << 16) ? 1u : 0u) != 0;
        memset(&rp.rows, 0, sizeof(rp.rows));
        rp.rows.S = (uint32_t)S;
        rp.segs0 = (nb + rp.ell0 - 1) / rp.ell0;
        rp.count1 = rp.segs0 - 1; // level 1 works on T[1..]
        rp.segs1 = (rp.count1 + (1u << rp.sh1) - 1) >> rp.sh1;
        rp.rows.levels = rp.count1 ? 2 : 1;
        rp.rows.mult[0] = rp.ell0;
        rp.rows.m[0] = rp.segs0;
        rp.rows.off[0] = 0;
        rp.rows.mult[1] = 1u << rp.sh1;
        rp.rows.m[1] = rp.segs1;
        rp.rows.off[1] = (uint32_t)((size_t)rp.segs0 * S);
        rp.total_segs = (size_t)rp.segs0 + rp.segs1;
        // the plain sum of all buckets = the sum of the level-0 segment sums T0: one more row for the tree sum
        rp.rows.m[rp.rows.levels] = rp.segs0;
        rp.rows.off[rp.rows.levels] = (uint32_t)(rp.total_segs * S);
        // parts: every first-stage CTA sums <= ~4 entries per team; second stage: parts / 64 per team
        rp.parts = std::min<uint32_t>(2 * TREE_TEAMS, std::max<uint32_t>(1, rp.segs0 / (4 * TREE_TEAMS)));
        rp.n_rows = (size_t)(rp.rows.levels + (plain ? 1 : 0)) * S;
        // layout: R0 | V1 | T0 | per-part sums | per-row sums
        rp.elems = (rp.total_segs + rp.segs0) * S + rp.n_rows * rp.parts + rp.n_rows;
        return rp;
    };
    const ReducePlan rp = plan_reduce(Bp, H > 1);
    if ((rc = ws.reduce.reserve((rp.elems * H + MAX_PARTS) * sizeof(xyzz_t)))) return rc;
    xyzz_t* part_results = (xyzz_t*)ws.reduce.p + rp.elems * H;

    // one part's chain on stream s; the last step writes XYZZ to out_xyzz (parts) or Jacobian to d_out (H = 1)
    auto run_chain = [&](cudaStream_t s, unsigned h, bool mark) -> int {
        Slot* slots = (Slot*)ws.partials.p + (size_t)h * slots_total;
        uint32_t* pend = pending + 16 * h;
        xyzz_t* bk = buckets + (size_t)h * Bp;
        const uint32_t g_lo = H > 1 ? h * Bp : 0, g_hi = H > 1 ? (h + 1) * Bp : (uint32_t)G;
        if (mark) pr.mark(s, PH_MSM_ACCUMULATE);
        if (J > 0) {
            k_msm_accumulate<true><<<div_up(num_chunks, ACC_THREADS), ACC_THREADS, 0, s>>>(
                nullptr, acc_offsets, (uint32_t)G, g_lo, g_hi, (uint32_t)chunk, (uint32_t)num_chunks, acc_points, 1, buckets, slots, pend);
        } else {
            k_msm_accumulate<false><<<div_up(num_chunks, ACC_THREADS), ACC_THREADS, 0, s>>>(
                sorted, offsets, (uint32_t)G, g_lo, g_hi, (uint32_t)chunk, (uint32_t)num_chunks, (const affine_t*)d_points,
                (uint32_t)point_stride, buckets, slots, pend);
        }
        ctx->launches += 1;
        if (mark) pr.mark(s, PH_MSM_FIXUP);
        {
            size_t n_in = 2 * num_chunks;
            Slot* in = slots;
            unsigned level = 0;
            while (true) {
                const size_t workers = n_in > 1 ? (n_in - 1 + MERGE_K - 1) / MERGE_K : 1;
                Slot* out = in + n_in;
                if (level < MERGE_GRID_LEVELS) {
                    // many workers (level 0 of >= 2^18 points: ~57 k): throughput bound, one thread each; else a team each
                    if (workers >= env_uint("BBG_MSM_MERGE_WIDE_FROM", 32768)) {
                        k_msm_merge<false><<<div_up(workers, 128), 128, 0, s>>>(in, (uint32_t)n_in, (uint32_t)workers, pend + level, buckets, out,
                                                                               pend + level + 1);
                    } else {
                        k_msm_merge<true><<<div_up(workers * 4, 128), 128, 0, s>>>(in, (uint32_t)n_in, (uint32_t)workers, pend + level, buckets, out,
                                                                                  pend + level + 1);
                    }
                    ctx->launches += 1;
                } else {
                    k_msm_merge_rest<<<1, MERGE_REST_THREADS, 0, s>>>(in, (uint32_t)n_in, level, pend, buckets);
                    ctx->launches += 1;
                    break;
                }
                ++level;
                if (workers == 1) break;
                in = out;
                n_in = 2 * workers;
            }
        }
        // bucket reduction (see the comment above k_msm_segments)
        if (mark) pr.mark(s, PH_MSM_REDUCE);
        xyzz_t* r_all = (xyzz_t*)ws.reduce.p + rp.elems * h;
        xyzz_t* t0 = r_all + rp.total_segs * S;
        xyzz_t* part_out = t0 + (size_t)rp.segs0 * S;
        xyzz_t* row_out = part_out + rp.n_rows * rp.parts;
        {
            const uint32_t workers = rp.segs0 * (uint32_t)S;
            if (rp.team0) {
                k_msm_segments<false, true><<<div_up((size_t)workers * 4, SEG_THREADS), SEG_THREADS, 0, s>>>(bk, Bp, Bp, rp.ell0, rp.segs0, workers, r_all, t0);
            } else {
                k_msm_segments<false, false><<<div_up(workers, SEG_THREADS), SEG_THREADS, 0, s>>>(bk, Bp, Bp, rp.ell0, rp.segs0, workers, r_all, t0);
            }
            ctx->launches += 1;
        }
        if (rp.count1) {
            const uint32_t workers = rp.segs1 * (uint32_t)S;
            k_msm_segments<true, true><<<div_up((size_t)workers * 4, SEG_THREADS), SEG_THREADS, 0, s>>>(t0 + 1, rp.segs0, rp.count1, 1u << rp.sh1, rp.segs1,
                                                                                                       workers, r_all + rp.rows.off[1], nullptr);
            ctx->launches += 1;
        }
        ReduceRows tree_rows = rp.rows;
        tree_rows.levels = (uint32_t)(rp.n_rows / S); // + the plain-sum row when the set is cut into parts
        k_msm_tree_sum<<<dim3(rp.parts, (unsigned)rp.n_rows), TREE_THREADS, 0, s>>>(r_all, tree_rows, part_out);
        ctx->launches += 1;
        const xyzz_t* sums = part_out;
        if (rp.parts > 1) {
            ReduceRows second;
            memset(&second, 0, sizeof(second));
            second.S = (uint32_t)rp.n_rows; // every row of the first stage is one "set" of a single level with `parts` entries
            second.levels = 1;
            second.m[0] = rp.parts;
            k_msm_tree_sum<<<dim3(1, (unsigned)rp.n_rows), TREE_THREADS, 0, s>>>(part_out, second, row_out);
            ctx->launches += 1;
            sums = row_out;
        }
        if (mark) pr.mark(s, PH_MSM_COMBINE);
        k_msm_finish<<<1, FINISH_THREADS, S * sizeof(xyzz_t), s>>>(sums, rp.rows, c, (jac_t*)d_out, rp.rows.levels, h * Bp,
                                                                   H > 1 ? part_results + h : nullptr, members);
        ctx->launches += 1;
        return BBG_OK;
    };

    if (H == 1) {
        if ((rc = run_chain(st, 0, true))) return rc;
    } else {
        BBG_CUDA(cudaEventRecord(ctx->ev_part_fork, st));
        for (unsigned h = H - 1; h >= 1; --h) {
            cudaStream_t s = ctx->part_stream[h - 1];
            BBG_CUDA(cudaStreamWaitEvent(s, ctx->ev_part_fork, 0));
            if ((rc = run_chain(s, h, false))) return rc;
            BBG_CUDA(cudaEventRecord(ctx->ev_part_join[h - 1], s));
        }
        if ((rc = run_chain(st, 0, true))) return rc;
        for (unsigned h = 1; h < H; ++h) BBG_CUDA(cudaStreamWaitEvent(st, ctx->ev_part_join[h - 1], 0));
        k_msm_parts_sum<<<1, 32, 0, st>>>(part_results, H, (jac_t*)d_out);
        ctx->launches += 1;
    }
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
