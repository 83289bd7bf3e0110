// msm.cu -- BN254 G1 multi-scalar multiplication (Pippenger bucket method) for sm_100a.
//
// Replaces barretenberg's scalar_multiplication::pippenger / pippenger_unsafe
// (bb/ecc/curves/bn254/scalar_multiplication/scalar_multiplication.cpp:853-929).  Same inputs
// (Montgomery fr scalars, Montgomery affine points), same group element out (96-byte Jacobian).
// The CPU reference's pipeline (compute_wnaf_states :188-252 -> organize_buckets :260-271 ->
// reduce_buckets/add_affine_points :305-521 -> evaluate_pippenger_rounds :720-838) is re-thought
// for the GPU rather than translated:
//
//   1. k_msm_digits<false>   scalar -> canonical -> signed c-bit windows, histogram of bucket sizes
//   2. scan                  bucket sizes -> bucket start offsets (one flat key space: window*B + |digit|-1)
//   3. k_msm_digits<true>    counting-sort scatter of (point index, sign) by bucket   [order inside a
//                            bucket is irrelevant: the group is commutative]
//   4. k_msm_accumulate      the sorted list is cut into equal-length chunks, one per thread, so
//                            every thread does the same number of mixed additions no matter how
//                            skewed the scalar distribution is; buckets cut by a chunk boundary leave
//                            partial sums that k_msm_fixup stitches together
//   5. k_msm_reduce_level    sum_b (b+1)*bucket[b] per window by a log_16-depth hierarchy of
//                            running sums (the reference's running-sum trick :773-783, parallelised)
//   6. k_msm_combine         Horner over the windows (c doublings each), XYZZ -> Jacobian
//
// Arithmetic is integer-ALU bound (IMAD.WIDE); HBM traffic is ~(32 + 64*W) B per point.
#include "g1.cuh"
#include "internal.hpp"

namespace bbg {

struct MsmPlan {
    unsigned c;      // window bits
    unsigned W;      // number of windows, W*c >= 255 so the top window never carries out
    unsigned B;      // buckets per window = 2^(c-1) (signed digits)
    size_t G;        // total buckets W*B
};

static MsmPlan msm_plan(size_t n)
{
    unsigned lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) {
        ++lg;
    }
    int c = (int)lg - 4;
    if (c < 2) c = 2;
    if (c > 20) c = 20;
    MsmPlan p;
    p.c = (unsigned)c;
    p.W = (255 + p.c - 1) / p.c;
    p.B = 1u << (p.c - 1);
    p.G = (size_t)p.W * p.B;
    return p;
}

// ------------------------------------------------------------------------------------------------
// 1/3. digits: histogram (SCATTER = false) or counting-sort scatter (SCATTER = true)
// ------------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_msm_digits(const fr_t* __restrict__ scalars,
                                                    uint32_t n,
                                                    unsigned c,
                                                    unsigned W,
                                                    uint32_t* __restrict__ counters,
                                                    uint32_t* __restrict__ sorted)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    // from_montgomery_form => canonical integer in [0, r)  (scalar_multiplication.cpp:224)
    fr_t s = fe_from_mont(fe_load_nc<FrParams>(scalars + i));
    uint32_t k[9];
#pragma unroll
    for (int j = 0; j < 8; ++j) k[j] = s.l[j];
    k[8] = 0;
    const uint32_t B = 1u << (c - 1);
    const uint32_t mask = (1u << c) - 1;
    uint32_t carry = 0;
    for (unsigned w = 0; w < W; ++w) {
        unsigned pos = w * c;
        unsigned limb = pos >> 5, off = pos & 31;
        uint64_t two = (uint64_t)k[limb] | ((uint64_t)k[limb + 1] << 32);
        uint32_t v = ((uint32_t)(two >> off) & mask) + carry;
        // signed digit in (-B, B]: v > B  =>  digit v - 2^c, borrow one from the next window
        uint32_t neg = v > B ? 1u : 0u;
        uint32_t mag = neg ? (mask + 1 - v) : v;
        carry = neg;
        if (mag) {
            uint32_t g = w * B + (mag - 1);
            if (SCATTER) {
                uint32_t dst = atomicAdd(&counters[g], 1u);
                sorted[dst] = (i << 1) | neg;
            } else {
                atomicAdd(&counters[g], 1u);
            }
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
struct alignas(16) Partial {
    xyzz_t acc;
    int32_t bucket; // -1: none
    int32_t pad[3];
};

static constexpr int ACC_THREADS = 128;

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
                                                                   const affine_t* __restrict__ points,
                                                                   uint32_t point_stride,
                                                                   xyzz_t* __restrict__ buckets,
                                                                   Partial* __restrict__ heads,
                                                                   Partial* __restrict__ tails)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = __ldg(offsets + G); // number of non-zero digits, produced by the scan
    const uint64_t start64 = (uint64_t)t * chunk;
    if (start64 >= total) {
        return;
    }
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (start64 + chunk < total) ? start + chunk : total;

    heads[t].bucket = -1;
    tails[t].bucket = -1;

    // the (non-empty) bucket that contains position `start`
    uint32_t g = upper_bound_u32(offsets, G + 1, start) - 1;
    uint32_t g_begin = __ldg(offsets + g);
    uint32_t g_end = __ldg(offsets + g + 1);

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
            // flush
            const bool complete = (g_begin >= start) && (g_end <= end);
            if (complete) {
                xyzz_store(buckets + g, acc);
            } else if (g_begin < start) {
                heads[t].acc = acc; // bucket began in an earlier chunk
                heads[t].bucket = (int32_t)g;
            } else {
                tails[t].acc = acc; // bucket continues into a later chunk
                tails[t].bucket = (int32_t)g;
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
}

// one thread per chunk whose tail opens a cut bucket: add the heads of the following chunks
__global__ void __launch_bounds__(128) k_msm_fixup(uint32_t num_chunks,
                                                    const Partial* __restrict__ heads,
                                                    const Partial* __restrict__ tails,
                                                    xyzz_t* __restrict__ buckets)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_chunks) {
        return;
    }
    const int32_t g = tails[t].bucket;
    if (g < 0) {
        return;
    }
    xyzz_t acc = xyzz_load(&tails[t].acc);
    for (uint32_t u = t + 1; u < num_chunks && heads[u].bucket == g; ++u) {
        xyzz_t h = xyzz_load(&heads[u].acc);
        xyzz_add(acc, h);
    }
    xyzz_store(buckets + g, acc);
}

// ------------------------------------------------------------------------------------------------
// 5. bucket reduction.  Invariant per window:  sum_i i * X0[i]  ==  sum_i ( scale_k * i * R_k[i] + C_k[i] )
//    with scale_k = prod of the segment lengths of the levels below.  One thread folds `ell` entries.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_msm_reduce_level(const xyzz_t* __restrict__ r_in,
                                                          const xyzz_t* __restrict__ c_in, // may be null (level 0)
                                                          uint32_t m_in,                   // entries per window
                                                          uint32_t ell,                    // segment length (power of 2)
                                                          uint32_t log_scale,              // log2(scale_k)
                                                          uint32_t num_windows,
                                                          xyzz_t* __restrict__ r_out,
                                                          xyzz_t* __restrict__ c_out)
{
    const uint32_t m_out = m_in / ell;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m_out * num_windows) {
        return;
    }
    const uint32_t w = t / m_out, s = t % m_out;
    const xyzz_t* r = r_in + (size_t)w * m_in + (size_t)s * ell;
    xyzz_t run = xyzz_infinity(), acc = xyzz_infinity();
    for (uint32_t j = ell - 1; j >= 1; --j) {
        xyzz_t x = xyzz_load(r + j);
        xyzz_add(run, x);
        xyzz_add(acc, run);
    }
    {
        xyzz_t x = xyzz_load(r);
        xyzz_add(run, x);
    }
    for (uint32_t d = 0; d < log_scale; ++d) {
        acc = xyzz_dbl(acc);
    }
    if (c_in != nullptr) {
        const xyzz_t* cc = c_in + (size_t)w * m_in + (size_t)s * ell;
        for (uint32_t j = 0; j < ell; ++j) {
            xyzz_t x = xyzz_load(cc + j);
            xyzz_add(acc, x);
        }
    }
    xyzz_store(r_out + (size_t)w * m_out + s, run);
    xyzz_store(c_out + (size_t)w * m_out + s, acc);
}

// 6. result = sum_w 2^(c w) * (C_w + R_w)   [bucket b (0-based) carries weight b+1 = i + 1]
__global__ void k_msm_combine(const xyzz_t* __restrict__ r_top,
                              const xyzz_t* __restrict__ c_top,
                              uint32_t num_windows,
                              uint32_t c,
                              jac_t* __restrict__ out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) {
        return;
    }
    xyzz_t acc = xyzz_infinity();
    for (int w = (int)num_windows - 1; w >= 0; --w) {
        for (uint32_t d = 0; d < c; ++d) {
            acc = xyzz_dbl(acc);
        }
        xyzz_t s = xyzz_load(c_top + w);
        xyzz_t r = xyzz_load(r_top + w);
        xyzz_add(s, r);
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
__device__ __forceinline__ xyzz_t xyzz_shfl_down(const xyzz_t& p, int delta)
{
    xyzz_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, p.x.l[i], delta);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, p.y.l[i], delta);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, p.zz.l[i], delta);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, p.zzz.l[i], delta);
    }
    return r;
}
__global__ void __launch_bounds__(32) k_g1_sum(const jac_t* __restrict__ in, uint32_t n, jac_t* __restrict__ out)
{
    const unsigned lane = threadIdx.x;
    xyzz_t acc = xyzz_infinity();
    for (uint32_t i = lane; i < n; i += 32) {
        jac_t j;
        j.x = fe_load<FqParams>(&in[i].x);
        j.y = fe_load<FqParams>(&in[i].y);
        j.z = fe_load<FqParams>(&in[i].z);
        xyzz_t p = xyzz_from_jacobian(j);
        xyzz_add(acc, p);
    }
    for (int d = 16; d >= 1; d >>= 1) {
        xyzz_t o = xyzz_shfl_down(acc, d);
        xyzz_add(acc, o);
    }
    if (lane == 0) {
        jac_t j = xyzz_to_jacobian(acc);
        fe_store(&out->x, j.x);
        fe_store(&out->y, j.y);
        fe_store(&out->z, j.z);
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

// Device-pointer MSM: d_scalars (n x 32 B), d_points (affine, stride in points), d_out (96 B Jacobian).
int msm_device(Context* ctx, const void* d_scalars, size_t n, const void* d_points, size_t point_stride, void* d_out, cudaStream_t st)
{
    if (n == 0) {
        k_set_infinity<<<1, 1, 0, st>>>((jac_t*)d_out);
        ctx->launches += 1;
        BBG_CUDA(cudaGetLastError());
        return BBG_OK;
    }
    if (n >= (1ull << 31)) {
        set_last_error("msm: n must be < 2^31");
        return BBG_ERR_ARG;
    }
    const MsmPlan pl = msm_plan(n);
    const size_t max_entries = (size_t)pl.W * n;
    if (max_entries >= (1ull << 32)) {
        set_last_error("msm: W*n must be < 2^32 (shard the MSM by point range)");
        return BBG_ERR_ARG;
    }
    int rc;
    if ((rc = ctx->msm_counts.reserve((pl.G + 1) * 4))) return rc;
    if ((rc = ctx->msm_offsets.reserve((pl.G + 1) * 4))) return rc;
    if ((rc = ctx->msm_cursors.reserve((pl.G + 1) * 4))) return rc;
    if ((rc = ctx->msm_sorted.reserve(max_entries * 4))) return rc;
    if ((rc = ctx->msm_buckets.reserve(pl.G * sizeof(xyzz_t)))) return rc;
    uint32_t* counts = (uint32_t*)ctx->msm_counts.p;
    uint32_t* offsets = (uint32_t*)ctx->msm_offsets.p;
    uint32_t* cursors = (uint32_t*)ctx->msm_cursors.p;
    uint32_t* sorted = (uint32_t*)ctx->msm_sorted.p;
    xyzz_t* buckets = (xyzz_t*)ctx->msm_buckets.p;

    BBG_CUDA(cudaMemsetAsync(counts, 0, (pl.G + 1) * 4, st));
    BBG_CUDA(cudaMemsetAsync(buckets, 0, pl.G * sizeof(xyzz_t), st)); // all-zero XYZZ = infinity

    const unsigned dig_blocks = div_up(n, 256);
    k_msm_digits<false><<<dig_blocks, 256, 0, st>>>((const fr_t*)d_scalars, (uint32_t)n, pl.c, pl.W, counts, nullptr);
    ctx->launches += 1;
    if ((rc = exclusive_scan(ctx, counts, pl.G, offsets, cursors, st))) return rc;
    k_msm_digits<true><<<dig_blocks, 256, 0, st>>>((const fr_t*)d_scalars, (uint32_t)n, pl.c, pl.W, cursors, sorted);
    ctx->launches += 1;

    // chunking: equal work per thread; about two waves of resident threads, chunk >= 16 entries.
    // The number of non-zero digits is only known on the device; size for the maximum.
    const size_t resident = (size_t)ctx->num_sms * ACC_THREADS * 4;
    size_t chunk = (max_entries + 2 * resident - 1) / (2 * resident);
    if (chunk < 16) chunk = 16;
    const size_t num_chunks = (max_entries + chunk - 1) / chunk;
    if ((rc = ctx->msm_partials.reserve(2 * num_chunks * sizeof(Partial)))) return rc;
    Partial* heads = (Partial*)ctx->msm_partials.p;
    Partial* tails = heads + num_chunks;
    // chunks past the real total never run their body: mark every partial "none" first
    BBG_CUDA(cudaMemsetAsync(heads, 0xff, 2 * num_chunks * sizeof(Partial), st));

    // the real total lives in offsets[G]; the kernel reads it from there (no host round trip)
    k_msm_accumulate<<<div_up(num_chunks, ACC_THREADS), ACC_THREADS, 0, st>>>(
        sorted, offsets, (uint32_t)pl.G, (uint32_t)chunk, (const affine_t*)d_points, (uint32_t)point_stride, buckets,
        heads, tails);
    k_msm_fixup<<<div_up(num_chunks, 128), 128, 0, st>>>((uint32_t)num_chunks, heads, tails, buckets);
    ctx->launches += 2;

    // bucket reduction hierarchy
    size_t level_elems = 0;
    {
        uint32_t m = pl.B;
        while (m > 1) {
            uint32_t ell = m >= 16 ? 16 : m;
            m /= ell;
            level_elems += (size_t)m * pl.W;
        }
        if (pl.B == 1) level_elems = pl.W;
    }
    if ((rc = ctx->msm_reduce.reserve(2 * (level_elems + pl.W) * sizeof(xyzz_t)))) return rc;
    xyzz_t* pool = (xyzz_t*)ctx->msm_reduce.p;
    const xyzz_t* r_in = buckets;
    const xyzz_t* c_in = nullptr;
    uint32_t m = pl.B, log_scale = 0;
    if (m == 1) {
        // c == 1 never happens (c >= 2), kept for completeness
        set_last_error("msm: unsupported window");
        return BBG_ERR_ARG;
    }
    while (m > 1) {
        uint32_t ell = m >= 16 ? 16 : m;
        uint32_t m_out = m / ell;
        xyzz_t* r_out = pool;
        xyzz_t* c_out = pool + (size_t)m_out * pl.W;
        pool += 2 * (size_t)m_out * pl.W;
        unsigned threads = m_out * pl.W;
        k_msm_reduce_level<<<div_up(threads, 128), 128, 0, st>>>(r_in, c_in, m, ell, log_scale, pl.W, r_out, c_out);
        ctx->launches += 1;
        r_in = r_out;
        c_in = c_out;
        unsigned lg = 0;
        while ((1u << lg) < ell) ++lg;
        log_scale += lg;
        m = m_out;
    }
    k_msm_combine<<<1, 32, 0, st>>>(r_in, c_in, pl.W, pl.c, (jac_t*)d_out);
    ctx->launches += 1;
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
