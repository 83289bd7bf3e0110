// ntt.cu -- radix-2 NTT / iNTT / coset-FFT over BN254 fr for sm_100a.
//
// Replaces barretenberg's polynomial_arithmetic::{fft, ifft, coset_fft, coset_ifft, *_with_constant,
// coset_fft_with_generator_shift} (bb/polynomials/polynomial_arithmetic.cpp:374-484).  The reference
// runs log2(n) radix-2 DIT passes over the whole array with per-round twiddle tables
// (fft_inner_parallel :140-255).  Here the transform is factored N = N_1 * ... * N_P (P = 2..4,
// N_p <= 256) in the four-step / Cooley-Tukey sense, one kernel launch per factor:
//
//   pass p:  for every fixed (o_1..o_{p-1}, i_{p+1}..i_P): an N_p-point DIF NTT over digit i_p held in
//            registers (radix-8 butterflies, 8 elements per thread) with shared-memory exchanges between
//            radix-8 rounds, then the inter-pass twiddle w_N^(N_1..N_{p-1} * o_p * rest) fused into the
//            store.  A CTA tile is N_p rows x 8 adjacent columns, so every global access is a 256-byte
//            contiguous run of 32-byte field elements (128-bit loads/stores).
//   last  :  reads rows contiguously, writes the digit-reversed (= natural) index so the output is in
//            natural order like the reference's, with the 1/n, constant and coset g^-i scalings fused.
//   first :  the coset pre-scaling g^i (scale_by_generator :97-117, i < generator_size only) fused
//            into the load.
//
// HBM traffic: P reads + P writes of the array (+ one 32-byte twiddle per element per inner pass);
// arithmetic: (log2 N)/2 + P - 1 (+ scalings) Montgomery multiplies per element -- the kernel is
// bound by the integer pipe (IMAD.WIDE), not by HBM; see DESIGN.md.
#include <cuda/barrier>
#include <cuda/ptx>

#include "ctx.cuh"
#include "field.cuh"
#include "internal.hpp"
#include "poly.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace bbg {

using fr = Fe<FrParams>;

// The pass kernel is a template over LOGE: every thread holds E = 2^LOGE elements (radix-E rounds) and a tile is
// N_p rows x E adjacent columns.  E = 8: 128 registers, 2 CTAs / SM (4 warps per scheduler); E = 4: <= 80 registers,
// 3 CTAs / SM (6 warps per scheduler) at the price of one more shared-memory exchange per pass.  The passes are
// bound by dependent IMAD.WIDE chains ("stall_wait" is the top stall reason in the ncu capture), so the extra
// warps are what raises the pipe utilisation.
static constexpr int NTT_THREADS = 256;
template <int LOGE> struct NttGeom {
    static constexpr int E = 1 << LOGE;                         // elements per thread = columns per tile
    // E = 4 uses an XOR-swizzled layout with no padding (sm_idx below); the others pad every row by one 16-byte unit
    static constexpr int PAD = LOGE == 2 ? E : E + 1;           // row pitch in 16-byte units
    static constexpr int HALF_STRIDE = NTT_THREADS * PAD;       // 16-byte units between the two halves of an element
    static constexpr int SMEM_BYTES = 2 * HALF_STRIDE * 16;     // E = 8: 73,728 B; E = 4: 32,768 B; E = 2: 24,576 B
    static constexpr int MIN_CTAS = LOGE == 3 ? 2 : (LOGE == 2 ? 3 : 4);
};

// Shared-memory slot (in 16-byte units) of element (row, col) of tile `tile_local` (R rows per tile).
// A 128-bit shared access is served a quarter-warp at a time: 8 lanes x 16 B = all 32 banks once if the 8 lanes hit 8
// different 16-byte bank groups (slot mod 8).  With E = 4 columns a quarter-warp touches either (a) two rows that differ
// in exactly ONE row bit x four columns (every radix-round exchange and the register load after the staging loop) or
// (b) eight consecutive rows of one column (the transposing store of the last pass).  Row pitch E + 1 (round 1)
// serves (b) but not (a): ncu showed 6.7-8 wavefronts per request where 4 is the floor.  The layout below serves both:
//   two rows share a 128-byte line; bank group = parity(row) * 4 + (col ^ ((row >> 1) & 3))
// (a): one differing bit flips the parity -> the two rows sit in different halves, columns stay distinct;
// (b): (row >> 1) & 3 walks the four column slots and each slot's two rows (row, row + 1) have opposite parity.
template <int LOGE> __device__ __forceinline__ uint32_t sm_idx(uint32_t tile_local, uint32_t R, uint32_t row, uint32_t col)
{
    if constexpr (LOGE == 2) {
        const uint32_t line = (tile_local * R + row) >> 1; // R is even: a pair of rows never straddles two tiles
        return (line << 3) | ((__popc(row) & 1u) << 2) | (col ^ ((row >> 1) & 3u));
    } else {
        return (tile_local * R + row) * (uint32_t)NttGeom<LOGE>::PAD + col;
    }
}
static constexpr int NTT_MAX_PASSES = 4;
static constexpr unsigned NTT_DEFAULT_LOGE = 2; // measured at 2^22: fft 0.892 ms with E = 4 vs 0.948 ms with E = 8 (BBG_NTT_LOGE=3)

struct PassParams {
    const fr* src;
    fr* dst;
    const fr* tw_big;   // w_N^e, e in [0, N)
    const fr* stage_tw; // w_{2^g}^j (or inverse), j in [0, 2^(g-1))
    uint32_t log_n;
    uint32_t g;       // this pass transforms a digit of g bits
    uint32_t below;   // bits below the digit
    uint32_t above;   // bits above the digit
    uint32_t last;
    uint32_t inverse;
    // last pass: where each higher digit lands in the natural-order output index
    uint32_t g1;               // bits of the first digit (the tile's columns)
    uint32_t num_mid;          // digits 2..P-1
    uint32_t mid_bits[2], mid_src_shift[2], mid_dst_shift[2];
    // fused pre-scaling (first pass): x[i] *= pro_hi[i >> split] * pro_lo[i & mask], i < pro_size
    const fr* pro_lo;
    const fr* pro_hi;
    const fr* pro_full; // cached full table start * shift^i (one multiply instead of two); overrides pro_lo / pro_hi
    uint32_t pro_split;
    uint64_t pro_size;
    // fused post-scaling (last pass): mode 0 none, 1 constant, 2 epi_hi[o >> split] * epi_lo[o & mask], 3 epi_full[o]
    uint32_t epi_mode;
    const fr* epi_full;
    // non-last pass: inter-pass twiddles taken from a table already multiplied by the final constant (1/n of an
    // ifft), indexed by the forward exponent; every element is multiplied (entry 0 carries the constant)
    const fr* tw_scaled;
    const fr* epi_lo;
    const fr* epi_hi;
    uint32_t epi_split;
    fr epi_const;
    // output interleave for the extended coset FFT: natural index o lands at (o << out_shift) + out_off
    uint32_t out_shift, out_off;
    // ---- multi-GPU four-step (one process per GPU; see ntt_device): this rank holds the sub-array whose index bits
    // [rk_pos, rk_pos + rk_bits) equal rk_val, packed.  log_n / below / g1 describe the LOCAL array; twiddles and
    // scalings use the true index, rebuilt by inserting the rank bits.
    uint32_t tw_log_n;       // log2 of the full transform
    uint32_t rk_bits, rk_pos, rk_val;
    uint32_t mid_bits_total; // last pass: total width of digits 2..P-1
    uint32_t split_low;      // last pass, after the all-to-all: a row is rk_bits source-rank chunks of 2^split_low elements,
    uint32_t chunk_log;      //   chunk s starting at s << chunk_log
    // Fused exchange (multi-GPU, the pass before the last): instead of writing the intermediate locally and handing it
    // to an NCCL all-to-all, every element is stored straight into the receive buffer of the rank that owns its chunk,
    // over NVLink peer memory: element a of the local intermediate belongs to chunk a >> chunk_log, and lands at slot
    // (rk_val << chunk_log) + (a & chunk mask) of that rank's buffer -- exactly where the all-to-all would have put it.
    uint32_t peer_on;
    fr* peer_dst[8];
    // Natural block distribution over peer memory (rank q holds elements [q << block_log, (q + 1) << block_log) of the
    // input / output array): the FIRST pass loads element `at` (true coefficient index) from peer_src[at >> block_log],
    // the LAST pass stores output o to peer_out[o >> block_log] -- no re-distribution step on either side.
    uint32_t src_on, out_on, block_log;
    const fr* peer_src[8];
    fr* peer_out[8];
};

__device__ __forceinline__ uint32_t bitrev(uint32_t x, uint32_t bits) { return __brev(x) >> (32 - bits); }
__device__ __forceinline__ uint64_t insert_bits(uint64_t x, uint32_t pos, uint32_t bits, uint64_t val)
{
    return ((x >> pos) << (pos + bits)) | (val << pos) | (x & ((1ull << pos) - 1));
}
__device__ __forceinline__ uint64_t squeeze_bits(uint64_t x, uint32_t pos, uint32_t bits)
{
    return ((x >> (pos + bits)) << pos) | (x & ((1ull << pos) - 1));
}

__device__ __forceinline__ void smem_store(uint4* sm, uint32_t half_stride, uint32_t idx, const fr& v)
{
    sm[idx] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    sm[half_stride + idx] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ fr smem_load(const uint4* sm, uint32_t half_stride, uint32_t idx)
{
    uint4 a = sm[idx], b = sm[half_stride + idx];
    fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// DIF butterfly: (u, v) -> (u + v, (u - v) * w)
__device__ __forceinline__ void bfly(fr& u, fr& v, const fr& w)
{
    fr s = fe_add(u, v);
    fr d = fe_sub_lazy(u, v); // (0, 4p): legal because w comes from a canonical table (k_powers stores reduce_once'd values)
    u = s;
    v = fe_mul(d, w);
}
__device__ __forceinline__ void bfly_notw(fr& u, fr& v)
{
    fr s = fe_add(u, v);
    v = fe_sub(u, v);
    u = s;
}

// LOGE (or fewer) DIF stages on the E registers x[j], j = LOGE-bit value of row bits [w0, w0 + LOGE).
// Stages run for row bits b_hi, b_hi-1, ..., w0 (b_hi <= w0 + LOGE - 1). lo = row bits below w0.
// Stage s (row bit w0 + s) pairs j with j + 2^s; its twiddle exponent is ((j mod 2^s) << w0 | lo) << (g-1-w0-s).
// SMEM_TW: the stage twiddles were staged into shared memory by TMA (cp.async.bulk) at kernel start; otherwise they are read
// from global memory through the read-only path.
template <bool SMEM_TW> __device__ __forceinline__ fr load_tw(const fr* p)
{
    if constexpr (SMEM_TW) {
        return fe_load<FrParams>(p);
    } else {
        return fe_load_nc<FrParams>(p);
    }
}
template <int LOGE, bool SMEM_TW = false>
__device__ __forceinline__ void radix_round(fr (&x)[1 << LOGE], uint32_t g, uint32_t w0, uint32_t b_hi, uint32_t lo,
                                            const fr* __restrict__ stage_tw)
{
    // In the last round (w0 == 0, hence lo == 0 for every thread) the exponent is zero exactly when
    // j mod 2^s == 0: those butterflies skip the multiply.  The test is warp-uniform, so nothing diverges.
    const bool tail = (w0 == 0);
#pragma unroll
    for (int s = LOGE - 1; s >= 0; --s) {
        if (b_hi >= w0 + (uint32_t)s) {
            const uint32_t sh = g - 1 - w0 - (uint32_t)s;
#pragma unroll
            for (int v = 0; v < (1 << s); ++v) {
                if (tail && v == 0) {
#pragma unroll
                    for (int u = 0; u < (1 << (LOGE - 1 - s)); ++u) {
                        const int j = (u << (s + 1)) | v;
                        bfly_notw(x[j], x[j + (1 << s)]);
                    }
                } else {
                    const uint32_t e = (((uint32_t)v << w0) | lo) << sh;
                    fr w = load_tw<SMEM_TW>(stage_tw + e);
#pragma unroll
                    for (int u = 0; u < (1 << (LOGE - 1 - s)); ++u) {
                        const int j = (u << (s + 1)) | v;
                        bfly(x[j], x[j + (1 << s)], w);
                    }
                }
            }
        }
    }
}

// Three (or fewer) DIF stages on the 8 registers x[j], j = 3-bit value of row bits [w0, w0+3).
// Stages run for row bits b_hi, b_hi-1, ..., w0 (b_hi <= w0 + 2). lo = row bits below w0.
__device__ __forceinline__ void radix8_round(fr (&x)[8], uint32_t g, uint32_t w0, uint32_t b_hi, uint32_t lo, const fr* __restrict__ stage_tw)
{
    // In the last round (w0 == 0, hence lo == 0 for every thread) the exponent is zero exactly when
    // j == 0: those butterflies skip the multiply.  The test is warp-uniform, so nothing diverges.
    const bool tail = (w0 == 0);
    if (b_hi >= w0 + 2) {
        // bit w0+2: pairs (j, j+4); exponent ((j&3) << w0 | lo) << (g-1-(w0+2))
        const uint32_t sh = g - 3 - w0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (tail && j == 0) {
                bfly_notw(x[j], x[j + 4]);
            } else {
                const uint32_t e = (((uint32_t)j << w0) | lo) << sh;
                fr w = fe_load_nc<FrParams>(stage_tw + e);
                bfly(x[j], x[j + 4], w);
            }
        }
    }
    if (b_hi >= w0 + 1) {
        const uint32_t sh = g - 2 - w0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (tail && j == 0) {
                bfly_notw(x[j], x[j + 2]);
                bfly_notw(x[j + 4], x[j + 6]);
            } else {
                const uint32_t e = (((uint32_t)j << w0) | lo) << sh;
                fr w = fe_load_nc<FrParams>(stage_tw + e);
                bfly(x[j], x[j + 2], w);
                bfly(x[j + 4], x[j + 6], w);
            }
        }
    }
    {
        if (tail) {
            bfly_notw(x[0], x[1]);
            bfly_notw(x[2], x[3]);
            bfly_notw(x[4], x[5]);
            bfly_notw(x[6], x[7]);
        } else {
            const uint32_t e = lo << (g - 1 - w0);
            fr w = fe_load_nc<FrParams>(stage_tw + e);
            bfly(x[0], x[1], w);
            bfly(x[2], x[3], w);
            bfly(x[4], x[5], w);
            bfly(x[6], x[7], w);
        }
    }
}

template <> __device__ __forceinline__ void radix_round<3, false>(fr (&x)[8], uint32_t g, uint32_t w0, uint32_t b_hi, uint32_t lo,
                                                                   const fr* __restrict__ stage_tw)
{
    radix8_round(x, g, w0, b_hi, lo, stage_tw); // hand-unrolled form: ptxas spills less with it than with the generic loops
}

// LAST is a template parameter so that each flavour only carries its own index maths in registers
// TMA_TW: one thread issues a TMA bulk copy (cp.async.bulk, completion on an mbarrier) of this pass's stage-twiddle table
// (2^(g-1) entries, <= 4 KB) into shared memory behind the exchange buffer; the copy flies while the threads load their
// elements from HBM, and the radix rounds then read twiddles from shared memory instead of the L1 / read-only path.
// PERSIST (experiment, OFF by default: BBG_NTT_PERSIST=1): the grid is one wave of resident CTAs and every CTA walks over its
// tile groups; the elements of the NEXT group are fetched with cp.async (LDGSTS, 16 bytes per request, no registers) into a
// staging area behind the exchange buffer while the current group is in its radix rounds, the idea being that a CTA never
// sits in a load phase with nothing to multiply.  Measured on B200 (fft, ms, one-group-per-CTA -> persistent): 2^20 0.254 ->
// 0.298, 2^22 0.906 -> 1.017, 2^24 3.55 -> 4.06, 2^26 16.2 -> 18.7: 12-15 % SLOWER.  The hardware's own CTA turnover already
// overlaps one CTA's loads with its neighbours' arithmetic, and the persistent form pays two more block barriers per group,
// 12-32 more spilled bytes and an extra trip through shared memory.  Kept, parity-tested, as the record of the experiment.
//   !LAST: a thread stages exactly the elements it will own (slot = (j, half) * threads + thread: conflict-free, no barrier);
//   LAST : rows are contiguous in memory, lanes run along the rows and the staging area has the exchange layout (the
//          transposition the non-persistent form does through registers).
static constexpr int NTT_TW_SMEM_BYTES = 128 * 32;
static constexpr int NTT_PERSIST_SMEM_BYTES = 2 * NttGeom<2>::SMEM_BYTES + NTT_TW_SMEM_BYTES; // exchange + staging + twiddles = 69,632 B
static constexpr bool NTT_PERSIST_DEFAULT = false; // BBG_NTT_PERSIST overrides
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int LOGE, bool LAST, int CTAS = NttGeom<LOGE>::MIN_CTAS, bool TMA_TW = false, bool PERSIST = false>
__global__ void __launch_bounds__(NTT_THREADS, CTAS) k_ntt_pass(const PassParams P)
{
    extern __shared__ uint4 sm[];
    constexpr int E = NttGeom<LOGE>::E;
    const uint32_t half_stride = NttGeom<LOGE>::HALF_STRIDE;
    uint4* const stage = sm + 2 * NttGeom<LOGE>::HALF_STRIDE + (TMA_TW ? NTT_TW_SMEM_BYTES / 16 : 0); // PERSIST only
    using tw_barrier = cuda::barrier<cuda::thread_scope_block>;
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ tw_barrier tw_bar;
    const fr* stage_tw = P.stage_tw;
    tw_barrier::arrival_token tw_token;
    if constexpr (TMA_TW) {
        fr* tw_sm = reinterpret_cast<fr*>(sm + 2 * NttGeom<LOGE>::HALF_STRIDE);
        const uint32_t tw_bytes = 32u << (P.g - 1);
        if (threadIdx.x == 0) {
            init(&tw_bar, NTT_THREADS);
            cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            cuda::device::memcpy_async_tx(tw_sm, P.stage_tw, cuda::aligned_size_t<16>(tw_bytes), tw_bar);
            tw_token = cuda::device::barrier_arrive_tx(tw_bar, 1, tw_bytes);
        } else {
            tw_token = tw_bar.arrive();
        }
        stage_tw = tw_sm;
    }

    const uint32_t g = P.g;
    const uint32_t R = 1u << g;                   // rows per tile
    const uint32_t tiles_per_cta = NTT_THREADS >> g; // R threads per tile (g <= 8)
    const uint32_t tile_local = threadIdx.x >> g;
    const uint32_t tau = threadIdx.x & (R - 1);
    const uint32_t col = tau & (E - 1);
    const uint32_t q = tau >> LOGE;               // [0, R/E)
    const uint32_t Nmask = (1u << P.tw_log_n) - 1u;         // full transform (twiddle exponents); n <= 2^28: 32-bit index maths throughout
    const uint32_t num_tiles = (1u << P.log_n) >> (g + LOGE); // local array
    const uint32_t rows8 = R >> LOGE;
    const uint32_t num_groups = (num_tiles + tiles_per_cta - 1) / tiles_per_cta;

    // source element index of item (it, thread) of the last pass's row-contiguous load
    auto last_src_index = [&](uint32_t rest0, uint32_t mid, uint32_t c, uint32_t row) -> uint32_t {
        const uint32_t hi_idx = ((rest0 + c) << P.mid_bits_total) | mid;
        if (P.rk_bits == 0) {
            return (hi_idx << g) + row;
        }
        // after the all-to-all the row is split by source rank: chunk s holds i_P = (s, low)
        const uint32_t src_rank = row >> P.split_low, low = row & ((1u << P.split_low) - 1);
        return (src_rank << P.chunk_log) + (hi_idx << P.split_low) + low;
    };
    auto tile_coords = [&](uint32_t tile, uint32_t& in_base, uint32_t& rest0, uint32_t& mid) {
        in_base = 0; // element index of (row 0, col 0)
        mid = 0;
        if constexpr (!LAST) {
            const uint32_t chunk_bits = P.below - LOGE; // log2(column chunks per hi value)
            const uint32_t hi = tile >> chunk_bits;
            rest0 = (tile & ((1u << chunk_bits) - 1u)) << LOGE; // first value of the low index
            in_base = (hi << (g + P.below)) + rest0;
        } else {
            const uint32_t o1_bits = P.g1 - LOGE;
            rest0 = (tile & ((1u << o1_bits) - 1u)) << LOGE; // first o_1 of the tile
            mid = tile >> o1_bits;
        }
    };
    // where element `a` of this rank's packed input lives: locally, or (natural blocks over peer memory) with its owner
    auto first_pass_src = [&](uint32_t a) -> const fr* {
        if (P.src_on) {
            const uint32_t at = (uint32_t)insert_bits(a, P.rk_pos, P.rk_bits, P.rk_val); // true coefficient index
            return P.peer_src[at >> P.block_log] + (at & ((1u << P.block_log) - 1u));
        }
        return P.src + a;
    };
    // PERSIST: start the copies of tile group `grp` into the staging area
    auto prefetch = [&](uint32_t grp) {
        const uint32_t t = grp * tiles_per_cta + tile_local;
        if (t < num_tiles) {
            uint32_t ib, r0, md;
            tile_coords(t, ib, r0, md);
            if constexpr (!LAST) {
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    const uint32_t row = (uint32_t)j * rows8 + q;
                    const uint4* src = reinterpret_cast<const uint4*>(first_pass_src(ib + (row << P.below) + col));
                    cp_async16(stage + (2 * j) * NTT_THREADS + threadIdx.x, src);
                    cp_async16(stage + (2 * j + 1) * NTT_THREADS + threadIdx.x, src + 1);
                }
            } else {
#pragma unroll
                for (uint32_t it = 0; it < (uint32_t)E; ++it) {
                    const uint32_t idx = it * R + tau; // [0, E R): column-major
                    const uint32_t c = idx >> g, row = idx & (R - 1);
                    const uint4* src = reinterpret_cast<const uint4*>(P.src + last_src_index(r0, md, c, row));
                    const uint32_t si = sm_idx<LOGE>(tile_local, R, row, c);
                    cp_async16(stage + si, src);
                    cp_async16(stage + half_stride + si, src + 1);
                }
            }
        }
        cp_async_commit();
    };

    uint32_t group = blockIdx.x;
    bool tw_pending = TMA_TW;
    if constexpr (PERSIST) {
        prefetch(group);
    }
    while (true) {
    const uint32_t tile = group * tiles_per_cta + tile_local;
    const bool active = tile < num_tiles;
    uint32_t in_base, rest0, mid;
    tile_coords(tile, in_base, rest0, mid);

    fr x[E];
    // the coset pre-scale of the first pass, fused into the load
    auto pre_scale = [&](fr& v, uint32_t a) {
        const uint32_t at = (uint32_t)insert_bits(a, P.rk_pos, P.rk_bits, P.rk_val); // true coefficient index
        if (P.pro_full != nullptr) {
            if (at < P.pro_size) v = fe_mul(v, fe_load_nc<FrParams>(P.pro_full + at));
        } else if (P.pro_lo != nullptr && at < P.pro_size) {
            fr sc = fe_mul(fe_load_nc<FrParams>(P.pro_hi + (at >> P.pro_split)),
                           fe_load_nc<FrParams>(P.pro_lo + (at & ((1u << P.pro_split) - 1u))));
            v = fe_mul(v, sc);
        }
    };
    // ---- load (round-0 register layout: rows j * R/8 + q, column col)
    if constexpr (PERSIST) {
        cp_async_wait_all();
        __syncthreads(); // the staged elements are visible to every thread, and the exchange buffer is free again
        if (active) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                if constexpr (!LAST) {
                    const uint4 a = stage[(2 * j) * NTT_THREADS + threadIdx.x], b = stage[(2 * j + 1) * NTT_THREADS + threadIdx.x];
                    x[j].l[0] = a.x; x[j].l[1] = a.y; x[j].l[2] = a.z; x[j].l[3] = a.w;
                    x[j].l[4] = b.x; x[j].l[5] = b.y; x[j].l[6] = b.z; x[j].l[7] = b.w;
                } else {
                    const uint32_t row = (uint32_t)j * rows8 + q;
                    x[j] = smem_load(stage, half_stride, sm_idx<LOGE>(tile_local, R, row, col));
                }
            }
        }
        if constexpr (LAST) {
            __syncthreads(); // every thread has taken its elements: the staging area may be refilled
        }
        if (group + gridDim.x < num_groups) {
            prefetch(group + gridDim.x);
        }
        if constexpr (!LAST) {
            if (active && (P.pro_full != nullptr || P.pro_lo != nullptr)) {
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    const uint32_t row = (uint32_t)j * rows8 + q;
                    pre_scale(x[j], in_base + (row << P.below) + col);
                }
            }
        }
    } else if constexpr (!LAST) {
        if (active) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const uint32_t row = (uint32_t)j * rows8 + q;
                const uint32_t a = in_base + (row << P.below) + col;
                x[j] = fe_load<FrParams>(first_pass_src(a));
                pre_scale(x[j], a);
            }
        }
    } else {
        // rows are contiguous in memory: stage through shared memory with lanes along the rows
        if (active) {
#pragma unroll 1
            for (uint32_t it = 0; it < (uint32_t)E; ++it) {
                const uint32_t idx = it * R + tau;   // [0, E R): column-major
                const uint32_t c = idx >> g, row = idx & (R - 1);
                fr v = fe_load<FrParams>(P.src + last_src_index(rest0, mid, c, row));
                smem_store(sm, half_stride, sm_idx<LOGE>(tile_local, R, row, c), v);
            }
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const uint32_t row = (uint32_t)j * rows8 + q;
                x[j] = smem_load(sm, half_stride, sm_idx<LOGE>(tile_local, R, row, col));
            }
        }
    }

    // ---- radix-E rounds over row bits g-1 .. 0
    int b_hi = (int)g - 1;
    uint32_t w0 = g - LOGE; // first round always has a full LOGE-bit window (g >= LOGE)
    bool first = true;
    while (true) {
        if (!first) {
            // exchange through shared memory: gather rows {hi_part, j, lo_part}
            __syncthreads();
            if (active) {
                const uint32_t lo_part = q & ((1u << w0) - 1), hi_part = q >> w0;
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    const uint32_t row = (hi_part << (w0 + LOGE)) | ((uint32_t)j << w0) | lo_part;
                    x[j] = smem_load(sm, half_stride, sm_idx<LOGE>(tile_local, R, row, col));
                }
            }
        }
        const uint32_t lo_part = q & ((1u << w0) - 1);
        if constexpr (TMA_TW) {
            if (tw_pending) { // the twiddles have landed (every thread waits: inactive ones too); once per CTA
                tw_bar.wait(std::move(tw_token));
                tw_pending = false;
            }
        }
        if (active) {
            radix_round<LOGE, TMA_TW>(x, g, w0, (uint32_t)b_hi, lo_part, stage_tw);
        }
        b_hi = (int)w0 - 1;
        if (b_hi < 0) {
            break;
        }
        // write back for the next round (each thread overwrites exactly the rows it read: no hazard)
        if (active) {
            const uint32_t hi_part = q >> w0;
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const uint32_t row = (hi_part << (w0 + LOGE)) | ((uint32_t)j << w0) | lo_part;
                smem_store(sm, half_stride, sm_idx<LOGE>(tile_local, R, row, col), x[j]);
            }
        }
        first = false;
        w0 = b_hi >= LOGE - 1 ? (uint32_t)b_hi - (LOGE - 1) : 0;
    }

    // ---- store.  After the last round (w0 == 0) thread holds rows (q << LOGE) | j; row rho holds X[bitrev_g(rho)].
    if (active) {
    if constexpr (!LAST) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const uint32_t rho = (q << LOGE) | (uint32_t)j;
            const uint32_t o = bitrev(rho, g);
            const uint32_t rest = (uint32_t)insert_bits(rest0 + col, P.rk_pos, P.rk_bits, P.rk_val);
            // inter-pass twiddle w_N^( 2^above * o * rest ); (x << above) mod N == (x mod (N >> above)) << above
            uint32_t e = ((o * rest) & (Nmask >> P.above)) << P.above;
            if (P.tw_scaled != nullptr) {
                if (P.inverse) {
                    e = (Nmask + 1u - e) & Nmask;
                }
                x[j] = fe_mul(x[j], fe_load_nc<FrParams>(P.tw_scaled + e));
            } else if (e != 0) {
                if (P.inverse) {
                    e = Nmask + 1u - e;
                }
                x[j] = fe_mul(x[j], fe_load_nc<FrParams>(P.tw_big + e));
            }
            const uint32_t a = in_base + (o << P.below) + col;
            if (P.peer_on) {
                const uint32_t owner = a >> P.chunk_log;
                fe_store(P.peer_dst[owner] + ((P.rk_val << P.chunk_log) | (a & ((1u << P.chunk_log) - 1u))), x[j]);
            } else {
                fe_store(P.dst + a, x[j]);
            }
        }
    } else {
        // natural-order output index: o_1 + sum_q o_q * 2^(bits before q) + o_P * 2^above
        uint32_t obase = (P.rk_val << P.g1) | (rest0 + col); // true o_1 (this rank owns its top rk_bits)
        if (P.rk_bits == 0) obase = rest0 + col;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            if ((uint32_t)d < P.num_mid) {
                const uint32_t dig = (mid >> P.mid_src_shift[d]) & ((1u << P.mid_bits[d]) - 1u);
                obase |= dig << P.mid_dst_shift[d];
            }
        }
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const uint32_t rho = (q << LOGE) | (uint32_t)j;
            const uint32_t o = obase | (bitrev(rho, g) << P.above);
            if (P.epi_mode == 1) {
                x[j] = fe_mul(x[j], P.epi_const);
            } else if (P.epi_mode == 2) {
                fr s = fe_mul(fe_load_nc<FrParams>(P.epi_hi + (o >> P.epi_split)),
                              fe_load_nc<FrParams>(P.epi_lo + (o & ((1u << P.epi_split) - 1u))));
                x[j] = fe_mul(x[j], s);
            } else if (P.epi_mode == 3) {
                x[j] = fe_mul(x[j], fe_load_nc<FrParams>(P.epi_full + o));
            }
            if (P.out_on) {
                fe_store(P.peer_out[o >> P.block_log] + (o & ((1u << P.block_log) - 1u)), x[j]); // the owner of natural index o
            } else {
                const uint32_t ol = P.rk_bits ? (uint32_t)squeeze_bits(o, P.g1, P.rk_bits) : o; // local slot of natural index o
                fe_store(P.dst + (((size_t)ol << P.out_shift) + P.out_off), x[j]);
            }
        }
    }
    } // active
    if constexpr (!PERSIST) {
        break;
    }
    group += gridDim.x;
    if (group >= num_groups) {
        break;
    }
    } // tile groups
}

// ---- N <= 32: direct O(N^2) DFT by one warp (everything fused, nothing worth tiling)
struct SmallParams {
    const fr* src;
    fr* dst;
    uint32_t log_n;
    uint32_t inverse;
    fr root;        // w_N or its inverse
    uint32_t has_pro;
    fr pro_start, pro_shift;
    uint64_t pro_size;
    uint32_t epi_mode; // 0 none, 1 const, 2 const * shift^o
    fr epi_const, epi_shift;
    uint32_t out_shift, out_off;
};
__global__ void __launch_bounds__(32) k_ntt_small(const SmallParams P)
{
    __shared__ fr xs[32];
    const uint32_t n = 1u << P.log_n;
    const uint32_t o = threadIdx.x;
    if (o < n) {
        fr v = fe_load<FrParams>(P.src + o);
        if (P.has_pro && o < P.pro_size) {
            v = fe_mul(v, fe_mul(P.pro_start, fe_pow(P.pro_shift, o)));
        }
        xs[o] = v;
    }
    __syncthreads();
    if (o >= n) {
        return;
    }
    const fr wo = fe_pow(P.root, o); // w^o
    fr acc = fe_zero<FrParams>();
    // Horner over i: sum_i x_i (w^o)^i
    for (int i = (int)n - 1; i >= 0; --i) {
        acc = fe_add(fe_mul(acc, wo), xs[i]);
    }
    if (P.epi_mode == 1) {
        acc = fe_mul(acc, P.epi_const);
    } else if (P.epi_mode == 2) {
        acc = fe_mul(acc, fe_mul(P.epi_const, fe_pow(P.epi_shift, o)));
    }
    fe_store(P.dst + (((uint64_t)o << P.out_shift) + P.out_off), acc);
}

// ---- table builders
// out[e] = base^e * start, e in [0, count): each thread does one exponentiation + a short running product
__global__ void __launch_bounds__(256) k_powers(fr* __restrict__ out, uint64_t count, fr base, fr start, uint32_t log_stride)
{
    // out[j] = start * base^(j << log_stride)
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t j0 = t * 8;
    if (j0 >= count) {
        return;
    }
    const fr step = fe_pow(base, 1ull << log_stride);
    fr cur = fe_mul(start, fe_pow(base, j0 << log_stride));
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
        if (j0 + j < count) {
            fe_store(out + j0 + j, fe_reduce_once(cur));
        }
        cur = fe_mul(cur, step);
    }
}

// fr: a primitive 2^28-th root of unity, Montgomery form (bb/ecc/curves/bn254/fr.hpp:27-30)
static const uint64_t FR_ROOT_28[4] = { 0x636e735580d13d9cULL, 0xa22bf3742445ffd6ULL, 0x56452ac01eb203d8ULL,
                                        0x1860ef942963f9e7ULL };

static fr to_dev(const hf::Fr& a)
{
    fr r;
    for (int i = 0; i < 4; ++i) {
        r.l[2 * i] = (uint32_t)a.d[i];
        r.l[2 * i + 1] = (uint32_t)(a.d[i] >> 32);
    }
    return r;
}

// get_root_of_unity(k): the 2^28-th root squared 28 - k times (bb/ecc/fields/field_impl.hpp:496-503)
hf::Fr ntt_root_of_unity(unsigned log_n)
{
    hf::Fr r;
    for (int i = 0; i < 4; ++i) r.d[i] = FR_ROOT_28[i];
    for (unsigned i = 28; i > log_n; --i) {
        r = hf::sqr(r);
    }
    return r;
}

static int launch_powers(Context* ctx, fr* out, uint64_t count, const hf::Fr& base, const hf::Fr& start, uint32_t log_stride, cudaStream_t st)
{
    const uint64_t threads = (count + 7) / 8;
    k_powers<<<div_up(threads, 256), 256, 0, st>>>(out, count, to_dev(base), to_dev(start), log_stride);
    ctx->launches += 1;
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

// stage twiddles: for g = 1..8, table_g[j] = w_{2^g}^j (dir 0) or w_{2^g}^-j (dir 1), j < 2^(g-1); offset(g) = 2^(g-1) - 1
static int ensure_stage_tables(Context* ctx, cudaStream_t st)
{
    if (ctx->ntt_stage_tw[0] != nullptr) {
        return BBG_OK;
    }
    for (int dir = 0; dir < 2; ++dir) {
        fr* tab = nullptr;
        BBG_CUDA(cudaMalloc(&tab, 256 * sizeof(fr)));
        for (unsigned g = 1; g <= 8; ++g) {
            hf::Fr w = ntt_root_of_unity(g);
            if (dir) {
                w = hf::invert(w);
            }
            int rc = launch_powers(ctx, tab + ((1u << (g - 1)) - 1), 1u << (g - 1), w, hf::one(), 0, st);
            if (rc) return rc;
        }
        ctx->ntt_stage_tw[dir] = tab;
    }
    return BBG_OK;
}

// ---- cached tables (w_N^e per size, geometric scaling tables) share ONE byte budget with LRU eviction across both kinds,
// so a run of large coset transforms cannot crowd out the MSM workspaces (each table is up to 32 N bytes).
static size_t ntt_table_budget()
{
    static const size_t budget = [] {
        const char* v = getenv("BBG_NTT_TABLE_MAX_MB");
        if (v && *v) return (size_t)atoll(v) << 20;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) total_b = (size_t)64 << 30;
        return std::max<size_t>(total_b / 12, (size_t)1 << 30); // 15 GB of a B200's 180 GB
    }();
    return budget;
}
// frees least-recently-used tables until `need` more bytes fit; `keep_log_n` / `keep_tab` are in use by the caller
static void ntt_tables_make_room(Context* ctx, size_t need, unsigned keep_log_n, const void* keep_tab)
{
    const size_t budget = ntt_table_budget();
    while (ctx->ntt_table_bytes + need > budget) {
        // oldest of both kinds
        uint64_t best = ~0ull;
        int kind = -1;
        unsigned tw_key = 0;
        size_t sc_idx = 0;
        for (auto& kv : ctx->ntt_twiddles) {
            if (kv.first == keep_log_n) continue;
            const uint64_t u = ctx->ntt_twiddle_use[kv.first];
            if (u < best) {
                best = u;
                kind = 0;
                tw_key = kv.first;
            }
        }
        for (size_t i = 0; i < ctx->ntt_scale_cache.size(); ++i) {
            const auto& e = ctx->ntt_scale_cache[i];
            if (e.tab == nullptr || e.tab == keep_tab) continue;
            if (e.last_use < best) {
                best = e.last_use;
                kind = 1;
                sc_idx = i;
            }
        }
        if (kind < 0) return; // nothing left to evict
        if (kind == 0) {
            cudaFree(ctx->ntt_twiddles[tw_key]); // cudaFree synchronises with any kernel still reading the table
            ctx->ntt_table_bytes -= sizeof(fr) << tw_key;
            ctx->ntt_twiddles.erase(tw_key);
            ctx->ntt_twiddle_use.erase(tw_key);
        } else {
            auto& e = ctx->ntt_scale_cache[sc_idx];
            cudaFree(e.tab);
            ctx->ntt_table_bytes -= e.count * sizeof(fr);
            ctx->ntt_scale_cache.erase(ctx->ntt_scale_cache.begin() + (long)sc_idx);
        }
    }
}

static int ensure_big_table(Context* ctx, unsigned log_n, const fr** out, cudaStream_t st)
{
    ctx->ntt_twiddle_use[log_n] = ++ctx->ntt_scale_clock;
    auto it = ctx->ntt_twiddles.find(log_n);
    if (it != ctx->ntt_twiddles.end()) {
        *out = (const fr*)it->second;
        return BBG_OK;
    }
    fr* tab = nullptr;
    const uint64_t N = 1ull << log_n;
    ntt_tables_make_room(ctx, N * sizeof(fr), log_n, nullptr);
    BBG_CUDA(cudaMalloc(&tab, N * sizeof(fr)));
    int rc = launch_powers(ctx, tab, N, ntt_root_of_unity(log_n), hf::one(), 0, st);
    if (rc) return rc;
    ctx->ntt_twiddles[log_n] = tab;
    ctx->ntt_table_bytes += N * sizeof(fr);
    *out = tab;
    return BBG_OK;
}

// Cached full table T[i] = start * shift^i, i < count (see Context::ntt_scale_cache).  *out stays nullptr when the key
// is new (first sighting), too large, or the build failed; the caller then uses the two-level tables.
static constexpr uint64_t NTT_SCALE_MAX_BYTES = 1ull << 30;  // per table
static constexpr size_t NTT_SCALE_MAX_TABLES = 8;
static int get_scale_table(Context* ctx, uint64_t count, const hf::Fr& start, const hf::Fr& shift, const fr** out, cudaStream_t st,
                           unsigned cur_log_n)
{
    *out = nullptr;
    if (count == 0 || count * sizeof(fr) > NTT_SCALE_MAX_BYTES || count * sizeof(fr) > ntt_table_budget() / 4) return BBG_OK;
    static const bool disabled = [] {
        const char* v = getenv("BBG_NTT_NO_SCALE_CACHE");
        return v && *v && atoi(v) != 0;
    }();
    if (disabled) return BBG_OK;
    auto& cache = ctx->ntt_scale_cache;
    const uint64_t now = ++ctx->ntt_scale_clock;
    for (size_t idx = 0; idx < cache.size(); ++idx) {
        auto& e = cache[idx];
        if (e.count == count && memcmp(e.start, start.d, 32) == 0 && memcmp(e.shift, shift.d, 32) == 0) {
            e.last_use = now;
            if (e.tab == nullptr) {
                uint64_t key_start[4], key_shift[4];
                memcpy(key_start, e.start, 32);
                memcpy(key_shift, e.shift, 32);
                ntt_tables_make_room(ctx, count * sizeof(fr), cur_log_n, nullptr); // may reshuffle `cache`: look the key up again
                Context::ScaleTab* ep = nullptr;
                for (auto& f : cache) {
                    if (f.count == count && memcmp(f.start, key_start, 32) == 0 && memcmp(f.shift, key_shift, 32) == 0) ep = &f;
                }
                if (ep == nullptr) return BBG_OK;
                fr* tab = nullptr;
                if (cudaMalloc(&tab, count * sizeof(fr)) != cudaSuccess) {
                    cudaGetLastError(); // out of memory: keep using the two-level path
                    return BBG_OK;
                }
                int rc = launch_powers(ctx, tab, count, shift, start, 0, st);
                if (rc) {
                    cudaFree(tab);
                    return rc;
                }
                ep->tab = tab;
                ctx->ntt_table_bytes += count * sizeof(fr);
                *out = (const fr*)tab;
                return BBG_OK;
            }
            *out = (const fr*)e.tab;
            return BBG_OK;
        }
    }
    if (cache.size() >= NTT_SCALE_MAX_TABLES) {
        size_t victim = 0;
        for (size_t i = 1; i < cache.size(); ++i) {
            if (cache[i].last_use < cache[victim].last_use) victim = i;
        }
        if (cache[victim].tab) {
            cudaFree(cache[victim].tab); // synchronises with any kernel still reading it
            ctx->ntt_table_bytes -= cache[victim].count * sizeof(fr);
        }
        cache.erase(cache.begin() + (long)victim);
    }
    Context::ScaleTab e;
    e.count = count;
    memcpy(e.start, start.d, 32);
    memcpy(e.shift, shift.d, 32);
    e.last_use = now;
    cache.push_back(e);
    return BBG_OK;
}

int ntt_root_table(Context* ctx, unsigned log_n, const void** table, cudaStream_t st)
{
    const fr* t = nullptr;
    int rc = ensure_big_table(ctx, log_n, &t, st);
    *table = t;
    return rc;
}
int ntt_root_table_at_least(Context* ctx, unsigned log_n, const void** table, unsigned* stride_log, cudaStream_t st)
{
    for (auto& kv : ctx->ntt_twiddles) {
        if (kv.first >= log_n) { // std::map: the smallest resident table that is large enough
            ctx->ntt_twiddle_use[kv.first] = ++ctx->ntt_scale_clock;
            *table = kv.second;
            *stride_log = kv.first - log_n;
            return BBG_OK;
        }
    }
    *stride_log = 0;
    return ntt_root_table(ctx, log_n, table, st);
}

// In-place (src == dst allowed) transform of 2^log_n elements on the device.
//   pro : x[i] *= start * shift^i for i < size, before the transform
//   epi : X[o] *= start (* shift^o), after the transform
//   out index = (o << out_shift) + out_off inside dst
//
// Multi-GPU (dist.rank_bits = log2(world) > 0): the transform is the same P-pass factorisation with ONE exchange.
//   phase 0: passes 1..P-1 on the local sub-array {i : bits [g_P - rb, g_P) of i == rank} (packed), all of whose
//            sub-transforms are local because those bits lie below every digit handled so far;
//   [all-to-all of equal contiguous chunks: chunk r' of every rank goes to rank r' -- done by the caller, NCCL]
//   phase 1: the last pass on the received buffer; this rank ends up with {k : bits [g_1 - rb, g_1) of k == rank}, packed.
int ntt_device(Context* ctx, const void* d_src, void* d_dst, unsigned log_n, bool inverse, const NttScale& pro, const NttScale& epi,
               unsigned out_shift, unsigned out_off, cudaStream_t st, const NttDist& dist)
{
    if (log_n > 28) {
        set_last_error("ntt: fr has 2-adicity 28, log2(n) must be <= 28");
        return BBG_ERR_ARG;
    }
    const uint64_t N = 1ull << log_n;
    hf::Fr root = ntt_root_of_unity(log_n);
    if (inverse) {
        root = hf::invert(root);
    }
    const unsigned rb = dist.rank_bits;
    if (rb > 0 && (log_n < 12 || out_shift != 0 || out_off != 0 || dist.rank >= (1u << rb))) {
        set_last_error("ntt: the multi-GPU path needs log2(n) >= 12 and no output interleave");
        return BBG_ERR_ARG;
    }
    if (log_n <= 5) {
        SmallParams sp;
        sp.src = (const fr*)d_src;
        sp.dst = (fr*)d_dst;
        sp.log_n = log_n;
        sp.inverse = inverse;
        sp.root = to_dev(root);
        sp.has_pro = pro.present;
        sp.pro_start = to_dev(pro.present ? pro.start : hf::one());
        sp.pro_shift = to_dev(pro.present && pro.has_shift ? pro.shift : hf::one());
        sp.pro_size = pro.size;
        sp.epi_mode = epi.present ? (epi.has_shift ? 2 : 1) : 0;
        sp.epi_const = to_dev(epi.present ? epi.start : hf::one());
        sp.epi_shift = to_dev(epi.present && epi.has_shift ? epi.shift : hf::one());
        sp.out_shift = out_shift;
        sp.out_off = out_off;
        k_ntt_small<<<1, 32, 0, st>>>(sp);
        ctx->launches += 1;
        BBG_CUDA(cudaGetLastError());
        return BBG_OK;
    }

    int rc;
    Profiler& pr = ctx->prof;
    pr.begin();
    pr.mark(st, PH_NTT_TABLES);
    if ((rc = ensure_stage_tables(ctx, st))) return rc;
    const fr* tw_big = nullptr;
    if ((rc = ensure_big_table(ctx, log_n, &tw_big, st))) return rc;
    if (rb == 0 && (rc = ctx->ntt_scratch.reserve(N * sizeof(fr)))) return rc;
    fr* scratch = (fr*)ctx->ntt_scratch.p;

    // factorisation: P passes of nearly equal size, each 3..8 bits
    const unsigned num_passes = log_n <= 16 ? 2 : (log_n <= 24 ? 3 : 4);
    unsigned gb[NTT_MAX_PASSES];
    for (unsigned p = 0; p < num_passes; ++p) {
        gb[p] = log_n / num_passes + (p < log_n % num_passes ? 1 : 0);
    }

    // two-level scaling tables (lo: shift^j, hi: start * shift^(j << split))
    const uint32_t split = (log_n + 1) / 2;
    const fr *pro_lo = nullptr, *pro_hi = nullptr, *epi_lo = nullptr, *epi_hi = nullptr;
    const uint64_t lo_count = 1ull << split, hi_count = 1ull << (log_n - split);
    const fr *pro_full = nullptr, *epi_full = nullptr, *tw_scaled = nullptr;
    if (pro.present && rb == 0) {
        if ((rc = get_scale_table(ctx, std::min<uint64_t>(pro.size, N), pro.start, pro.has_shift ? pro.shift : hf::one(), &pro_full, st, log_n))) return rc;
    }
    if (epi.present && epi.has_shift && rb == 0) {
        if ((rc = get_scale_table(ctx, N, epi.start, epi.shift, &epi_full, st, log_n))) return rc;
    }
    if (epi.present && !epi.has_shift && rb == 0 && inverse) {
        // plain ifft: fold 1/n into the last inter-pass twiddle (table 1/n * w^e, indexed by the forward exponent)
        const hf::Fr n_inv = hf::invert(hf::from_u64(N));
        if (memcmp(hf::reduce(epi.start).d, hf::reduce(n_inv).d, 32) == 0) {
            if ((rc = get_scale_table(ctx, N, hf::reduce(n_inv), ntt_root_of_unity(log_n), &tw_scaled, st, log_n))) return rc;
        }
    }
    if (pro.present && pro_full == nullptr) {
        if ((rc = ctx->ntt_pro.reserve((lo_count + hi_count) * sizeof(fr)))) return rc;
        fr* lo = (fr*)ctx->ntt_pro.p;
        fr* hi = lo + lo_count;
        const hf::Fr sh = pro.has_shift ? pro.shift : hf::one();
        if ((rc = launch_powers(ctx, lo, lo_count, sh, hf::one(), 0, st))) return rc;
        if ((rc = launch_powers(ctx, hi, hi_count, sh, pro.start, split, st))) return rc;
        pro_lo = lo;
        pro_hi = hi;
    }
    if (epi.present && epi.has_shift && epi_full == nullptr) {
        if ((rc = ctx->ntt_epi.reserve((lo_count + hi_count) * sizeof(fr)))) return rc;
        fr* lo = (fr*)ctx->ntt_epi.p;
        fr* hi = lo + lo_count;
        if ((rc = launch_powers(ctx, lo, lo_count, epi.shift, hf::one(), 0, st))) return rc;
        if ((rc = launch_powers(ctx, hi, hi_count, epi.shift, epi.start, split, st))) return rc;
        epi_lo = lo;
        epi_hi = hi;
    }

    if (!ctx->ntt_attr_set) { // per context: a re-init on another device needs the attributes again
        BBG_CUDA(cudaFuncSetAttribute(k_ntt_pass<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NttGeom<3>::SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute(k_ntt_pass<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NttGeom<3>::SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute(k_ntt_pass<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NttGeom<2>::SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute(k_ntt_pass<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NttGeom<2>::SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute((k_ntt_pass<2, true, 3, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, NttGeom<2>::SMEM_BYTES + NTT_TW_SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute((k_ntt_pass<2, false, 3, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, NttGeom<2>::SMEM_BYTES + NTT_TW_SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute((k_ntt_pass<2, true, 3, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_PERSIST_SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute((k_ntt_pass<2, false, 3, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_PERSIST_SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute((k_ntt_pass<2, true, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, NttGeom<2>::SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute((k_ntt_pass<2, false, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, NttGeom<2>::SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute(k_ntt_pass<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NttGeom<1>::SMEM_BYTES));
        BBG_CUDA(cudaFuncSetAttribute(k_ntt_pass<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NttGeom<1>::SMEM_BYTES));
        ctx->ntt_attr_set = true;
    }
    // elements per thread (see NttGeom): 2^loge.  Measured on B200 (fft, ms; E = 2 / 4 / 8): 2^16 0.039 / 0.052 / 0.102 (the grid is
    // only n / (256 E) CTAs there), 2^18 0.091 / 0.088 / 0.116, 2^20 0.249 / 0.241 / 0.287, 2^22 0.948 / 0.892 / 0.948,
    // 2^24 3.94 / 3.55 / 3.75.
    static const unsigned loge_env = [] {
        const char* v = getenv("BBG_NTT_LOGE");
        return v && *v ? (unsigned)atoi(v) : 0u;
    }();
    const unsigned loge = (loge_env >= 1 && loge_env <= 3) ? loge_env : (log_n <= 16 ? 1u : NTT_DEFAULT_LOGE);
    static const bool tma_twiddles = [] {
        const char* v = getenv("BBG_NTT_TMA_TWIDDLES"); // default on; 0 = read the stage twiddles through the read-only path
        return !(v && *v == '0');
    }();
    static const int persist_env = [] {
        const char* v = getenv("BBG_NTT_PERSIST"); // 1: persistent CTAs with cp.async prefetch of the next tile group (E = 4); 0: one tile group per CTA
        return v && *v ? atoi(v) : -1;
    }();
    static const bool e4_two_ctas = [] {
        const char* v = getenv("BBG_NTT_E4_CTAS");
        return v && *v == '2';
    }();

    if (rb > 0 && (gb[num_passes - 1] < rb + 3 || gb[0] < rb + 3)) {
        set_last_error("ntt: too many ranks for this transform size");
        return BBG_ERR_ARG;
    }
    const unsigned log_local = log_n - rb;
    unsigned above = 0;
    for (unsigned p = 0; p < num_passes; ++p) {
        PassParams pp;
        const bool last = (p == num_passes - 1);
        if (rb == 0) {
            pp.src = (p == 0) ? (const fr*)d_src : scratch;
            pp.dst = last ? (fr*)d_dst : scratch;
        } else {
            if ((dist.phase == 0) == last) { // phase 0 runs every pass but the last, phase 1 only the last
                above += gb[p];
                continue;
            }
            pp.src = (p == 0 || last) ? (const fr*)d_src : (const fr*)d_dst; // middle passes run in place in dst
            pp.dst = (fr*)d_dst;
        }
        pp.tw_big = tw_big;
        pp.stage_tw = (const fr*)ctx->ntt_stage_tw[inverse ? 1 : 0] + ((1u << (gb[p] - 1)) - 1);
        pp.log_n = log_local;
        pp.tw_log_n = log_n;
        pp.g = gb[p];
        pp.above = above;
        pp.below = log_local - above - gb[p];
        pp.last = last;
        pp.inverse = inverse;
        pp.g1 = gb[0];
        pp.rk_bits = rb;
        pp.rk_pos = gb[num_passes - 1] - rb; // position of the rank bits inside the input index
        pp.rk_val = dist.rank;
        pp.mid_bits_total = log_n - gb[0] - gb[num_passes - 1];
        pp.split_low = gb[num_passes - 1] - rb;
        pp.chunk_log = log_local - rb;
        if (rb > 0 && last) {
            pp.g1 = gb[0] - rb;   // local o_1 (its top rb bits are this rank)
            pp.below = 0;
        }
        pp.num_mid = 0;
        for (int d = 0; d < 2; ++d) {
            pp.mid_bits[d] = pp.mid_src_shift[d] = pp.mid_dst_shift[d] = 0;
        }
        if (last) {
            // mid = (o_2, ..., o_{P-1}) with o_2 most significant
            pp.num_mid = num_passes - 2;
            unsigned dst_shift = gb[0];
            unsigned remaining = above - gb[0];
            for (unsigned d = 0; d < pp.num_mid; ++d) {
                const unsigned bits = gb[1 + d];
                remaining -= bits;
                pp.mid_bits[d] = bits;
                pp.mid_src_shift[d] = remaining;
                pp.mid_dst_shift[d] = dst_shift;
                dst_shift += bits;
            }
        }
        pp.peer_on = 0;
        for (auto& q : pp.peer_dst) q = nullptr;
        if (rb > 0 && dist.peer_recv != nullptr && dist.phase == 0 && p + 2 == num_passes) {
            pp.peer_on = 1;
            for (unsigned r = 0; r < (1u << rb); ++r) pp.peer_dst[r] = (fr*)dist.peer_recv[r];
        }
        pp.src_on = pp.out_on = 0;
        pp.block_log = log_local;
        for (auto& q : pp.peer_src) q = nullptr;
        for (auto& q : pp.peer_out) q = nullptr;
        if (rb > 0 && dist.peer_src != nullptr && dist.phase == 0 && p == 0) {
            pp.src_on = 1;
            for (unsigned r = 0; r < (1u << rb); ++r) pp.peer_src[r] = (const fr*)dist.peer_src[r];
        }
        if (rb > 0 && dist.peer_out != nullptr && dist.phase == 1 && last) {
            pp.out_on = 1;
            for (unsigned r = 0; r < (1u << rb); ++r) pp.peer_out[r] = (fr*)dist.peer_out[r];
        }
        pp.pro_lo = (p == 0) ? pro_lo : nullptr;
        pp.pro_hi = (p == 0) ? pro_hi : nullptr;
        pp.pro_full = (p == 0) ? pro_full : nullptr;
        pp.epi_full = nullptr;
        pp.tw_scaled = (p + 2 == num_passes) ? tw_scaled : nullptr;
        pp.pro_split = split;
        pp.pro_size = pro.present ? pro.size : 0;
        pp.epi_mode = 0;
        pp.epi_lo = pp.epi_hi = nullptr;
        pp.epi_split = split;
        pp.epi_const = to_dev(hf::one());
        pp.out_shift = 0;
        pp.out_off = 0;
        if (last) {
            if (epi.present) {
                pp.epi_mode = epi.has_shift ? 2 : 1;
                pp.epi_lo = epi_lo;
                pp.epi_hi = epi_hi;
                pp.epi_const = to_dev(epi.start);
                if (epi_full != nullptr) {
                    pp.epi_mode = 3;
                    pp.epi_full = epi_full;
                } else if (tw_scaled != nullptr) {
                    pp.epi_mode = 0; // the constant rode in on the previous pass's twiddles
                }
            }
            pp.out_shift = out_shift;
            pp.out_off = out_off;
        }
        const uint64_t num_tiles = (N >> rb) >> (gb[p] + loge);
        const unsigned tiles_per_cta = NTT_THREADS >> gb[p];
        const unsigned blocks = (unsigned)((num_tiles + tiles_per_cta - 1) / tiles_per_cta);
        pr.mark(st, PH_NTT_PASS0 + (int)p);
        if (loge == 1) {
            if (last) k_ntt_pass<1, true><<<blocks, NTT_THREADS, NttGeom<1>::SMEM_BYTES, st>>>(pp);
            else k_ntt_pass<1, false><<<blocks, NTT_THREADS, NttGeom<1>::SMEM_BYTES, st>>>(pp);
        } else if (loge == 2 && e4_two_ctas) {
            // experiment knob (BBG_NTT_E4_CTAS=2): 128 registers, no spills, 4 instead of 6 warps per scheduler
            if (last) k_ntt_pass<2, true, 2><<<blocks, NTT_THREADS, NttGeom<2>::SMEM_BYTES, st>>>(pp);
            else k_ntt_pass<2, false, 2><<<blocks, NTT_THREADS, NttGeom<2>::SMEM_BYTES, st>>>(pp);
        } else if (loge == 2 && tma_twiddles && (persist_env > 0 || (persist_env < 0 && NTT_PERSIST_DEFAULT)) && blocks > 3u * ctx->num_sms) {
            const unsigned grid = 3u * (unsigned)ctx->num_sms; // one wave: three CTAs per SM
            if (last) k_ntt_pass<2, true, 3, true, true><<<grid, NTT_THREADS, NTT_PERSIST_SMEM_BYTES, st>>>(pp);
            else k_ntt_pass<2, false, 3, true, true><<<grid, NTT_THREADS, NTT_PERSIST_SMEM_BYTES, st>>>(pp);
        } else if (loge == 2 && tma_twiddles) {
            if (last) k_ntt_pass<2, true, 3, true><<<blocks, NTT_THREADS, NttGeom<2>::SMEM_BYTES + NTT_TW_SMEM_BYTES, st>>>(pp);
            else k_ntt_pass<2, false, 3, true><<<blocks, NTT_THREADS, NttGeom<2>::SMEM_BYTES + NTT_TW_SMEM_BYTES, st>>>(pp);
        } else if (loge == 2) {
            if (last) k_ntt_pass<2, true><<<blocks, NTT_THREADS, NttGeom<2>::SMEM_BYTES, st>>>(pp);
            else k_ntt_pass<2, false><<<blocks, NTT_THREADS, NttGeom<2>::SMEM_BYTES, st>>>(pp);
        } else {
            if (last) k_ntt_pass<3, true><<<blocks, NTT_THREADS, NttGeom<3>::SMEM_BYTES, st>>>(pp);
            else k_ntt_pass<3, false><<<blocks, NTT_THREADS, NttGeom<3>::SMEM_BYTES, st>>>(pp);
        }
        ctx->launches += 1;
        above += gb[p];
    }
    pr.mark(st, -1);
    BBG_CUDA(cudaGetLastError());
    return BBG_OK;
}

} // namespace bbg
