// poly.hpp -- parameter blocks of the pointwise / scan kernels (poly.cu) and their host launchers.
#pragma once
#include "ctx.cuh"
#include "field.cuh"
#include "host_field.hpp"

namespace bbg {

// indices of the polynomial table: waffle::PolynomialIndex (bb/plonk/proof_system/types/polynomial_manifest.hpp:10-50)
enum {
    BBG_POLY_Q_1 = 0, BBG_POLY_Q_2, BBG_POLY_Q_3, BBG_POLY_Q_4, BBG_POLY_Q_5, BBG_POLY_Q_M, BBG_POLY_Q_C, BBG_POLY_Q_ARITHMETIC_SELECTOR,
    BBG_POLY_Q_FIXED_BASE_SELECTOR, BBG_POLY_Q_RANGE_SELECTOR, BBG_POLY_Q_SORT_SELECTOR, BBG_POLY_Q_LOGIC_SELECTOR, BBG_POLY_TABLE_1,
    BBG_POLY_TABLE_2, BBG_POLY_TABLE_3, BBG_POLY_TABLE_4, BBG_POLY_TABLE_INDEX, BBG_POLY_TABLE_TYPE, BBG_POLY_Q_MIMC_COEFFICIENT,
    BBG_POLY_Q_MIMC_SELECTOR, BBG_POLY_Q_ELLIPTIC, BBG_POLY_SIGMA_1, BBG_POLY_SIGMA_2, BBG_POLY_SIGMA_3, BBG_POLY_SIGMA_4, BBG_POLY_ID_1,
    BBG_POLY_ID_2, BBG_POLY_ID_3, BBG_POLY_ID_4, BBG_POLY_W_1, BBG_POLY_W_2, BBG_POLY_W_3, BBG_POLY_W_4, BBG_POLY_S, BBG_POLY_Z,
    BBG_POLY_Z_LOOKUP, BBG_POLY_COUNT
};
static_assert(BBG_POLY_COUNT == BBG_NUM_POLYNOMIALS, "include/bbg.h BBG_NUM_POLYNOMIALS");

struct TurboParams {
    const Fe<FrParams>* p[BBG_POLY_COUNT];
    Fe<FrParams>* quotient;
    uint32_t n_large;
    Fe<FrParams> alpha_pow[7]; // alpha_base * alpha^k
    Fe<FrParams> alpha;
    Fe<FrParams> c_one, c_two, c_three, c_six, c_seven, c_17, c_81, c_83;
};
struct PermParams {
    const Fe<FrParams>* wires[4];
    const Fe<FrParams>* sigmas[4];
    const Fe<FrParams>* z;
    const Fe<FrParams>* l_start;
    const Fe<FrParams>* roots; // w_{4n}^i
    Fe<FrParams>* quotient;
    uint32_t n_large, width, roots_cut;
    Fe<FrParams> g_beta, beta, gamma, alpha_base, alpha_squared, public_input_delta, c_one;
    Fe<FrParams> coset_gen[3];
};
struct VanishParams {
    uint32_t n_large, subgroup, roots_cut;
    Fe<FrParams> g;
    Fe<FrParams> inv_sub[8];
    Fe<FrParams> numer[4];
};
struct LagrangeParams {
    uint32_t n_large, subgroup;
    Fe<FrParams> g, c_one;
    Fe<FrParams> numer_sub[8];
};
struct GrandParams {
    const Fe<FrParams>* wires[4];
    const Fe<FrParams>* sigmas[4];
    const Fe<FrParams>* roots;
    uint32_t root_stride_log, n, width;
    Fe<FrParams> beta, gamma;
    Fe<FrParams> coset_gen[3];
};

static constexpr unsigned EVAL_BATCH_MAX = 40;
struct EvalBatchParams {
    const Fe<FrParams>* coeffs[EVAL_BATCH_MAX];
    uint32_t n[EVAL_BATCH_MAX];
    Fe<FrParams> z[EVAL_BATCH_MAX];
    Fe<FrParams> z_chunk[EVAL_BATCH_MAX];
};
static constexpr unsigned LINCOMB_MAX = 48;
struct LinCombParams {
    Fe<FrParams>* dest;
    const Fe<FrParams>* base;
    uint32_t n, count;
    const Fe<FrParams>* polys[LINCOMB_MAX];
    Fe<FrParams> scalars[LINCOMB_MAX];
};

// host-side argument bundles (device pointers)
struct PermArgs {
    const void* d_wires[4];
    const void* d_sigmas[4];
    const void* d_z;
    const void* d_l_start;
    void* d_quotient;
    size_t n_large;
    unsigned width, roots_cut;
    hf::Fr alpha_base, beta, gamma, public_input_delta;
};
struct GrandArgs {
    const void* d_wires[4];
    const void* d_sigmas[4];
    void* d_z;
    size_t n;
    unsigned width;
    hf::Fr beta, gamma;
};

int poly_turbo_quotient_device(Context* ctx, int kind, const void* const* d_polys, size_t n_large, const void* alpha_base, const void* alpha,
                               void* d_quotient, cudaStream_t st);
int poly_permutation_quotient_device(Context* ctx, const PermArgs& A, cudaStream_t st);
int poly_divide_vanishing_device(Context* ctx, void* d_q, size_t n_small, size_t n_large, unsigned roots_cut, cudaStream_t st);
int poly_lagrange_l1_device(Context* ctx, void* d_out, size_t n_small, size_t n_large, cudaStream_t st);
int poly_grand_product_device(Context* ctx, const GrandArgs& A, cudaStream_t st);
int poly_evaluate_batch_device(Context* ctx, const void* const* d_coeffs, const size_t* n, const void* zs, size_t count, void* d_out, cudaStream_t st);
int poly_evaluate_device(Context* ctx, const void* d_coeffs, size_t n, const hf::Fr& z, void* d_out, cudaStream_t st);
int poly_opening_device(Context* ctx, const void* d_src, size_t n_in, size_t n_out, const hf::Fr& z, void* d_dest, void* d_f_at_z, cudaStream_t st);
int poly_linear_combination_device(Context* ctx, void* d_dest, const void* d_base, const void* const* d_polys, const void* scalars, size_t count,
                                   size_t n, cudaStream_t st);
int poly_copy_pad_device(Context* ctx, const void* d_src, void* d_dst, size_t n, size_t total, cudaStream_t st);

// ntt.cu: the cached table w_N^e, e in [0, N) (canonical values), N = 2^log_n; built on `st` on first use
int ntt_root_table(Context* ctx, unsigned log_n, const void** table, cudaStream_t st);
// a table for some N >= 2^log_n that is already resident (else this size's own); w_n^i = table[i << *stride_log]
int ntt_root_table_at_least(Context* ctx, unsigned log_n, const void** table, unsigned* stride_log, cudaStream_t st);

} // namespace bbg
