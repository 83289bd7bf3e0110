/* oracle/bn254_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the barretenberg hot path (BN254 fq/fr Montgomery arithmetic,
 * g1 group law, Pippenger MSM semantics, radix-2 NTT family, SRS transcript decode).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arms
 * may load this library, and only as the checker.  The product (aztec-2.0_b200/) never
 * links, loads or calls it.
 *
 * Parity is PINNED: tests/test_oracle_*.py check every function here against
 *   (a) the reference's own L0 known-answer constants (fq.test.cpp / g1.test.cpp, cited there),
 *   (b) tests/golden/ *.json vectors generated from the unmodified reference (oracle/_ref),
 *   (c) oracle/_ref/libbbref.so itself when present.
 *
 * Layouts are the reference's: field = 4 x u64 little-endian limbs, Montgomery form R = 2^256,
 * values in [0, 2p) ("coarse"); affine = {x, y} 64 B, infinity <=> bit 63 of x.data[3];
 * Jacobian = {x, y, z} 96 B, infinity <=> bit 63 of x.data[3].
 */
#ifndef BN254_ORACLE_H
#define BN254_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t d[4]; } orc_fe;              /* fq or fr element */
typedef struct { orc_fe x, y; } orc_affine;            /* 64 B  */
typedef struct { orc_fe x, y, z; } orc_jac;            /* 96 B  */

enum { ORC_FQ = 0, ORC_FR = 1 };

/* op codes shared with oracle/ref_shim.cpp: 0 mul 1 add 2 sub 3 sqr 4 to_mont 5 from_mont 6 invert 7 reduce_once 8 neg */
void orc_field_op(int field, int op, const orc_fe* a, const orc_fe* b, orc_fe* out);
void orc_fr_pow(const orc_fe* a, uint64_t e, orc_fe* out);
void orc_fr_root_of_unity(unsigned log2n, orc_fe* out);          /* field_impl.hpp:496-503 */
void orc_fr_coset_generator(orc_fe* out);                        /* fr.hpp:44-59 idx 0 == 5 */
void orc_fq_beta(orc_fe* out);                                   /* fq.hpp:21-24 */

void orc_g1_mixed_add(const orc_jac* a, const orc_affine* b, orc_jac* out);
void orc_g1_add(const orc_jac* a, const orc_jac* b, orc_jac* out);
void orc_g1_dbl(const orc_jac* a, orc_jac* out);
void orc_g1_set_infinity(orc_jac* out);
void orc_g1_to_affine(const orc_jac* a, orc_affine* out);
void orc_g1_affine_to_buffer(const orc_affine* a, uint8_t* buf64);
void orc_g1_jac_to_buffer(const orc_jac* a, uint8_t* buf64);
void orc_g1_mul(const orc_affine* p, const orc_fe* scalar_mont, orc_jac* out);
int  orc_g1_on_curve(const orc_affine* a);
void orc_g1_sum(const orc_jac* pts, size_t n, orc_jac* out);
void orc_g1_one(orc_affine* out);

/* srs/io.cpp */
void orc_read_g1_elements_from_buffer(orc_affine* elements, const uint8_t* buffer, size_t buffer_size);
int  orc_read_transcript_g1(orc_affine* monomials, size_t degree, const char* dir);
/* scalar_multiplication.cpp:104-112 */
void orc_generate_pippenger_point_table(const orc_affine* points, orc_affine* table, size_t n);

/* MSM: sum_i scalars[i] * table[2*i]   (table = the 2n interleaved table, or stride 1 for plain points) */
void orc_pippenger(const orc_fe* scalars, const orc_affine* points, size_t n, size_t point_stride, orc_jac* out);
void orc_naive_msm(const orc_fe* scalars, const orc_affine* points, size_t n, size_t point_stride, orc_jac* out);

/* NTT family. kind: 0 fft 1 ifft 2 coset_fft 3 coset_ifft 4 fft_with_constant 5 ifft_with_constant
 *                   6 coset_fft_with_constant 7 coset_fft_with_generator_shift */
void orc_ntt(int kind, orc_fe* coeffs, size_t n, size_t generator_size, const orc_fe* constant);
void orc_coset_fft_ext(orc_fe* coeffs, size_t n, size_t ext);
void orc_evaluate(const orc_fe* coeffs, const orc_fe* z, size_t n, orc_fe* out);
/* root, root_inverse, domain, domain_inverse, generator, generator_inverse */
void orc_domain_constants(size_t n, orc_fe* out6);

void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
