/* bbg.h -- C-ABI of libbbg.so: B200-native BN254 G1 MSM + fr NTT behind barretenberg's prover API.
 *
 * Plain C, plain pointers and sizes, int status codes (0 = ok; bbg_last_error() has the text).
 * Every entry point names the barretenberg interface it replaces; "bb/" abbreviates
 * barretenberg/src/aztec/ in AztecProtocol/aztec-2.0.  Data layouts are barretenberg's:
 *   fr / fq            32 B, 4 x u64 little-endian limbs, Montgomery form (R = 2^256), any value in [0, 2p)
 *   g1::affine_element 64 B {x, y}; point at infinity <=> bit 255 of x
 *   g1::element        96 B {x, y, z} Jacobian; same infinity flag
 * "host" pointers are ordinary (pageable or pinned) host memory; "_dev" variants take device
 * pointers and a cudaStream_t (as void*) and never synchronise.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point returns BBG_ERR_NO_DEVICE.
 */
#ifndef BBG_H
#define BBG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* every entry point below has default visibility; everything else in libbbg.so is hidden */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define BBG_OK 0
#define BBG_ERR_CUDA 1
#define BBG_ERR_ARG 2
#define BBG_ERR_NO_DEVICE 3
#define BBG_ERR_SRS 4
#define BBG_ERR_IO 5

/* ---- lifecycle ------------------------------------------------------------------------------ */
/* Bind the process to a CUDA device (-1 = the caller's current device).  Idempotent for the same device; returns
 * BBG_ERR_ARG if the library is already bound to a DIFFERENT device (any earlier call creates the context on the
 * then-current device).  Every entry point binds that device for its own duration and restores the caller's.
 *
 * Streams: the host-pointer entry points run on the library's own stream and return when the result is in host
 * memory.  The "_dev" entry points queue on the caller's stream and return immediately; they share ONE set of
 * workspaces and cached tables, so the library orders every call after the previous one with an event when the
 * streams differ -- calls from different streams are safe but execute one after the other on the device. */
int bbg_init(int device);
void bbg_shutdown(void);
const char* bbg_last_error(void);
int bbg_device_count(void);
uint64_t bbg_kernel_launches(void);    /* kernels this library has launched so far (bench.py reports it) */
double bbg_last_device_ms(void);       /* CUDA-event time of the kernels of the last host-pointer call */

/* Totals of the host-pointer entry points since the library was loaded -- calls, H2D bytes, D2H bytes for each of
 * msm, ntt, srs, poly (12 values).  BBG_STATS=1 prints the same (plus wall / kernel time) at exit. */
int bbg_stats_totals(uint64_t* out12);

/* Per-phase device timing of the LAST compute call (CUDA events on the launching stream; measurement aid
 * for bench.py, no reference counterpart).  Phases: 0 msm digits+histogram, 1 scan, 2 scatter, 3 pairwise affine
 * passes (optional path), 4 bucket accumulate, 5 slot merge, 6 bucket reduce, 7 window combine, 8 ntt tables,
 * 9..12 ntt pass 0..3.
 * bbg_profile_read synchronises on the recorded events and fills ms[0..n) (0 for phases that did not run). */
#define BBG_NUM_PHASES 13
int bbg_profile(int enable);
int bbg_profile_read(double* ms, int n);

/* bb/ecc/curves/bn254/scalar_multiplication/c_bind.cpp:11-19  bbmalloc / bbfree.
 * Returns page-locked host memory so the host-pointer entry points copy at full PCIe rate. */
void* bbg_malloc(size_t size);
void bbg_free(void* ptr);

/* ---- SRS + Pippenger object ----------------------------------------------------------------- */
/* bb/ecc/curves/bn254/scalar_multiplication/c_bind.cpp:21-29 new_pippenger(points, num_points):
 * `points` = raw transcript bytes of num_points-1 G1 points (64 B each, bb/srs/io.cpp:47-67 format);
 * monomial 0 is the generator.  Decoded (bswap + to-Montgomery) on the device. */
void* bbg_new_pippenger(const uint8_t* points, size_t num_points);
/* bb/.../pippenger.cpp:18-25 Pippenger(path, num_points): reads <dir>/transcriptNN.dat (bb/srs/io.cpp:134-162). */
void* bbg_new_pippenger_from_path(const char* srs_dir, size_t num_points);
/* Adopt a host table already in barretenberg's 2n-entry interleaved form [P0, phi(P0), P1, ...]
 * (what ProverReferenceString::get_monomials() returns, bb/plonk/reference_string/reference_string.hpp:24-38).
 * The host pointer is remembered so bbg_pippenger() can recognise it. */
void* bbg_new_pippenger_from_table(const void* table2n, size_t num_points);
/* Tell the library that `table2n` (host) holds this object's 2n interleaved table, so that bbg_pippenger() calls
 * made with a pointer into it (ProverReferenceString::get_monomials(), pippenger.cpp:27-31 monomials_ + 2*from)
 * run on the resident device copy instead of re-uploading the bases. */
int bbg_pippenger_bind_host_table(void* pippenger, const void* table2n);
/* When enabled (or BBG_AUTO_ADOPT=1 in the environment), bbg_pippenger() adopts an unknown table of >= 2^12 points on
 * first sight as bbg_new_pippenger_from_table would -- valid only if the caller never changes that memory
 * (true for a PLONK SRS); off by default. */
int bbg_set_auto_adopt(int enable);
/* Adopt a plain host array of num_points affine elements. */
void* bbg_new_pippenger_from_points(const void* points, size_t num_points);
/* Same from device memory (device-to-device copy; used for synthetic bases built on the GPU). */
void* bbg_new_pippenger_from_device_points(const void* d_points, size_t num_points);
/* c_bind.cpp:31-34 delete_pippenger */
void bbg_delete_pippenger(void* pippenger);
/* pippenger.hpp:47-49 get_num_points / get_point_table (copies the 2n interleaved table to host memory) */
size_t bbg_pippenger_num_points(void* pippenger);
int bbg_pippenger_get_point_table(void* pippenger, void* table2n_out);
const void* bbg_pippenger_device_points(void* pippenger); /* n contiguous affine points in HBM */
/* No reference counterpart (the CPU picks its window in runtime_states.hpp:9-63): the signed-window width c and the
 * number of precomputed fixed-base levels 2^(c l) P_i this object holds in HBM; reported by bench.py. */
unsigned bbg_pippenger_window_bits(void* pippenger);
unsigned bbg_pippenger_levels(void* pippenger);

/* c_bind.cpp:36-43 pippenger_unsafe(pippenger, scalars, from, range, result) ==
 * Pippenger::pippenger_unsafe(scalars, from, range) (pippenger.cpp:27-31): MSM over monomials [from, from+range).
 * scalars: `range` fr elements; result: one g1::element (96 B). */
int bbg_pippenger_unsafe(void* pippenger, const void* scalars, size_t from, size_t range, void* result);
int bbg_pippenger_unsafe_dev(void* pippenger, const void* d_scalars, size_t from, size_t range, void* d_result, void* stream);

/* `count` MSMs over the same monomials [from, from+range) in one call: what work_queue::process_queue
 * (bb/plonk/proof_system/prover/work_queue.hpp:213-243) does item by item for the prover's W_1..W_4 and T_1..T_4
 * commitments.  scalars[i]: `range` fr elements; results: count x 96 B g1::element.  Consecutive MSMs run on up to
 * four streams / workspaces so the latency-bound tail of one overlaps the bucket accumulation of the next ones. */
int bbg_pippenger_unsafe_batch(void* pippenger, const void* const* scalars, size_t count, size_t from, size_t range, void* results);
int bbg_pippenger_unsafe_batch_dev(void* pippenger, const void* const* d_scalars, size_t count, size_t from, size_t range,
                                   void* d_results, void* stream);
/* same, addressed like bbg_pippenger(): by the host address of an adopted 2n interleaved table */
int bbg_pippenger_batch(const void* const* scalars, size_t count, const void* points_table2n, size_t num_points, void* results);

/* ---- resident polynomials (no reference counterpart; bb/plonk/proof_system/prover/work_queue.hpp:208-282 is the caller
 * they exist for).  With residency on, the host-pointer entry points keep the device copy of every array they
 * touch, keyed by host address, and skip the upload when the same array (or a slice of it) comes back: ifft ->
 * commitment MSM -> coset FFT of one wire polynomial crosses PCIe once.  Every mirror carries a fingerprint of the
 * host words it was made from (first / last elements + stratified samples) that is re-checked on each use, so an
 * array the caller rewrote in between is uploaded again.  Off by default (BBG_RESIDENT=1 or bbg_resident_mode(1)). */
int bbg_resident_mode(int enable);                           /* 1 on, 0 off (drops every mirror), -1 query */
int bbg_resident_invalidate(const void* host, size_t bytes); /* forget mirrors overlapping the range (bytes 0: all) */
/* Write mirrors that are ahead of host memory (BBG_KEEP_ON_DEVICE) back to the arrays overlapping [host, host+bytes)
 * (bytes 0: all of them).  The ONLY way, besides an entry point called with that array, that the library writes host
 * memory: the caller vouches that the arrays still exist.  Evictions and bbg_resident_mode(0) never write back. */
int bbg_resident_flush(const void* host, size_t bytes);
int bbg_resident_stats(uint64_t* out4);                      /* hits, misses, H2D bytes saved, bytes resident */

/* bb/.../scalar_multiplication.hpp:139-148  pippenger(scalars, points, num_points, state, handle_edge_cases)
 * and pippenger_unsafe(...).  `points` is the 2n interleaved table (even entries are read).  If it lies inside a
 * table adopted with bbg_new_pippenger_from_table the resident device copy is used, otherwise the points
 * are uploaded for this call.  The runtime state argument of the reference has no device counterpart.
 * Both variants are safe for repeated points and points at infinity. */
int bbg_pippenger(const void* scalars, const void* points_table2n, size_t num_points, int handle_edge_cases, void* result);
/* same with a plain (stride-1) affine array */
int bbg_msm_points(const void* scalars, const void* points, size_t num_points, void* result);
int bbg_msm_points_dev(const void* d_scalars, const void* d_points, size_t point_stride, size_t num_points, void* d_result, void* stream);

/* scalar_multiplication.hpp:94 generate_pippenger_point_table(points, table, num_points); table may alias points */
int bbg_generate_pippenger_point_table(const void* points, void* table, size_t num_points);
/* c_bind.cpp:45-51 g1_sum(points, num_points, result): sum of Jacobian elements (partial-MSM combiner) */
int bbg_g1_sum(const void* elements, size_t num_points, void* result);
int bbg_g1_sum_dev(const void* d_elements, size_t num_points, void* d_result, void* stream);

/* bb/srs/io.hpp:10-18 */
int bbg_read_transcript_g1(void* monomials, size_t degree, const char* srs_dir);
int bbg_read_g1_elements_from_buffer(void* elements, const char* buffer, size_t buffer_size);

/* ---- NTT family (bb/polynomials/polynomial_arithmetic.hpp:23-39), in place, natural order ---- */
#define BBG_FFT 0                              /* fft(coeffs, domain)                          */
#define BBG_IFFT 1                             /* ifft                                         */
#define BBG_COSET_FFT 2                        /* coset_fft(coeffs, domain)                    */
#define BBG_COSET_IFFT 3                       /* coset_ifft                                   */
#define BBG_FFT_WITH_CONSTANT 4                /* fft_with_constant(coeffs, domain, value)     */
#define BBG_IFFT_WITH_CONSTANT 5               /* ifft_with_constant                           */
#define BBG_COSET_FFT_WITH_CONSTANT 6          /* coset_fft_with_constant                      */
#define BBG_COSET_FFT_WITH_GENERATOR_SHIFT 7   /* coset_fft_with_generator_shift               */
/* n = domain.size (power of two, <= 2^28); generator_size = domain.generator_size (0 = n);
 * constant: one fr (Montgomery) for kinds 4..7, ignored otherwise. */
int bbg_ntt(void* coeffs, size_t n, int kind, size_t generator_size, const void* constant);
int bbg_ntt_dev(void* d_coeffs, size_t n, int kind, size_t generator_size, const void* constant, void* stream);
/* Multi-GPU four-step NTT, one process per GPU, world = 2^k ranks, n >= 2^12.  The transform is the same
 * pass factorisation n = n_1 ... n_P as on one GPU, with ONE exchange between pass P-1 and pass P:
 *   phase 0 : d_src = this rank's input sub-array {x[i] : bits [in_pos, in_pos + k) of i == rank}, packed in index order
 *             (n / world elements); runs passes 1..P-1; d_dst (n / world elements, may equal d_src) receives the
 *             intermediate, whose world equal contiguous chunks are the all-to-all send buffers (chunk r -> rank r);
 *   exchange: the caller's all-to-all (NCCL ncclSend/ncclRecv group, torch.distributed.all_to_all_single), chunk from
 *             rank s landing at offset s * n / world^2 of the receive buffer;
 *   phase 1 : d_src = receive buffer, d_dst (distinct) = this rank's output sub-array {X[k] : bits [out_pos, out_pos + k)
 *             of k == rank}, packed in index order.
 * bbg_ntt_dist_layout gives in_pos / out_pos for (n, world).  Same kinds, scalings and bit-exact results as bbg_ntt. */
int bbg_ntt_dist_layout(size_t n, int world, unsigned* in_pos, unsigned* out_pos);
int bbg_ntt_dist_dev(const void* d_src, void* d_dst, size_t n, int kind, size_t generator_size, const void* constant, int rank,
                     int world, int phase, void* stream);

/* bbg_ntt with flags: BBG_KEEP_ON_DEVICE leaves the result in the array's device mirror (resident polynomials on;
 * the host copy is stale until bbg_resident_flush) -- for transforms whose only consumers are further device steps */
#define BBG_KEEP_ON_DEVICE 1u
/* same, but only when the array's mirror is ALREADY ahead of host memory (an earlier step deferred its write-back):
 * a generic entry point can then take part in a device-resident chain without ever hiding data from a host-side caller */
#define BBG_KEEP_IF_AHEAD 2u
int bbg_ntt_ex(void* coeffs, size_t n, int kind, size_t generator_size, const void* constant, unsigned flags);

/* The same with the exchange FUSED into the pass before it: phase 0's last pass stores every element straight into the
 * receive buffer of the rank that owns its chunk, over NVLink peer memory, at the slot the all-to-all would have put it --
 * the transfer overlaps the pass's arithmetic tile by tile and no NCCL all-to-all runs.  peer_recv: `world` device pointers
 * (<= 8), peer_recv[r] = rank r's receive buffer of n / world elements mapped into this process (bbg_peer_buffer_open),
 * peer_recv[rank] = this rank's own (bbg_peer_buffer_alloc); d_work: n / world elements of local scratch.  Before phase 1
 * (bbg_ntt_dist_dev(..., phase = 1) on the own receive buffer) every rank must have finished this call: a stream-ordered
 * all-reduce of one word is enough.  Alternate two receive buffers when transforms follow each other back to back. */
int bbg_ntt_dist_fused_dev(const void* d_src, void* d_work, void* const* peer_recv, size_t n, int kind, size_t generator_size,
                           const void* constant, int rank, int world, void* stream);
/* The multi-GPU transform on NATURAL contiguous blocks, the contract of SURVEY.md 8e: rank q holds x[q n / W, (q + 1) n / W)
 * and ends with X[q n / W, (q + 1) n / W).  All movement is peer memory traffic issued by the passes themselves: phase 0
 * loads its sub-array from the owners' input blocks (peer_in) in its first pass and stores into the owners' receive buffers
 * (peer_recv) in its last; phase 1 transforms peer_recv[rank] and stores every output to the owner of its natural index
 * (peer_out).  Every table holds `world` device pointers (<= 8) to buffers of n / world elements, entry `rank` being this
 * rank's own; d_work: n / world elements of local scratch (phase 0).  The caller orders the phases across ranks with any
 * stream-ordered barrier: all inputs written -> phase 0 everywhere -> phase 1 everywhere -> outputs readable. */
int bbg_ntt_dist_natural_dev(void* const* peer_in, void* d_work, void* const* peer_recv, void* const* peer_out, size_t n, int kind,
                             size_t generator_size, const void* constant, int rank, int world, int phase, void* stream);
/* Peer-mapped device buffers (CUDA IPC, one process per GPU): alloc returns the pointer and a 64-byte handle to send to
 * the other ranks (any transport); open maps another rank's buffer into this process (lazy peer access over NVLink). */
int bbg_peer_buffer_alloc(size_t bytes, void** d_ptr, void* ipc_handle64);
int bbg_peer_buffer_open(const void* ipc_handle64, void** d_ptr);
int bbg_peer_buffer_close(void* d_ptr);
int bbg_peer_buffer_free(void* d_ptr);

/* coset_fft(coeffs, small_domain, large_domain, domain_extension) (polynomial_arithmetic.cpp:401-456):
 * coeffs holds ext*n elements, the first n are the input; output interleaved out[ext*i + k]. */
int bbg_coset_fft_ext(void* coeffs, size_t n, size_t domain_extension);
int bbg_coset_fft_ext_dev(void* d_coeffs, size_t n, size_t domain_extension, void* stream);

/* bb/plonk/proof_system/prover/c_bind.cpp:101-121 */
void* bbg_new_evaluation_domain(size_t circuit_size);
void bbg_delete_evaluation_domain(void* domain);
int bbg_ifft(void* coeffs, void* domain);
int bbg_coset_fft_with_generator_shift(void* coeffs, const void* constant, void* domain);
/* evaluation_domain constants (bb/polynomials/evaluation_domain.cpp:57-76): root, root_inverse, domain,
 * domain_inverse, generator, generator_inverse -- 6 x 32 B, computed by the library's own host arithmetic */
int bbg_domain_constants(size_t n, void* out6);

/* ---- quotient-stage pointwise kernels and scans: the fr arithmetic that sits between the
 * prover's NTTs and MSMs.  Host-pointer semantics like everything above; `flags` = 0 or BBG_KEEP_ON_DEVICE (outputs stay
 * in their device mirror; needs resident polynomials).  All challenge / constant arguments are single fr elements. ---- */
/* work_queue FFT item (bb/plonk/proof_system/prover/work_queue.hpp:260-270): wire_fft[0, ext*n + ext) = the ext*n-point coset
 * FFT of the n coefficients in `wire` (zero padded, generator_size = n), followed by its first `ext` values again
 * (polynomial::add_lagrange_base_coefficient x ext) */
int bbg_wire_coset_fft(const void* wire, void* wire_fft, size_t n, size_t ext, unsigned flags);
/* work_queue IFFT item (work_queue.hpp:272-276): wire <- ifft(wire), in place, n coefficients.  `lagrange_copy` (may be null):
 * a host array whose first n elements the caller has just copied from `wire` (prover.cpp:184-186 keeps the Lagrange-base
 * wires in w_i_fft[0, n) for the permutation widget); with resident polynomials its device mirror is seeded from the data
 * uploaded for the transform instead of being uploaded again in round 3.  flags: 0 or BBG_KEEP_ON_DEVICE (the coefficients
 * stay in the wire's mirror: every later reader of a Turbo proof -- commitment, coset FFT, openings -- is a device step). */
int bbg_wire_ifft(void* wire, size_t n, const void* lagrange_copy, unsigned flags);
/* the IFFT items of one queue flush together (<= 16 columns of n elements; lagrange_copies or its entries may be null): every
 * column is staged into pinned memory by the copy pool and uploaded asynchronously, so the staging of column k + 1 overlaps the
 * upload and transform of column k; one synchronisation at the end */
int bbg_wire_ifft_batch(void* const* wires, size_t n, const void* const* lagrange_copies, size_t count, unsigned flags);
/* TransitionWidget::compute_quotient_contribution (bb/plonk/proof_system/widgets/transition_widgets/transition_widget.hpp:293-307)
 * for the TurboPLONK gate kernels: quotient[i] += identity(i) over the n_large-point coset domain.
 * polys: BBG_NUM_POLYNOMIALS pointers indexed like waffle::PolynomialIndex (types/polynomial_manifest.hpp:10-50) to the
 * "_fft" arrays (n_large elements each; the ones the widget does not read may be null); alpha_base = the widget's first
 * alpha power, alpha = the challenge. */
#define BBG_NUM_POLYNOMIALS 36
#define BBG_WIDGET_TURBO_ARITHMETIC 0   /* turbo_arithmetic_widget.hpp:17-143 */
#define BBG_WIDGET_TURBO_FIXED_BASE 1   /* turbo_fixed_base_widget.hpp:17-160 */
#define BBG_WIDGET_TURBO_RANGE 2        /* turbo_range_widget.hpp:30-161      */
#define BBG_WIDGET_TURBO_LOGIC 3        /* turbo_logic_widget.hpp:17-183      */
int bbg_turbo_quotient(int kind, const void* const* polys, size_t n_large, const void* alpha_base, const void* alpha, void* quotient,
                       unsigned flags);
/* ProverPermutationWidget<program_width, false>::compute_quotient_contribution (widgets/random_widgets/permutation_widget_impl.hpp:317-437):
 * quotient[i] = alpha_base * (numerator - denominator) (assignment), identity permutation polynomials, coset generator 5 */
int bbg_permutation_quotient(const void* const* wire_ffts, const void* const* sigma_ffts, unsigned program_width, const void* z_fft,
                             const void* lagrange_1, size_t n_large, unsigned num_roots_cut, const void* alpha_base, const void* beta,
                             const void* gamma, const void* public_input_delta, void* quotient, unsigned flags);
/* polynomial_arithmetic::divide_by_pseudo_vanishing_polynomial(evaluations, src_domain, target_domain, k) (bb/polynomials/polynomial_arithmetic.cpp:628-725) */
int bbg_divide_by_pseudo_vanishing_polynomial(void* evaluations, size_t n_small, size_t n_large, unsigned num_roots_cut, unsigned flags);
/* polynomial_arithmetic::compute_lagrange_polynomial_fft(l_1, src_domain, target_domain) (:546-626) */
int bbg_compute_lagrange_polynomial_fft(void* l_1_coefficients, size_t n_small, size_t n_large);
/* the grand product of ProverPermutationWidget::compute_round_commitments (permutation_widget_impl.hpp:48-270): z[0] = 1,
 * z[i] = prod_{j<i} prod_k (w_k[j] + gamma + beta k_k w^j) / (w_k[j] + gamma + beta sigma_k[j]), i < n; inputs in Lagrange base */
int bbg_permutation_grand_product(const void* const* wires_lagrange, const void* const* sigmas_lagrange, unsigned program_width, size_t n,
                                  const void* beta, const void* gamma, void* z, unsigned flags);
/* polynomial_arithmetic::evaluate(coeffs, z, n) (:507-538): canonical fr result */
int bbg_evaluate(const void* coeffs, size_t n, const void* z, void* result);
/* `count` (<= 40) evaluations in one launch: results[k] = sum_i polys[k][i] * z_k^i over ns[k] coefficients, zs = count
 * consecutive fr.  KateCommitmentScheme::add_opening_evaluations_to_transcript (kate_commitment_scheme.cpp:373-436) evaluates
 * every polynomial of the manifest at zeta (and some at zeta * omega) one after the other. */
int bbg_evaluate_batch(const void* const* polys, const size_t* ns, size_t count, const void* zs, void* results);
/* KateCommitmentScheme::compute_opening_polynomial / compute_kate_opening_coefficients (commitment_scheme/kate_commitment_scheme.cpp:25-57,
 * polynomial_arithmetic.cpp:727-751): dest[0, n) = coefficients of (F(X) - F(z)) / (X - z), F = src[0, n_eval); *f_at_z = F(z)
 * (may be null); dest may equal src */
int bbg_compute_opening_polynomial(const void* src, void* dest, const void* z, size_t n_eval, size_t n, void* f_at_z, unsigned flags);
/* dest[i] = (base ? base[i] : 0) + sum_k polys[k][i] * scalars[k], i < n; count <= 48; scalars = count consecutive fr.
 * The accumulation of the two opening polynomials in KateCommitmentScheme::batch_open (kate_commitment_scheme.cpp:213-222). */
int bbg_linear_combination(void* dest, const void* base, const void* const* polys, const void* scalars, size_t count, size_t n, unsigned flags);
/* host_array[elem_offset, +count) = values, in host memory and in the array's device mirror (blinding scalars written
 * between two device steps: prover.cpp:181-183, permutation_widget_impl.hpp:289-291) */
int bbg_poly_write(void* host_array, size_t elem_offset, const void* values, size_t count);

/* element-wise probe used by the L0 parity tests: out[i] = op(a[i], b[i]) on the device.
 * field: 0 fq, 1 fr.  op: 0 mul 1 add 2 sub 3 sqr 4 to_montgomery 5 from_montgomery 7 reduce_once 8 neg */
int bbg_field_op(int field, int op, const void* a, const void* b, void* out, size_t n);
int bbg_field_op_dev(int field, int op, const void* d_a, const void* d_b, void* d_out, size_t n, void* stream);
/* g1::affine_element(element) (bb/ecc/groups/element_impl.hpp:51-68) for n Jacobian elements: canonical affine
 * coordinates (what affine_element::to_buffer() serialises), infinity flag kept */
int bbg_g1_normalize(const void* elements, size_t n, void* affine_out);
/* g1 probe: op 0 mixed add (jac, affine) 1 add (jac, jac) 2 dbl (jac); inputs/outputs 96-byte Jacobian */
int bbg_g1_op(int op, const void* a, const void* b, void* out, size_t n);

/* ---- measurement / synthetic-input utilities (no reference counterpart) ------------------------ */
/* Integer-pipe roofline: back-to-back register-resident Montgomery multiplies (field 0 fq, 1 fr) on every SM;
 * writes the sustained rate in multiplies per second.  bench.py reports MSM/NTT arithmetic against it. */
int bbg_bench_field_mul(int field, int iters, double* muls_per_second);
/* out[i] = affine(in[i] + q): distinct synthetic bases beyond the 2^20-point SRS (Q_j = Q_{j-1} + D).
 * in/out: n affine points in device memory (may alias); q: one affine point in host memory. */
int bbg_g1_add_affine_dev(const void* d_in, void* d_out, size_t n, const void* q_affine, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* BBG_H */
