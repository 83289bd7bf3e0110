"""oracle/pypoly.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Plain-Python restatement, over Python integers mod r, of the quotient-stage arithmetic of barretenberg's PLONK prover
(SURVEY.md 8f ranks 2-3); small cases only.  "bb/" = barretenberg/src/aztec/.  Every function cites the reference lines
it follows.  Pinned (tests/test_oracle_poly.py) against the compiled reference where the reference exposes the function
without a proving key (the four Turbo gate kernels through their own templates, divide_by_pseudo_vanishing_polynomial,
compute_lagrange_polynomial_fft, compute_kate_opening_coefficients, evaluate).  The permutation-argument functions need
a proving key and a transcript in the reference; their restatement here is pinned end to end instead: the CUDA kernels
that agree with it produce the byte-identical join-split proof (tests/test_gpu_prover.py).
"""
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
ROOT_28 = None  # set lazily from the oracle (2^28-th primitive root, bb/ecc/curves/bn254/fr.hpp:27-30)
GENERATOR = 5   # fr::coset_generator(0) = evaluation_domain::generator (bb/polynomials/evaluation_domain.cpp:69-70)
COSET_GENERATORS = (5, 6, 7)  # fr::coset_generator(0..2), bb/ecc/curves/bn254/fr.hpp:44-59

# waffle::PolynomialIndex (bb/plonk/proof_system/types/polynomial_manifest.hpp:10-50)
(Q_1, Q_2, Q_3, Q_4, Q_5, Q_M, Q_C, Q_ARITHMETIC_SELECTOR, Q_FIXED_BASE_SELECTOR, Q_RANGE_SELECTOR, Q_SORT_SELECTOR, Q_LOGIC_SELECTOR,
 TABLE_1, TABLE_2, TABLE_3, TABLE_4, TABLE_INDEX, TABLE_TYPE, Q_MIMC_COEFFICIENT, Q_MIMC_SELECTOR, Q_ELLIPTIC, SIGMA_1, SIGMA_2, SIGMA_3,
 SIGMA_4, ID_1, ID_2, ID_3, ID_4, W_1, W_2, W_3, W_4, S, Z, Z_LOOKUP, MAX_NUM_POLYNOMIALS) = range(37)


def inv(a):
    return pow(a, -1, R_MOD)


def root_of_unity(orc, log2n):
    """fr::get_root_of_unity (bb/ecc/fields/field_impl.hpp:496-503) as a plain integer"""
    from oracle import pyoracle as po
    return orc.from_mont_ints(po.FR, orc.fr_root_of_unity(log2n).reshape(1, 4))[0]


def turbo_quotient(kind, polys, n_large, alpha_base, alpha, quotient):
    """TransitionWidget::compute_quotient_contribution (bb/plonk/proof_system/widgets/transition_widgets/transition_widget.hpp:293-307)
    with the gate kernel `kind`; polys: dict PolynomialIndex -> list of ints; returns the new quotient list."""
    p = R_MOD
    mask = n_large - 1
    ap = [alpha_base * pow(alpha, k, p) % p for k in range(7)]
    out = list(quotient)
    for i in range(n_large):
        s = (i + 4) & mask
        w1, w2, w3, w4 = polys[W_1][i], polys[W_2][i], polys[W_3][i], polys[W_4][i]
        w1n, w2n, w3n, w4n = polys[W_1][s], polys[W_2][s], polys[W_3][s], polys[W_4][s]
        if kind == 0:
            # turbo_arithmetic_widget.hpp:17-143
            qa = polys[Q_ARITHMETIC_SELECTOR][i]
            lin = (qa * w1 * w2 * polys[Q_M][i] + qa * w1 * polys[Q_1][i] + qa * w2 * polys[Q_2][i] + qa * w3 * polys[Q_3][i]
                   + qa * w4 * polys[Q_4][i] + (w4 * w4 - w4) * (w4 - 2) * qa * alpha * polys[Q_5][i] + qa * polys[Q_C][i]) * ap[0]
            d = w3 - 4 * w4
            nl = (qa * qa - qa) * d * (9 * d - 2 * d * d - 7) * ap[0]
            v = lin + nl
        elif kind == 1:
            # turbo_fixed_base_widget.hpp:17-160
            qc, qe = polys[Q_C][i], polys[Q_FIXED_BASE_SELECTOR][i]
            d = w4n - 4 * w4
            q1m = d * d * qe * ap[1]
            q2m = ap[1] * qe
            q3m = (w1n - w1) * d * w3n * ap[3] * qe + 2 * d * w3n * w2 * ap[2] * qe
            q4m = w3 * qe * qc * ap[5]
            q5m = (1 - w4) * qe * qc * ap[5]
            qmm = w3 * qe * qc * ap[6]
            lin = (qmm * polys[Q_M][i] + q1m * polys[Q_1][i] + q2m * polys[Q_2][i] + q3m * polys[Q_3][i] + q4m * polys[Q_4][i]
                   + q5m * polys[Q_5][i])
            acc_id = (d + 1) * (d + 3) * (d - 1) * (d - 3) * ap[0]
            x_alpha_id = -(w3n * ap[1])
            t0 = (w1n + w1 + w3n) * (w3n - w1) ** 2
            t1 = -(w3n ** 3 + w2 * w2 + (-17))
            t2 = 2 * d * w2 * qe
            x_acc = (t0 + t1 + t2) * ap[2]
            y_acc = ((w2n + w2) * (w3n - w1) + (w1 - w1n) * (w2 - qe * d)) * ap[3]
            acc_init = (w4 - 1) * (w4 - 1 - w3) * ap[4]
            x_init = -(w1 * w3) * ap[5]
            y_init = ((1 - w4) * qc - w2 * w3) * ap[6]
            gate = ((acc_init + x_init + y_init) * qc + acc_id + x_alpha_id + x_acc + y_acc) * qe
            v = lin + gate
        elif kind == 2:
            # turbo_range_widget.hpp:30-161
            ds = (w3 - 4 * w4, w2 - 4 * w3, w1 - 4 * w2, w4n - 4 * w1)
            acc = 0
            for k, d in enumerate(ds):
                acc += (d * d - d) * (d - 2) * (d - 3) * ap[k]
            v = acc * polys[Q_RANGE_SELECTOR][i]
        else:
            # turbo_logic_widget.hpp:17-183 (the instruction sequence there evaluates exactly this polynomial)
            qc = polys[Q_C][i]
            a, b, c = w1n - 4 * w1, w2n - 4 * w2, w4n - 4 * w4
            s_, q = a + b, a * a + b * b
            ident = (s_ * s_ - q - 2 * w3) * alpha
            ident = (ident + (a * a - a) * (a * a - 5 * a + 6)) * alpha
            ident = (ident + (b * b - b) * (b * b - 5 * b + 6)) * alpha
            inner = (w3 * (4 * w3 - 18 * s_ + 81) + 18 * q - 81 * s_ + 83) * w3
            tail = 3 * c + 3 * s_ - 2 * inner + (9 * c - 3 * s_) * qc
            v = (ident + tail) * ap[0] * polys[Q_LOGIC_SELECTOR][i]
        out[i] = (out[i] + v) % p
    return out


def permutation_quotient(wires, sigmas, z, l1, n_large, roots_cut, alpha_base, beta, gamma, delta, root_large):
    """ProverPermutationWidget<width, false>::compute_quotient_contribution
    (bb/plonk/proof_system/widgets/random_widgets/permutation_widget_impl.hpp:317-437); returns quotient (assignment)."""
    p = R_MOD
    mask = n_large - 1
    width = len(wires)
    out = [0] * n_large
    x = GENERATOR * beta % p  # cur_root_times_beta at i = 0 (:356-359)
    a2 = alpha_base * alpha_base % p
    for i in range(n_large):
        num, den = 1, 1
        for k in range(width):
            wg = (wires[k][i] + gamma) % p
            idt = x if k == 0 else COSET_GENERATORS[k - 1] * x  # :366-384
            num = num * (idt + wg) % p
            den = den * (sigmas[k][i] * beta + wg) % p
        zn = z[(i + 4) & mask]
        num = num * z[i] % p
        den = den * zn % p
        num += (zn - delta) * alpha_base * l1[(i + 4 + 4 * roots_cut) & mask]  # :418-422
        num += (z[i] - 1) * a2 * l1[i]                                         # :424-427
        out[i] = (num - den) * alpha_base % p                                  # :429-430
        x = x * root_large % p
    return out


def grand_product(wires, sigmas, n, beta, gamma, root_small):
    """z of ProverPermutationWidget::compute_round_commitments (permutation_widget_impl.hpp:48-270) before blinding:
    z[0] = 1, z[i + 1] = prod_{j <= i} num_j / den_j for i <= n - 2 (accumulators[0] = &z[1], last index excluded :255-259)"""
    p = R_MOD
    width = len(wires)
    zs = [1] * n
    x = beta  # cur_root_times_beta = w^i beta (:125-127)
    num_acc, den_acc = 1, 1
    for i in range(n - 1):
        for k in range(width):
            wg = (wires[k][i] + gamma) % p
            idt = x if k == 0 else COSET_GENERATORS[k - 1] * x
            num_acc = num_acc * (idt + wg) % p
            den_acc = den_acc * (sigmas[k][i] * beta + wg) % p
        zs[i + 1] = num_acc * inv(den_acc) % p
        x = x * root_small % p
    return zs


def divide_by_pseudo_vanishing_polynomial(evals, n_small, n_large, roots_cut, root_small, root_large):
    """bb/polynomials/polynomial_arithmetic.cpp:628-725"""
    p = R_MOD
    ext = n_large // n_small
    gn = pow(GENERATOR, n_small, p)
    w_ext = pow(root_large, n_small, p)  # the primitive ext-th root get_root_of_unity(log2 ext) (:125): w_{4n}^n
    inv_sub = [inv((gn * pow(w_ext, j, p) - 1) % p) for j in range(ext)]
    winv = inv(root_small)
    consts = [(-pow(winv, k + 1, p)) % p for k in range(roots_cut)]
    out = []
    x = GENERATOR
    for i in range(n_large):
        v = evals[i] * inv_sub[i % ext] % p
        for c in consts:
            v = v * (x + c) % p
        out.append(v)
        x = x * root_large % p
    return out


def lagrange_l1_fft(n_small, n_large, root_large):
    """compute_lagrange_polynomial_fft (polynomial_arithmetic.cpp:546-626)"""
    p = R_MOD
    ext = n_large // n_small
    gn = pow(GENERATOR, n_small, p)
    w_ext = pow(root_large, n_small, p)
    n_inv = inv(n_small)
    numer = [(gn * pow(w_ext, j, p) - 1) * n_inv % p for j in range(ext)]
    out = []
    x = GENERATOR
    for i in range(n_large):
        out.append(inv((x - 1) % p) * numer[i % ext] % p)
        x = x * root_large % p
    return out


def evaluate(coeffs, z):
    """polynomial_arithmetic::evaluate (:507-538)"""
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z + c) % R_MOD
    return acc


def opening_polynomial(src, z, n):
    """compute_kate_opening_coefficients / compute_opening_polynomial (polynomial_arithmetic.cpp:727-751,
    bb/plonk/proof_system/commitment_scheme/kate_commitment_scheme.cpp:25-57): the reference's own recurrence"""
    p = R_MOD
    f = evaluate(src, z)
    d = (-inv(z)) % p
    dest = [0] * n
    dest[0] = (src[0] - f) * d % p
    for i in range(1, n):
        dest[i] = (src[i] - dest[i - 1]) * d % p
    return dest, f
