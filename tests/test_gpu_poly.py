"""GPU parity of the quotient-stage pointwise kernels and scans, through the C-ABI:
against the compiled reference (the Turbo gate kernels through the reference's own templates, the vanishing-polynomial
division, the Lagrange FFT, the opening polynomial, evaluate) and against oracle/pypoly.py (the permutation argument).
Bit-exact on canonical limbs; with and without resident polynomials."""
import numpy as np
import pytest

import inputs
from oracle import pyoracle as po
from oracle import pypoly as pp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bbg():
    import bbg as _bbg
    _bbg.init(0)
    return _bbg


def canon(orc, a):
    return np.array(orc.reduce(po.FR, np.asarray(a).reshape(-1, 4)))


def to_ints(orc, a):
    return orc.from_mont_ints(po.FR, np.asarray(a, dtype=np.uint64).reshape(-1, 4))


def turbo_inputs(seed, n_large):
    ids = [pp.W_1, pp.W_2, pp.W_3, pp.W_4, pp.Q_1, pp.Q_2, pp.Q_3, pp.Q_4, pp.Q_5, pp.Q_M, pp.Q_C, pp.Q_ARITHMETIC_SELECTOR,
           pp.Q_FIXED_BASE_SELECTOR, pp.Q_RANGE_SELECTOR, pp.Q_LOGIC_SELECTOR]
    return {k: inputs.fr_elements(seed + k, n_large, coarse_fraction=0.2) for k in ids}


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("lg", [6, 14])
def test_turbo_quotient_vs_reference_templates(bbg, orc, ref, kind, lg):
    n_large = 1 << lg
    polys = turbo_inputs(1000 * kind + lg, n_large)
    a0, a = inputs.fr_elements(7, 1)[0], inputs.fr_elements(8, 1)[0]
    q0 = inputs.fr_elements(9, n_large)
    want = ref.turbo_quotient(kind, polys, n_large, a0, a, q0)
    got = bbg.turbo_quotient(kind, polys, n_large, a0, a, q0.copy())
    assert np.array_equal(canon(orc, got), canon(orc, want))


def test_turbo_quotient_missing_polynomial_is_an_error(bbg):
    polys = turbo_inputs(1, 64)
    del polys[pp.Q_RANGE_SELECTOR]
    with pytest.raises(bbg.BbgError):
        bbg.turbo_quotient(bbg.WIDGET_TURBO_RANGE, polys, 64, inputs.fr_elements(1, 1)[0], inputs.fr_elements(2, 1)[0], np.zeros((64, 4), dtype=np.uint64))


@pytest.mark.parametrize("lg_small,ext", [(4, 4), (10, 4), (12, 2), (5, 1)])
def test_divide_by_pseudo_vanishing_polynomial_vs_reference(bbg, orc, ref, lg_small, ext):
    n, N = 1 << lg_small, (1 << lg_small) * ext
    ev = inputs.fr_elements(21 + lg_small, N, coarse_fraction=0.3)
    want = ref.divide_by_pseudo_vanishing_polynomial(ev, n, 4)
    got = bbg.divide_by_pseudo_vanishing_polynomial(ev.copy(), n, 4)
    assert np.array_equal(canon(orc, got), canon(orc, want))


@pytest.mark.parametrize("lg_small,ext", [(4, 4), (10, 4), (11, 2)])
def test_lagrange_polynomial_fft_vs_reference(bbg, orc, ref, lg_small, ext):
    n, N = 1 << lg_small, (1 << lg_small) * ext
    assert np.array_equal(canon(orc, bbg.compute_lagrange_polynomial_fft(n, N)), canon(orc, ref.compute_lagrange_polynomial_fft(n, N)))


@pytest.mark.parametrize("n", [1, 5, 16, 17, 4096, 65536 + 3])
def test_evaluate_vs_reference(bbg, orc, ref, n):
    c = inputs.fr_elements(40 + n % 97, n, coarse_fraction=0.2)
    z = inputs.fr_elements(41, 1)[0]
    assert np.array_equal(canon(orc, bbg.evaluate(c, z)), canon(orc, ref.evaluate(c, z)))


@pytest.mark.parametrize("n", [2, 16, 33, 4096, 40000])
def test_opening_polynomial_vs_reference(bbg, orc, ref, n):
    src = inputs.fr_elements(50 + n % 89, n)
    z = inputs.fr_elements(51, 1)[0]
    want_d, want_f = ref.compute_kate_opening_coefficients(src, z)
    got_d, got_f = bbg.compute_opening_polynomial(src.copy(), z)
    assert np.array_equal(canon(orc, got_f), canon(orc, want_f))
    assert np.array_equal(canon(orc, got_d), canon(orc, want_d))
    # in place (dest == src), as the reference's polynomial::compute_kate_opening_coefficients calls it
    buf = src.copy()
    bbg.compute_opening_polynomial(buf, z, dest=buf)
    assert np.array_equal(canon(orc, buf), canon(orc, want_d))


def perm_inputs(seed, n):
    wires = [inputs.fr_elements(seed + k, n) for k in range(4)]
    sigmas = [inputs.fr_elements(seed + 10 + k, n) for k in range(4)]
    return wires, sigmas


@pytest.mark.parametrize("width", [3, 4])
def test_permutation_quotient_vs_pypoly(bbg, orc, width):
    n_large, cut = 256, 4
    wires, sigmas = perm_inputs(60, n_large)
    wires, sigmas = wires[:width], sigmas[:width]
    z, l1 = inputs.fr_elements(71, n_large), inputs.fr_elements(72, n_large)
    a0, beta, gamma, delta = (inputs.fr_elements(73 + k, 1)[0] for k in range(4))
    q = inputs.fr_elements(77, n_large)  # overwritten: assignment, not accumulation
    got = bbg.permutation_quotient(wires, sigmas, z, l1, n_large, cut, a0, beta, gamma, delta, q)
    I = lambda x: to_ints(orc, x)  # noqa: E731
    want = pp.permutation_quotient([I(w) for w in wires], [I(s) for s in sigmas], I(z), I(l1), n_large, cut, I(a0)[0], I(beta)[0], I(gamma)[0],
                                   I(delta)[0], pp.root_of_unity(orc, 8))
    assert I(got) == want


@pytest.mark.parametrize("n,width", [(16, 4), (64, 3), (1024, 4)])
def test_grand_product_vs_pypoly(bbg, orc, n, width):
    wires, sigmas = perm_inputs(80 + n, n)
    wires, sigmas = wires[:width], sigmas[:width]
    beta, gamma = inputs.fr_elements(91, 1)[0], inputs.fr_elements(92, 1)[0]
    got = bbg.permutation_grand_product(wires, sigmas, n, beta, gamma)
    I = lambda x: to_ints(orc, x)  # noqa: E731
    lg = n.bit_length() - 1
    want = pp.grand_product([I(w) for w in wires], [I(s) for s in sigmas], n, I(beta)[0], I(gamma)[0], pp.root_of_unity(orc, lg))
    assert I(got) == want


def test_grand_product_is_a_permutation_check(bbg, orc):
    """semantic property (the reason the argument exists): when sigma is the identity permutation of the SAME values the
    running product telescopes, i.e. z[i] = 1 wherever the prefix sets coincide.  sigma_k[i] = k_k w^i makes num == den."""
    n = 256
    root = pp.root_of_unity(orc, 8)
    wires = [inputs.fr_elements(95 + k, n) for k in range(4)]
    ks = [1, 5, 6, 7]
    sig_ints = [[ks[k] * pow(root, i, pp.R_MOD) % pp.R_MOD for i in range(n)] for k in range(4)]
    sigmas = [orc.to_mont(po.FR, s) for s in sig_ints]
    z = bbg.permutation_grand_product(wires, sigmas, n, inputs.fr_elements(96, 1)[0], inputs.fr_elements(97, 1)[0])
    assert to_ints(orc, z) == [1] * n


def test_wire_coset_fft_matches_the_work_queue_sequence(bbg, orc, ref):
    """work_queue.hpp:260-270: copy n coefficients, zero-pad to 4n + 4, coset_fft over the 4n domain with generator_size n,
    then append the first four evaluations"""
    n = 1 << 10
    wire = inputs.fr_elements(101, n)
    out = np.zeros((4 * n + 4, 4), dtype=np.uint64)
    bbg.wire_coset_fft(wire, out, n, 4)
    padded = np.zeros((4 * n, 4), dtype=np.uint64)
    padded[:n] = wire
    want = ref.ntt(po.NTT_COSET_FFT, padded, generator_size=n)
    assert np.array_equal(canon(orc, out[: 4 * n]), canon(orc, want))
    assert np.array_equal(canon(orc, out[4 * n:]), canon(orc, want[:4]))


def test_device_resident_round_four_chain(bbg, orc, ref):
    """The chain the prover shim runs with resident polynomials: wire coset FFTs kept on the device -> permutation
    quotient (assignment, kept) -> a gate widget (kept because the mirror is ahead) -> division (kept) -> coset_ifft
    (written back).  Host arrays in between are stale by design; the final result must equal the host-side chain."""
    n, N = 1 << 8, 1 << 10
    polys = turbo_inputs(300, N)
    wires_coeff = [inputs.fr_elements(310 + k, n) for k in range(4)]
    sigmas = [inputs.fr_elements(320 + k, N) for k in range(4)]
    z, l1 = inputs.fr_elements(330, N), inputs.fr_elements(331, N)
    a0, a, beta, gamma, delta = (inputs.fr_elements(340 + k, 1)[0] for k in range(5))

    def chain(flags_keep, flags_ahead):
        wf = [np.zeros((N + 4, 4), dtype=np.uint64) for _ in range(4)]
        for k in range(4):
            bbg.wire_coset_fft(wires_coeff[k], wf[k], n, 4, flags_keep)
        p = dict(polys)
        for k, idx in enumerate((pp.W_1, pp.W_2, pp.W_3, pp.W_4)):
            p[idx] = wf[k]
        q = np.zeros((N, 4), dtype=np.uint64)
        bbg.permutation_quotient(wf, sigmas, z, l1, N, 4, a0, beta, gamma, delta, q, flags_keep)
        bbg.turbo_quotient(bbg.WIDGET_TURBO_ARITHMETIC, p, N, a0, a, q, flags_ahead)
        bbg.turbo_quotient(bbg.WIDGET_TURBO_RANGE, p, N, a0, a, q, flags_ahead)
        bbg.divide_by_pseudo_vanishing_polynomial(q, n, 4, flags_ahead)
        bbg.coset_ifft(q)
        return q

    plain = chain(0, 0)
    bbg.resident_mode(True)
    try:
        s0 = bbg.resident_stats()
        kept = chain(bbg.KEEP_ON_DEVICE, bbg.KEEP_IF_AHEAD)
        assert bbg.resident_stats()["hits"] > s0["hits"] + 8
        again = chain(bbg.KEEP_ON_DEVICE, bbg.KEEP_IF_AHEAD)  # second "proof": mirrors of the first are reused or refreshed
    finally:
        bbg.resident_mode(False)
    assert np.array_equal(canon(orc, kept), canon(orc, plain))
    assert np.array_equal(canon(orc, again), canon(orc, plain))


def test_poly_write_reaches_host_and_mirror(bbg, orc, srs_mini):
    n = 1 << 10
    bbg.resident_mode(True)
    try:
        z = inputs.fr_elements(400, n)
        bbg.ntt_ex(z, bbg.FFT, flags=bbg.KEEP_ON_DEVICE)      # mirror ahead of host
        blind = inputs.fr_elements(401, 3)
        bbg.poly_write(z, n - 3, blind)
        bbg.ifft(z)                                            # from the mirror, written back
    finally:
        bbg.resident_mode(False)
    want = inputs.fr_elements(400, n)
    bbg.fft(want)
    want[n - 3:] = blind
    bbg.ifft(want)
    assert np.array_equal(canon(orc, z), canon(orc, want))


@pytest.mark.parametrize("count,with_base", [(1, False), (5, True), (26, True), (48, False)])
def test_linear_combination_vs_field_ops(bbg, orc, count, with_base):
    """kate_commitment_scheme.cpp:213-222: opening_poly[i] = t[i] + sum_k poly_k[i] * nu_k"""
    n = 1 << 10
    polys = [inputs.fr_elements(500 + k, n, coarse_fraction=0.1) for k in range(count)]
    sc = inputs.fr_elements(600, count)
    base = inputs.fr_elements(601, n) if with_base else None
    got = bbg.linear_combination(polys, sc, n, base=base)
    want = base.copy() if with_base else np.zeros((n, 4), dtype=np.uint64)
    for k in range(count):
        term = orc.field_op(po.FR, po.OP_MUL, polys[k], np.repeat(sc[k:k + 1], n, axis=0))
        want = orc.field_op(po.FR, po.OP_ADD, want, term)
    assert np.array_equal(canon(orc, got), canon(orc, want))
    # in place: dest is also the base
    if with_base:
        buf = base.copy()
        bbg.linear_combination(polys, sc, n, base=buf, dest=buf)
        assert np.array_equal(canon(orc, buf), canon(orc, want))


def test_evaluate_batch_vs_reference(bbg, orc, ref):
    """round 5's opening evaluations in one launch: different lengths and different points per polynomial"""
    sizes = [4096, 4096, 1 << 14, 33, 4096, 1]
    polys = [inputs.fr_elements(700 + k, n, coarse_fraction=0.1) for k, n in enumerate(sizes)]
    zs = inputs.fr_elements(710, len(sizes))
    got = bbg.evaluate_batch(polys, zs)
    for k in range(len(sizes)):
        assert np.array_equal(canon(orc, got[k]), canon(orc, ref.evaluate(polys[k], zs[k]))), k


def test_wire_ifft_seeds_the_lagrange_mirror(bbg, orc, ref):
    """the IFFT work item keeps the Lagrange-base copy on the device for round 3's grand product"""
    n = 1 << 10
    lag = [inputs.fr_elements(800 + k, n) for k in range(4)]
    sigmas = [inputs.fr_elements(810 + k, n) for k in range(4)]
    beta, gamma = inputs.fr_elements(820, 1)[0], inputs.fr_elements(821, 1)[0]
    want_z = bbg.permutation_grand_product(lag, sigmas, n, beta, gamma)
    bbg.resident_mode(True)
    try:
        wires = [w.copy() for w in lag]
        copies = [np.zeros((4 * n + 4, 4), dtype=np.uint64) for _ in range(4)]
        for k in range(4):
            copies[k][:n] = wires[k]           # prover.cpp:184-186
            bbg.wire_ifft(wires[k], copies[k])
        s0 = bbg.resident_stats()
        z = bbg.permutation_grand_product([c[:n] for c in copies], sigmas, n, beta, gamma)
        assert bbg.resident_stats()["hits"] >= s0["hits"] + 4   # the four Lagrange copies were not uploaded again
        for k in range(4):
            assert np.array_equal(canon(orc, wires[k]), canon(orc, ref.ntt(po.NTT_IFFT, lag[k])))
    finally:
        bbg.resident_mode(False)
    assert np.array_equal(canon(orc, z), canon(orc, want_z))


@pytest.mark.parametrize("resident,keep", [(False, False), (True, False), (True, True)])
def test_wire_ifft_batch_matches_single_iffts(bbg, orc, ref, resident, keep):
    """bbg_wire_ifft_batch (the IFFT items of one queue flush: columns staged into pinned memory by the copy pool, uploaded
    asynchronously, one synchronisation): every column must be the reference's ifft, the Lagrange copies' mirrors must be
    seeded, and with BBG_KEEP_ON_DEVICE the coefficients must be readable by the next device step (a commitment-style
    evaluate) and come home on bbg_resident_flush."""
    n = 1 << 12
    lag = [inputs.fr_elements(1800 + k, n) for k in range(5)]
    want = [ref.ntt(po.NTT_IFFT, x) for x in lag]
    zeta = inputs.fr_elements(1810, 1)[0]
    bbg.resident_mode(resident)
    try:
        wires = [w.copy() for w in lag]
        copies = [np.zeros((4 * n + 4, 4), dtype=np.uint64) for _ in range(5)]
        for k in range(5):
            copies[k][:n] = wires[k]
        copies[4] = None  # an item without a Lagrange copy
        bbg.wire_ifft_batch(wires, copies, bbg.KEEP_ON_DEVICE if keep else 0)
        if keep:
            for k in range(5):
                assert np.array_equal(canon(orc, bbg.evaluate(wires[k], zeta)), canon(orc, ref.evaluate(want[k], zeta))), k
                bbg.resident_flush(wires[k])
        for k in range(5):
            assert np.array_equal(canon(orc, wires[k]), canon(orc, want[k])), k
        if resident:
            s0 = bbg.resident_stats()
            sigmas = [inputs.fr_elements(1820 + k, n) for k in range(4)]
            beta, gamma = inputs.fr_elements(1830, 1)[0], inputs.fr_elements(1831, 1)[0]
            z = bbg.permutation_grand_product([c[:n] for c in copies[:4]], sigmas, n, beta, gamma)
            assert bbg.resident_stats()["hits"] >= s0["hits"] + 4
            bbg.resident_mode(False)
            assert np.array_equal(canon(orc, z), canon(orc, bbg.permutation_grand_product(lag[:4], sigmas, n, beta, gamma)))
    finally:
        bbg.resident_mode(False)
