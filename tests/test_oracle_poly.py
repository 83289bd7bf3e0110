"""CPU: pins oracle/pypoly.py (plain-Python restatement of the prover's quotient-stage arithmetic) to the compiled
reference wherever the reference exposes the function without a proving key."""
import numpy as np
import pytest

import inputs
from oracle import pyoracle as po
from oracle import pypoly as pp

pytestmark = pytest.mark.ref


def to_ints(orc, a):
    return orc.from_mont_ints(po.FR, np.asarray(a, dtype=np.uint64).reshape(-1, 4))


def to_mont(orc, ints):
    return orc.to_mont(po.FR, [int(x) % pp.R_MOD for x in ints])


def turbo_inputs(seed, n_large):
    ids = [pp.W_1, pp.W_2, pp.W_3, pp.W_4, pp.Q_1, pp.Q_2, pp.Q_3, pp.Q_4, pp.Q_5, pp.Q_M, pp.Q_C, pp.Q_ARITHMETIC_SELECTOR,
           pp.Q_FIXED_BASE_SELECTOR, pp.Q_RANGE_SELECTOR, pp.Q_LOGIC_SELECTOR]
    return {k: inputs.fr_elements(seed + k, n_large, coarse_fraction=0.2) for k in ids}


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_turbo_kernels_match_the_reference_templates(orc, ref, kind):
    n_large = 64
    polys = turbo_inputs(100 * kind, n_large)
    a0, a = inputs.fr_elements(7, 1)[0], inputs.fr_elements(8, 1)[0]
    q0 = inputs.fr_elements(9, n_large)
    want = to_ints(orc, ref.turbo_quotient(kind, polys, n_large, a0, a, q0))
    ip = {k: to_ints(orc, v) for k, v in polys.items()}
    got = pp.turbo_quotient(kind, ip, n_large, to_ints(orc, a0)[0], to_ints(orc, a)[0], to_ints(orc, q0))
    assert got == want


def test_divide_by_pseudo_vanishing_polynomial(orc, ref):
    n, N = 16, 64
    ev = inputs.fr_elements(21, N)
    want = to_ints(orc, ref.divide_by_pseudo_vanishing_polynomial(ev, n, 4))
    got = pp.divide_by_pseudo_vanishing_polynomial(to_ints(orc, ev), n, N, 4, pp.root_of_unity(orc, 4), pp.root_of_unity(orc, 6))
    assert got == want


def test_lagrange_polynomial_fft(orc, ref):
    n, N = 16, 64
    want = to_ints(orc, ref.compute_lagrange_polynomial_fft(n, N))
    assert pp.lagrange_l1_fft(n, N, pp.root_of_unity(orc, 6)) == want


def test_kate_opening_coefficients_and_evaluate(orc, ref):
    n = 64
    src = inputs.fr_elements(31, n)
    z = inputs.fr_elements(32, 1)[0]
    d, f = ref.compute_kate_opening_coefficients(src, z)
    dest, fz = pp.opening_polynomial(to_ints(orc, src), to_ints(orc, z)[0], n)
    assert to_ints(orc, f)[0] == fz == to_ints(orc, ref.evaluate(src, z))[0]
    assert to_ints(orc, d) == dest
    # the recurrence the reference runs from the constant term equals the quotient's suffix sums (what the device computes)
    ints, zi = to_ints(orc, src), to_ints(orc, z)[0]
    for i in (0, 1, n // 2, n - 2, n - 1):
        assert dest[i] == sum(ints[j] * pow(zi, j - i - 1, pp.R_MOD) for j in range(i + 1, n)) % pp.R_MOD


def test_coset_generators_are_5_6_7(ref):
    """bb/ecc/curves/bn254/fr.hpp:44-59: coset_generator(0..2), used by the permutation argument"""
    o = po.Oracle()
    assert list(o.from_mont_ints(po.FR, np.array([[0x5eef048d8fffffe7, 0x12ee50ec1ce401d0, 0x29312d5a5e5ee7, 0x463456c802275bed],
                                                  [0xb8538a9dfffffe2, 0x49eac781bc44cefa, 0x6697d49cd2d7a515, 0x543ece899c2f3b1c],
                                                  [0x3057819e4fffffdb, 0x307f6d866832bb01, 0x5c65ec9f484e3a89, 0x180a96573d3d9f8]], dtype=np.uint64))) == [5, 6, 7]
