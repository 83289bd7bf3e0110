"""The reference's own test cases for the hot path, replayed against the CUDA library through the C-ABI.

One test per TEST() of
  bb/ecc/curves/bn254/scalar_multiplication/scalar_multiplication.test.cpp   (:619-927, the pippenger cases; the earlier
      cases exercise CPU-internal helpers -- reduce_buckets, add_affine_points, radix_sort -- that have no device counterpart)
  bb/polynomials/polynomial_arithmetic.test.cpp                                (the fft / ifft / coset cases)
with the same sizes and the same assertion, the expectation computed the way the reference test computes it (naive
sum of scalar multiplications, polynomial evaluation at the roots) by the oracle.  The MSM calls go through
bbg_pippenger with a 2n interleaved table, i.e. exactly the entry point barretenberg's signature maps to.
"""
import numpy as np
import pytest

import inputs
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bbg():
    import bbg as _bbg
    _bbg.init(0)
    return _bbg


def canon(orc, a):
    return np.array(orc.reduce(po.FR, np.asarray(a).reshape(-1, 4)))


def random_points(orc, srs_mini, seed, n):
    """Distinct 'random' affine points: k_i * P_i for SRS points P_i and seeded scalars k_i (g1::element::random_element
    is an RNG draw in the reference; any points do)."""
    pts, _ = srs_mini
    ks = inputs.fr_elements(seed, n)
    return np.stack([orc.g1_to_affine(orc.g1_mul(pts[i + 1], ks[i])) for i in range(n)])


def naive(orc, scalars, points):
    return orc.jac_to_buffer(orc.naive_msm(scalars, points, stride=1))


def run_pippenger(bbg, orc, scalars, points, unsafe):
    table = orc.point_table(points)  # generate_pippenger_point_table(points, points, n)
    n = scalars.shape[0]
    out = bbg.pippenger_unsafe(scalars, table, n) if unsafe else bbg.pippenger(scalars, table, n, True)
    return orc.jac_to_buffer(out)


# ----------------------------------------------------------------------------- scalar_multiplication.test.cpp
def test_undersized_inputs(bbg, orc, srs_mini):
    """:619-653 -- 17 points (the CPU falls back to naive multiplication there)."""
    pts = random_points(orc, srs_mini, 1, 17)
    sc = inputs.fr_elements(2, 17)
    assert run_pippenger(bbg, orc, sc, pts, unsafe=False) == naive(orc, sc, pts)


def test_pippenger(bbg, orc, srs_mini):
    """:655-686 -- random scalars, random points (reference: 2^13 points; here 2^11 because the expectation is the naive
    sum computed by the scalar C oracle)."""
    n = 1 << 11
    pts = random_points(orc, srs_mini, 3, n)
    sc = inputs.fr_elements(4, n)
    assert run_pippenger(bbg, orc, sc, pts, unsafe=False) == naive(orc, sc, pts)


def test_pippenger_edge_case_dbl(bbg, orc, srs_mini):
    """:688-721 -- 128 copies of ONE point: every bucket addition is a doubling."""
    p = random_points(orc, srs_mini, 5, 1)[0]
    pts = np.repeat(p[None, :], 128, axis=0)
    sc = inputs.fr_elements(6, 128)
    assert run_pippenger(bbg, orc, sc, pts, unsafe=False) == naive(orc, sc, pts)


def short_scalars(orc, seed, n):
    return orc.field_op(po.FR, po.OP_TO_MONT, po.ints_to_array(inputs.short_scalar_ints(seed, n)))


@pytest.mark.parametrize("unsafe", [False, True])
def test_pippenger_short_inputs(bbg, orc, srs_mini, unsafe):
    """:723-774 (safe) and :808-860 (unsafe) -- the 128-bit / 64-bit / 3-bit / zero / one scalar mix."""
    n = 1 << 10
    pts = random_points(orc, srs_mini, 7, n)
    sc = short_scalars(orc, 8, n)
    assert run_pippenger(bbg, orc, sc, pts, unsafe=unsafe) == naive(orc, sc, pts)


def test_pippenger_unsafe(bbg, orc, srs_mini):
    """:776-806."""
    n = 1 << 11
    pts = random_points(orc, srs_mini, 9, n)
    sc = inputs.fr_elements(10, n)
    assert run_pippenger(bbg, orc, sc, pts, unsafe=True) == naive(orc, sc, pts)


def test_pippenger_one(bbg, orc, srs_mini):
    """:862-893 -- a single point."""
    pts = random_points(orc, srs_mini, 11, 1)
    sc = inputs.fr_elements(12, 1)
    assert run_pippenger(bbg, orc, sc, pts, unsafe=False) == naive(orc, sc, pts)


def test_pippenger_zero_points(bbg, orc, srs_mini):
    """:895-908 -- num_points = 0 returns the point at infinity."""
    pts = random_points(orc, srs_mini, 13, 1)
    table = orc.point_table(pts)
    out = bbg.pippenger(np.zeros((0, 4), np.uint64), table, 0, True)
    assert orc.jac_to_buffer(out) == orc.jac_to_buffer(orc.g1_infinity())


def test_pippenger_mul_by_zero(bbg, orc, srs_mini):
    """:910-927 -- one point, scalar zero."""
    pts = random_points(orc, srs_mini, 14, 1)
    out = run_pippenger(bbg, orc, np.zeros((1, 4), np.uint64), pts, unsafe=False)
    assert out == orc.jac_to_buffer(orc.g1_infinity())


# ----------------------------------------------------------------------------- polynomial_arithmetic.test.cpp
def roots_pow(orc, lg, count):
    w = orc.fr_root_of_unity(lg)
    cur = orc.to_mont(po.FR, [1])[0]
    out = []
    for _ in range(count):
        out.append(cur)
        cur = orc.field_op(po.FR, po.OP_MUL, cur, w)[0]
    return out


def test_fft_with_small_degree(bbg, orc):
    """fft of 16 coefficients == evaluate(poly, w^i) for every i."""
    n = 16
    poly = inputs.fr_elements(21, n)
    got = canon(orc, bbg.fft(poly.copy()))
    for i, z in enumerate(roots_pow(orc, 4, n)):
        assert np.array_equal(got[i], canon(orc, orc.evaluate(poly, z))[0]), i


def test_basic_fft(bbg, orc):
    """ifft(fft(x)) == x at n = 2^14."""
    x = inputs.fr_elements(22, 1 << 14)
    assert np.array_equal(canon(orc, bbg.ifft(bbg.fft(x.copy()))), x)


@pytest.mark.parametrize("n", [2, 256, 1 << 14])
def test_fft_ifft_consistency(bbg, orc, n):
    x = inputs.fr_elements(23 + n, n)
    assert np.array_equal(canon(orc, bbg.ifft(bbg.fft(x.copy()))), x)


@pytest.mark.parametrize("n", [2, 256, 1 << 14])
def test_fft_coset_ifft_consistency(bbg, orc, n):
    """coset_ifft(coset_fft(x)) == x, and the coset evaluation equals evaluate(poly, g w^i) at sampled i."""
    x = inputs.fr_elements(24 + n, n)
    y = bbg.coset_fft(x.copy())
    assert np.array_equal(canon(orc, bbg.coset_ifft(y.copy())), x)
    lg = n.bit_length() - 1
    g = orc.to_mont(po.FR, [5])[0]  # fr::coset_generator(0)
    roots = roots_pow(orc, lg, min(n, 3))
    for i, w in enumerate(roots):
        z = orc.field_op(po.FR, po.OP_MUL, g, w)[0]
        assert np.array_equal(canon(orc, y)[i], canon(orc, orc.evaluate(x, z))[0]), i


def test_fft_coset_ifft_cross_consistency(bbg, orc):
    """n = 2: coset FFTs of the same zero-padded polynomial on domains n, 2n, 4n agree on the shared points
    (a[i] + b[2i] + c[4i] == 3 * coset_fft(a)[i]); coset_ifft of the sum returns 3x."""
    n = 2
    x = inputs.fr_elements(25, n)

    def padded(m):
        out = np.zeros((m, 4), dtype=np.uint64)
        out[:n] = x
        return out

    a = bbg.coset_fft(padded(n))
    b = bbg.coset_fft(padded(2 * n))
    c = bbg.coset_fft(padded(4 * n))
    s = orc.field_op(po.FR, po.OP_ADD, orc.field_op(po.FR, po.OP_ADD, a, c[::4].copy()), b[::2].copy())
    back = canon(orc, bbg.coset_ifft(np.ascontiguousarray(s)))
    three_x = canon(orc, orc.field_op(po.FR, po.OP_ADD, orc.field_op(po.FR, po.OP_ADD, x, x), x))
    assert np.array_equal(back, three_x)


def test_compute_lagrange_polynomial_fft_building_block(bbg, orc):
    """compute_lagrange_polynomial_fft (polynomial_arithmetic.cpp:572-650) is built on a coset FFT over the 2n domain of
    the n-th roots' Lagrange polynomial; the device-side building block it needs is coset_fft with
    generator_size < domain size (scale_by_generator only touches the first generator_size coefficients)."""
    n = 1 << 10
    x = np.zeros((4 * n, 4), dtype=np.uint64)
    x[:n] = inputs.fr_elements(26, n)
    got = canon(orc, bbg.coset_fft(x.copy(), generator_size=n))
    exp = canon(orc, orc.ntt(po.NTT_COSET_FFT, x, generator_size=n))
    assert np.array_equal(got, exp)
    # zero padding means scaling all 4n coefficients gives the same result
    assert np.array_equal(got, canon(orc, bbg.coset_fft(x.copy())))
