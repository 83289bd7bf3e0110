"""Pins the plain-C oracle's MSM and NTT against golden vectors generated from the unmodified
reference, against the reference's own property tests, and (when present) the reference .so live."""
import numpy as np
import pytest

import inputs
from helpers import msm_scalars, unhex
from oracle import pyoracle as po


def test_msm_golden(orc, golden, srs_mini):
    pts, table = srs_mini
    for case in golden["msm"]:
        sc = msm_scalars(case, orc)
        assert "%016x" % po.fnv1a64(sc) == case["fnv_scalars"], case
        if case["kind"] == "repeated_point7":
            rep = po.aligned_empty((case["n"], 8))
            rep[:] = pts[7]
            res = orc.pippenger(sc, rep, stride=1)
        else:
            res = orc.pippenger(sc, table, n=case["n"], stride=2)
        assert orc.jac_to_buffer(res).hex() == case["result"], case


# scalar_multiplication.test.cpp:655-686: pippenger == naive sum of points[i] * scalars[i]
def test_msm_matches_naive(orc, srs_mini):
    pts, table = srs_mini
    sc = inputs.fr_elements(77, 200)
    assert orc.jac_to_buffer(orc.pippenger(sc, table)) == orc.jac_to_buffer(orc.naive_msm(sc, table))


def test_ntt_golden(orc, golden):
    for e in golden["ntt"]:
        n = 1 << e["log2n"]
        x = inputs.fr_elements(e["seed"], n, coarse_fraction=0.25)
        const = inputs.fr_elements(e["const_seed"], 1)[0]
        y = orc.reduce(po.FR, orc.ntt(e["kind"], x, generator_size=e["generator_size"], constant=const))
        assert "%016x" % po.fnv1a64(y) == e["fnv"], e
        if "full" in e:
            assert np.array_equal(y, unhex(e["full"]))


def test_coset_fft_ext_golden(orc, golden):
    for e in golden["coset_fft_ext"]:
        n = 1 << e["log2n"]
        x = inputs.fr_elements(e["seed"], n)
        y = orc.reduce(po.FR, orc.coset_fft_ext(x, n, e["ext"]))
        assert "%016x" % po.fnv1a64(y) == e["fnv"], e


# polynomial_arithmetic.test.cpp:45-68 fft_with_small_degree: fft output i == evaluate(poly, w^i)
def test_fft_with_small_degree(orc):
    n = 16
    x = inputs.fr_elements(5, n)
    y = orc.reduce(po.FR, orc.ntt(po.NTT_FFT, x))
    root = orc.domain_constants(n)[0]
    w = orc.to_mont(po.FR, [1])[0]
    for i in range(n):
        assert np.array_equal(orc.reduce(po.FR, orc.evaluate(x, w).reshape(1, 4))[0], y[i])
        w = orc.field_op(po.FR, po.OP_MUL, w.reshape(1, 4), root.reshape(1, 4))[0]


# polynomial_arithmetic.test.cpp:70-134: fft∘ifft = id, coset_fft∘coset_ifft = id
@pytest.mark.parametrize("lg", [1, 3, 10, 14])
def test_fft_ifft_consistency(orc, lg):
    x = inputs.fr_elements(6 + lg, 1 << lg)
    canon = orc.reduce(po.FR, x)
    assert np.array_equal(orc.reduce(po.FR, orc.ntt(po.NTT_IFFT, orc.ntt(po.NTT_FFT, x))), canon)
    assert np.array_equal(orc.reduce(po.FR, orc.ntt(po.NTT_COSET_IFFT, orc.ntt(po.NTT_COSET_FFT, x))), canon)


# ---- staging check: the compiled reference reproduces SURVEY.md Appendix B (g++ build)
APPENDIX_B = {
    10: dict(s0="0f4783413dc3aad223503b8f25f251743c5f91377d55f414e6c9f8ce4b1815fc", scalars="59919e101ea0e3fc",
             msm="0aae9f65e29ca851dc7b0a2f5d5d2e9beae59420bb64ece41ea4326abb16689f1052e5bc099d2887eb5e66295802224cd47c8ede9f2be0716150d970881db19c",
             fft="5c4d50c7fbeec8bc", coset_fft="db4ea786e2f88503", ifft="8e48521830a3f8a2"),
    16: dict(s0="0f4783413dc3aad223503b8f25f251743c5f91377d55f414e6c9f8ce4b1815fc", scalars="cb1450069600a63f",
             msm="285e9c258cced5b63d573391bc95ddbf2bb5e39a99fa514afa0e79fa9e72f5cb10674d099c5eb0abe601b1a8e0103ed82e9c1e10561fbeaa7c3f9e0c8b9a744a",
             fft="9c69b08724dfe4de", coset_fft="5ccc75f538a37ceb", ifft="8bea2726fa48e23d"),
}


@pytest.mark.ref
@pytest.mark.parametrize("lg", [10, 16])
def test_reference_staging_reproduces_survey_appendix_b(ref, orc, lg):
    import os
    if not os.path.exists(os.path.join(po.REF_SRS_DIR, "transcript00.dat")):
        pytest.skip("full SRS not staged")
    n = 1 << lg
    kat = APPENDIX_B[lg]
    s = ref.debug_random_frs(n)
    red = np.array(ref.reduce(po.FR, s))
    assert "%016x" % po.fnv1a64(red) == kat["scalars"]
    pts = ref.read_transcript_g1(n)
    table = ref.point_table(pts)
    assert ref.jac_to_buffer(ref.pippenger(s, table, n=n, unsafe=True)).hex() == kat["msm"]
    for kind, key in ((po.NTT_FFT, "fft"), (po.NTT_COSET_FFT, "coset_fft"), (po.NTT_IFFT, "ifft")):
        assert "%016x" % po.fnv1a64(ref.reduce(po.FR, ref.ntt(kind, s))) == kat[key]
    # and the plain-C oracle agrees with the reference on the same inputs
    assert orc.jac_to_buffer(orc.pippenger(s, table, n=n)).hex() == kat["msm"]
    assert "%016x" % po.fnv1a64(orc.reduce(po.FR, orc.ntt(po.NTT_FFT, s))) == kat["fft"]
    assert "%016x" % po.fnv1a64(orc.reduce(po.FR, orc.ntt(po.NTT_COSET_FFT, s))) == kat["coset_fft"]
    assert "%016x" % po.fnv1a64(orc.reduce(po.FR, orc.ntt(po.NTT_IFFT, s))) == kat["ifft"]
