"""GPU parity tests: the CUDA path, called through the C-ABI (libbbg.so via the ctypes binding),
against the oracle on identical seeded inputs and against the committed golden vectors generated
from the unmodified reference.  Bit-exact on canonical encodings:
  field / NTT outputs : reduce_once'd limbs identical
  MSM / g1 outputs    : affine_element::to_buffer() 64 bytes identical
Every test here needs a B200: `python -m pytest tests -m gpu`.
"""
import os

import numpy as np
import pytest

import inputs
from helpers import msm_scalars, unhex
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bbg():
    import bbg as _bbg
    _bbg.init(0)
    return _bbg


def canon(orc, a):
    return np.array(orc.reduce(po.FR, np.asarray(a).reshape(-1, 4)))


# ------------------------------------------------------------------------------------------ L0
@pytest.mark.parametrize("name,fid", [("fq", po.FQ), ("fr", po.FR)])
def test_field_golden(bbg, golden, orc, name, fid):
    v = golden["fields"][name]
    a, b = unhex(v["a_mont"]), unhex(v["b_mont"])
    # raw limbs of mul / sqr match the reference bit for bit (same integer (ab + mp) / 2^256)
    assert np.array_equal(bbg.field_op(fid, 0, a, b), unhex(v["mul_raw"]))
    assert np.array_equal(bbg.field_op(fid, 3, a), unhex(v["sqr_raw"]))
    assert np.array_equal(bbg.field_op(fid, 7, bbg.field_op(fid, 1, a, b)), unhex(v["add"]))
    assert np.array_equal(bbg.field_op(fid, 7, bbg.field_op(fid, 2, a, b)), unhex(v["sub"]))
    # -0: the reference returns 2p (reduce_once -> p), the device returns 0; same field element, so
    # canonicalise fully (two conditional subtractions) before comparing
    neg = bbg.field_op(fid, 7, bbg.field_op(fid, 7, bbg.field_op(fid, 8, a)))
    assert np.array_equal(neg, bbg.field_op(fid, 7, unhex(v["neg"])))
    assert np.array_equal(bbg.field_op(fid, 5, a), unhex(v["from_mont"]))
    ints = po.ints_to_array([int(s, 16) for s in v["a_int"]])
    assert np.array_equal(bbg.field_op(fid, 4, ints), a)


@pytest.mark.parametrize("fid", [po.FQ, po.FR])
def test_field_random_vs_oracle(bbg, orc, fid):
    n = 4096
    a = inputs.fr_elements(11 + fid, n, coarse_fraction=0.3)  # top limb < modulus top limb for both fields
    b = inputs.fr_elements(13 + fid, n, coarse_fraction=0.3)
    for op in (0, 1, 2, 3, 5, 7, 8):
        got = bbg.field_op(fid, op, a, b if op in (0, 1, 2) else None)
        exp = orc.field_op(fid, op, a, b if op in (0, 1, 2) else None)
        if op in (0, 3, 5, 7):
            assert np.array_equal(got, exp), op  # raw limbs
        else:
            assert np.array_equal(orc.reduce(fid, orc.reduce(fid, got)), orc.reduce(fid, orc.reduce(fid, exp))), op


# bb/ecc/curves/bn254/fq.test.cpp:71-166 restated on the device
def test_fq_kats_on_device(bbg, orc):
    def F(*l):
        return np.array(l, dtype=np.uint64).reshape(1, 4)
    a = F(0x2523b6fa3956f038, 0x158aa08ecdd9ec1d, 0xf48216a4c74738d4, 0x2514cc93d6f0a1bf)
    b = F(0xb68aee5e4c8fc17c, 0xc5193de7f401d5e8, 0xb8777d4dde671db3, 0xe513e75c087b0bb)
    e = F(0x7ed4174114b521c4, 0x58f5bd1d4279fdc2, 0x6a73ac09ee843d41, 0x687a76ae9b3425c)
    assert np.array_equal(bbg.field_op(po.FQ, 7, bbg.field_op(po.FQ, 0, a, b)), e)
    a = F(0x329596aa978981e8, 0x8542e6e254c2a5d0, 0xc5b687d82eadb178, 0x2d242aaf48f56b8a)
    e = F(0xbf4fb34e120b8b12, 0xf64d70efbf848328, 0xefbb6a533f2e7d89, 0x1de50f941425e4aa)
    assert np.array_equal(bbg.field_op(po.FQ, 7, bbg.field_op(po.FQ, 3, a)), e)


def test_g1_golden(bbg, golden, orc, srs_mini):
    pts, table = srs_mini
    g = golden["g1"]
    s = unhex(g["scalars"])
    jacs = np.stack([orc.g1_mul(pts[i + 1], s[i]) for i in range(8)])
    for i in range(8):
        assert orc.jac_to_buffer(jacs[i]).hex() == g["mul"][i]
    madd = bbg.g1_op(0, jacs, np.stack([pts[100 + i] for i in range(8)]))
    add = bbg.g1_op(1, jacs, np.stack([jacs[(i + 3) % 8] for i in range(8)]))
    dbl = bbg.g1_op(2, jacs)
    for i in range(8):
        assert orc.jac_to_buffer(madd[i]).hex() == g["mixed_add"][i]
        assert orc.jac_to_buffer(add[i]).hex() == g["add"][i]
        assert orc.jac_to_buffer(dbl[i]).hex() == g["dbl"][i]
    # edge cases (g1.test.cpp:122-254): inf + P, P + P through the mixed adder, P + (-P)
    inf = orc.g1_infinity().reshape(1, 12)
    p5 = bbg.g1_op(0, inf, pts[5].reshape(1, 8))
    assert orc.jac_to_buffer(p5[0]).hex() == g["edge"]["inf_plus_affine"]
    assert orc.jac_to_buffer(bbg.g1_op(0, p5, pts[5].reshape(1, 8))[0]).hex() == g["edge"]["p_plus_p_mixed"]
    negp5 = pts[5].copy()
    negp5[4:] = orc.field_op(po.FQ, po.OP_NEG, pts[5][4:].reshape(1, 4))[0]
    assert orc.jac_to_buffer(bbg.g1_op(0, p5, negp5.reshape(1, 8))[0]).hex() == g["edge"]["p_minus_p_mixed"]
    assert orc.jac_to_buffer(bbg.g1_op(1, inf, inf)[0]).hex() == g["edge"]["inf_buffer"]
    assert orc.jac_to_buffer(bbg.g1_sum(jacs)).hex() == g["edge"]["sum8"]
    # P + P and P + (-P) through the full adder
    assert orc.jac_to_buffer(bbg.g1_op(1, jacs, jacs)[3]) == orc.jac_to_buffer(orc.g1_dbl(jacs[3]))


# ------------------------------------------------------------------------------------------ SRS
def test_srs_decode_and_point_table(bbg, golden, orc, srs_mini):
    pts, table = srs_mini
    n = inputs.SRS_MINI_POINTS
    got = bbg.read_transcript_g1(n, inputs.SRS_MINI_DIR)
    assert "%016x" % po.fnv1a64(got) == golden["srs_mini"]["fnv_points"]
    assert np.array_equal(got, pts)
    t = bbg.generate_pippenger_point_table(pts)
    assert "%016x" % po.fnv1a64(t) == golden["srs_mini"]["fnv_table"]
    pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)
    assert pip.get_num_points() == n
    assert np.array_equal(pip.get_point_table(), table[: 2 * n])
    with open(os.path.join(inputs.SRS_MINI_DIR, "transcript00.dat"), "rb") as f:
        raw = f.read()[28:28 + (n - 1) * 64]
    pip2 = bbg.Pippenger.from_raw(raw, n)
    assert np.array_equal(pip2.get_point_table(), table[: 2 * n])
    # srs/io.cpp:159-161: too few points is an error, not a short read
    with pytest.raises(bbg.BbgError) as ei:
        bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n + 1)
    assert "Is your srs large enough" in str(ei.value)


# ------------------------------------------------------------------------------------------ MSM
def test_msm_golden(bbg, golden, orc, srs_mini):
    pts, table = srs_mini
    pip = bbg.Pippenger.from_table(table, inputs.SRS_MINI_POINTS)
    for case in golden["msm"]:
        sc = msm_scalars(case, orc)
        n = case["n"]
        if case["kind"] == "repeated_point7":
            rep = np.zeros((n, 8), dtype=np.uint64)
            rep[:] = pts[7]
            res = bbg.msm_points(sc, rep)
        else:
            res = pip.pippenger_unsafe(sc, 0, n)
            # the free functions (points = 2n interleaved host table) give the same element
            res2 = bbg.pippenger_unsafe(sc, table, n)
            assert orc.jac_to_buffer(res2).hex() == case["result"], case
            res3 = bbg.pippenger(sc, po.aligned_copy(table[: 2 * max(n, 1)]), n)
            assert orc.jac_to_buffer(res3).hex() == case["result"], case
        assert orc.jac_to_buffer(res).hex() == case["result"], case


@pytest.mark.parametrize("n", [1, 3, 31, 33, 257, 1000, 4096])
def test_msm_vs_oracle_ragged(bbg, orc, srs_mini, n):
    pts, table = srs_mini
    sc = inputs.fr_elements(900 + n, n, coarse_fraction=0.2)
    exp = orc.jac_to_buffer(orc.pippenger(sc, table, n=n, stride=2))
    assert orc.jac_to_buffer(bbg.msm_points(sc, pts[:n])) == exp


def test_msm_from_range(bbg, orc, srs_mini):
    """Pippenger::pippenger_unsafe(scalars, from, range) (pippenger.cpp:27-31) + g1_sum (c_bind.cpp:40-45):
    the reference's own shard-by-range-then-sum hook."""
    pts, table = srs_mini
    pip = bbg.Pippenger.from_points(pts)
    n = 4096
    sc = inputs.fr_elements(950, n)
    whole = orc.jac_to_buffer(pip.pippenger_unsafe(sc, 0, n))
    parts = []
    for lo, hi in ((0, 1000), (1000, 1001), (1001, 3000), (3000, 4096)):
        parts.append(pip.pippenger_unsafe(sc[lo:hi], lo, hi - lo))
    assert orc.jac_to_buffer(bbg.g1_sum(np.stack(parts))) == whole
    assert whole == orc.jac_to_buffer(orc.pippenger(sc, table, n=n, stride=2))
    with pytest.raises(bbg.BbgError):
        pip.pippenger_unsafe(sc[:10], 4090, 10)


def test_msm_edge_cases(bbg, orc, srs_mini):
    pts, table = srs_mini
    inf_buf = orc.jac_to_buffer(orc.g1_infinity())
    # scalar_multiplication.test.cpp:895-908 zero points, :910-927 mul by zero
    assert orc.jac_to_buffer(bbg.msm_points(np.zeros((0, 4), np.uint64), np.zeros((0, 8), np.uint64), 0)) == inf_buf
    assert orc.jac_to_buffer(bbg.msm_points(np.zeros((100, 4), np.uint64), pts[:100])) == inf_buf
    # :862-893 pippenger_one
    sc = inputs.fr_elements(3, 1)
    assert orc.jac_to_buffer(bbg.msm_points(sc, pts[1:2])) == orc.jac_to_buffer(orc.g1_mul(pts[1], sc[0]))
    # points at infinity in the input are skipped (safe path semantics)
    p = pts[:64].copy()
    p[5, 3] |= np.uint64(1 << 63)
    sc = inputs.fr_elements(4, 64)
    exp = orc.pippenger(np.delete(sc, 5, axis=0), np.delete(pts[:64], 5, axis=0), stride=1)
    assert orc.jac_to_buffer(bbg.msm_points(sc, p)) == orc.jac_to_buffer(exp)
    # P and -P with the same scalar cancel; all-equal points with scalars summing to zero give infinity
    q = np.stack([pts[9], pts[9]])
    q[1, 4:] = orc.field_op(po.FQ, po.OP_NEG, pts[9][4:].reshape(1, 4))[0]
    s2 = np.stack([sc[0], sc[0]])
    assert orc.jac_to_buffer(bbg.msm_points(s2, q)) == inf_buf
    # scalar r - 1 (== -1) and the largest coarse representative 2r - 1
    m1 = orc.to_mont(po.FR, [po.FR_MODULUS - 1])
    assert orc.jac_to_buffer(bbg.msm_points(m1, pts[3:4])) == orc.jac_to_buffer(orc.g1_mul(pts[3], m1[0]))


def test_msm_skewed_buckets(bbg, orc, srs_mini):
    """All scalars equal: every point of a window lands in ONE bucket (the load-balance worst case)."""
    pts, table = srs_mini
    n = 4096
    one = inputs.fr_elements(12, 1)
    sc = np.repeat(one, n, axis=0)
    assert orc.jac_to_buffer(bbg.msm_points(sc, pts[:n])) == orc.jac_to_buffer(orc.pippenger(sc, table, n=n, stride=2))


@pytest.mark.parametrize("passes,c,k,mat", [(1, 0, 0, 1), (2, 6, 1, 0), (3, 0, 7, 1), (6, 5, 32, 0), (12, 4, 64, 1)])
def test_msm_pairwise_affine_passes(bbg, orc, srs_mini, passes, c, k, mat, monkeypatch):
    """Batched-affine bucket accumulation (msm.cu k_msm_pair_pass: pairwise affine additions sharing one Kaliski
    inversion per CTA batch, the GPU counterpart of scalar_multiplication.cpp:273-401, :523-718).  Large MSMs take
    this path by themselves; here it is forced on the 4096-point SRS for every number of levels, including more
    levels than any bucket needs, on uniform, skewed, repeated-point (doubling), cancelling and infinite inputs."""
    pts, table = srs_mini
    monkeypatch.setenv("BBG_MSM_PAIR_PASSES", str(passes))
    monkeypatch.setenv("BBG_MSM_MATERIALISE", str(mat))  # 1: the counting sort moves the points; 0: schedule words + gather
    if k:
        monkeypatch.setenv("BBG_MSM_PAIR_K", str(k))  # output slots per thread (default: chosen per level)
    if c:
        monkeypatch.setenv("BBG_MSM_C", str(c))
        monkeypatch.setenv("BBG_MSM_C1", str(c))
    pip = bbg.Pippenger.from_points(pts)
    inf_buf = orc.jac_to_buffer(orc.g1_infinity())
    for n in (4096, 3000, 33, 2, 1):
        sc = inputs.fr_elements(1200 + n + passes, n, coarse_fraction=0.2)
        exp = orc.jac_to_buffer(orc.pippenger(sc, pts[:n], stride=1))
        assert orc.jac_to_buffer(pip.pippenger_unsafe(sc, 0, n)) == exp, n
        assert orc.jac_to_buffer(bbg.msm_points(sc, pts[:n])) == exp, n
    # one bucket per window holds everything
    n = 4096
    one = np.repeat(inputs.fr_elements(5, 1), n, axis=0)
    assert orc.jac_to_buffer(pip.pippenger_unsafe(one, 0, n)) == orc.jac_to_buffer(orc.pippenger(one, pts[:n], stride=1))
    # the same point 64 times: every pairwise addition is a doubling
    rep = np.zeros((64, 8), dtype=np.uint64)
    rep[:] = pts[7]
    sc = inputs.fr_elements(77, 64)
    same = np.repeat(sc[:1], 64, axis=0)
    tot = orc.field_op(po.FR, po.OP_MUL, orc.to_mont(po.FR, [64]), same[:1])  # 64 * s
    assert orc.jac_to_buffer(bbg.msm_points(same, rep)) == orc.jac_to_buffer(orc.g1_mul(pts[7], tot[0]))
    assert orc.jac_to_buffer(bbg.msm_points(sc, rep)) == orc.jac_to_buffer(orc.pippenger(sc, rep, stride=1))
    # P and -P with equal scalars meet in the same buckets and cancel; an infinite input is skipped
    q = np.stack([pts[9], pts[9], pts[11], pts[12]])
    q[1, 4:] = orc.field_op(po.FQ, po.OP_NEG, pts[9][4:].reshape(1, 4))[0]
    s4 = np.stack([sc[0], sc[0], sc[1], sc[2]])
    exp = orc.pippenger(s4[2:], q[2:], stride=1)
    assert orc.jac_to_buffer(bbg.msm_points(s4, q)) == orc.jac_to_buffer(exp)
    assert orc.jac_to_buffer(bbg.msm_points(s4[:2], q[:2])) == inf_buf
    p = pts[:64].copy()
    p[5, 3] |= np.uint64(1 << 63)
    s64 = inputs.fr_elements(4, 64)
    exp = orc.pippenger(np.delete(s64, 5, axis=0), np.delete(pts[:64], 5, axis=0), stride=1)
    assert orc.jac_to_buffer(bbg.msm_points(s64, p)) == orc.jac_to_buffer(exp)
    assert orc.jac_to_buffer(bbg.msm_points(np.zeros((100, 4), np.uint64), pts[:100])) == inf_buf


# ------------------------------------------------------------------------------------------ NTT
def test_ntt_golden(bbg, golden, orc):
    for e in golden["ntt"]:
        n = 1 << e["log2n"]
        x = inputs.fr_elements(e["seed"], n, coarse_fraction=0.25)
        const = inputs.fr_elements(e["const_seed"], 1)[0]
        y = canon(orc, bbg.ntt(x.copy(), e["kind"], generator_size=e["generator_size"], constant=const))
        assert "%016x" % po.fnv1a64(y) == e["fnv"], e
        if "full" in e:
            assert np.array_equal(y, unhex(e["full"]))


def test_coset_fft_ext_golden(bbg, golden, orc):
    for e in golden["coset_fft_ext"]:
        n = 1 << e["log2n"]
        buf = np.zeros((n * e["ext"], 4), dtype=np.uint64)
        buf[:n] = inputs.fr_elements(e["seed"], n)
        y = canon(orc, bbg.coset_fft_ext(buf, n, e["ext"]))
        assert "%016x" % po.fnv1a64(y) == e["fnv"], e


@pytest.mark.parametrize("lg", list(range(0, 19)))
def test_ntt_vs_oracle_all_sizes(bbg, orc, lg):
    n = 1 << lg
    x = inputs.fr_elements(700 + lg, n, coarse_fraction=0.25)
    const = inputs.fr_elements(800 + lg, 1)[0]
    kinds = range(8) if lg <= 14 else (bbg.FFT, bbg.COSET_IFFT)
    for kind in kinds:
        gs = n // 4 if (kind in (2, 6, 7) and n >= 4) else 0
        got = canon(orc, bbg.ntt(x.copy(), kind, generator_size=gs, constant=const))
        exp = canon(orc, orc.ntt(kind, x, generator_size=gs, constant=const))
        assert np.array_equal(got, exp), (lg, kind)


@pytest.mark.parametrize("lg", [6, 9, 12, 13, 17])
def test_ntt_repeated_calls_use_cached_scale_tables(bbg, orc, lg):
    """The library promotes a (constant, shift, size) scaling to a cached full table the SECOND time it sees it, and a
    plain ifft then folds 1/n into its last inter-pass twiddles: results must not depend on which path ran.  Also cycles
    through more distinct constants than the cache holds (eviction) and back."""
    n = 1 << lg
    x = inputs.fr_elements(900 + lg, n, coarse_fraction=0.25)
    const = inputs.fr_elements(901 + lg, 1)[0]
    exp = {}
    for rep in range(3):
        for kind in range(8):
            for gs in ((0, n // 4) if kind in (2, 6, 7) else (0,)):
                if (kind, gs) not in exp:
                    exp[(kind, gs)] = canon(orc, orc.ntt(kind, x, generator_size=gs, constant=const))
                got = canon(orc, bbg.ntt(x.copy(), kind, generator_size=gs, constant=const))
                assert np.array_equal(got, exp[(kind, gs)]), (lg, kind, gs, rep)
    if lg == 9:
        consts = inputs.fr_elements(77, 12)
        want = [canon(orc, orc.ntt(po.NTT_COSET_FFT_GEN_SHIFT, x, constant=c)) for c in consts]
        for rep in range(3):
            for c, w in zip(consts, want):
                assert np.array_equal(canon(orc, bbg.coset_fft_with_generator_shift(x.copy(), c)), w), rep
        # and the first keys again, after they were evicted
        got = canon(orc, bbg.ntt(x.copy(), bbg.COSET_IFFT))
        assert np.array_equal(got, exp[(bbg.COSET_IFFT, 0)])


def test_ntt_device_pointer_entry(bbg, orc):
    import torch
    n = 1 << 12
    x = inputs.fr_elements(42, n)
    t = torch.from_numpy(x.view(np.int64)).cuda()
    bbg.coset_fft(t, generator_size=n // 4)
    torch.cuda.synchronize()
    got = t.cpu().numpy().view(np.uint64)
    assert np.array_equal(canon(orc, got), canon(orc, orc.ntt(po.NTT_COSET_FFT, x, generator_size=n // 4)))


def test_msm_device_pointer_entry(bbg, orc, srs_mini):
    import torch
    pts, table = srs_mini
    n = 2048
    pip = bbg.Pippenger.from_points(pts)
    sc = inputs.fr_elements(43, n)
    t = torch.from_numpy(sc.view(np.int64)).cuda()
    out = pip.pippenger_unsafe(t, 100, n)
    torch.cuda.synchronize()
    jac = out.cpu().numpy().view(np.uint64)
    assert orc.jac_to_buffer(jac) == orc.jac_to_buffer(orc.pippenger(sc, pts[100:100 + n], stride=1))


def test_ntt_rejects_bad_sizes(bbg):
    with pytest.raises(bbg.BbgError):
        bbg.fft(np.zeros((3, 4), dtype=np.uint64))
    with pytest.raises(bbg.BbgError):
        bbg.ntt(np.zeros((4, 4), dtype=np.uint64), bbg.FFT_WITH_CONSTANT)  # constant missing


def test_launch_counter_moves(bbg):
    before = bbg.kernel_launches()
    bbg.fft(inputs.fr_elements(1, 256))
    assert bbg.kernel_launches() > before


@pytest.mark.parametrize("levels,c", [(1, 0), (2, 0), (5, 7), (0, 0), (0, 11), (0, 12), (4, 13), (0, 16)])
def test_msm_fixed_base_levels(bbg, orc, srs_mini, levels, c, monkeypatch):
    """The Pippenger object's precomputed levels 2^(D l) * P_i (msm.cu k_msm_precompute): every split of the
    windows into levels x bucket sets must give the same group element, for sub-ranges too."""
    pts, table = srs_mini
    if levels:
        monkeypatch.setenv("BBG_MSM_LEVELS", str(levels))
    if c:
        monkeypatch.setenv("BBG_MSM_C", str(c))
    pip = bbg.Pippenger.from_points(pts)
    n = 3000
    sc = inputs.fr_elements(970 + levels + c, n, coarse_fraction=0.2)
    assert orc.jac_to_buffer(pip.pippenger_unsafe(sc, 0, n)) == orc.jac_to_buffer(orc.pippenger(sc, pts[:n], stride=1))
    assert orc.jac_to_buffer(pip.pippenger_unsafe(sc[:500], 77, 500)) == orc.jac_to_buffer(orc.pippenger(sc[:500], pts[77:577], stride=1))
    one = np.repeat(inputs.fr_elements(5, 1), n, axis=0)  # one bucket holds everything: exercises every merge level
    assert orc.jac_to_buffer(pip.pippenger_unsafe(one, 0, n)) == orc.jac_to_buffer(orc.pippenger(one, pts[:n], stride=1))


@pytest.mark.parametrize("wide_from", [1, 100, 10 ** 9])
def test_msm_slot_merge_wide_and_team_workers(bbg, orc, srs_mini, wide_from, monkeypatch):
    """k_msm_merge<TEAM>: a merge level with many workers runs one thread per worker (plain additions), otherwise a team of
    four lanes per worker; BBG_MSM_MERGE_WIDE_FROM moves the switch so that both forms see every level on uniform, skewed
    (all digits in one bucket: every level has work) and mixed inputs."""
    pts, table = srs_mini
    monkeypatch.setenv("BBG_MSM_MERGE_WIDE_FROM", str(wide_from))
    pip = bbg.Pippenger.from_points(pts)
    n = inputs.SRS_MINI_POINTS
    sc = inputs.fr_elements(2300, n, coarse_fraction=0.3)
    assert orc.jac_to_buffer(pip.pippenger_unsafe(sc, 0, n)) == orc.jac_to_buffer(orc.pippenger(sc, pts[:n], stride=1))
    one = np.repeat(inputs.fr_elements(8, 1), n, axis=0)
    assert orc.jac_to_buffer(pip.pippenger_unsafe(one, 0, n)) == orc.jac_to_buffer(orc.pippenger(one, pts[:n], stride=1))
    two = one.copy()
    two[::2] = inputs.fr_elements(9, 1)
    assert orc.jac_to_buffer(pip.pippenger_unsafe(two, 0, n)) == orc.jac_to_buffer(orc.pippenger(two, pts[:n], stride=1))
    assert orc.jac_to_buffer(bbg.msm_points(sc[:777], pts[:777])) == orc.jac_to_buffer(orc.pippenger(sc[:777], pts[:777], stride=1))


@pytest.mark.parametrize("c", [12, 16, 18, 20])
def test_msm_short_range_on_large_object_subdivides_windows(bbg, orc, srs_mini, c, monkeypatch):
    """Pippenger::pippenger_unsafe(scalars, from, range) with a range far below the size the object's window was chosen
    for: msm_device narrows the window to D / k (k bucket sets per fixed-base level) at call time.  Ranges on both sides of
    the switch, against the oracle; BBG_MSM_CALL_WINDOW=0 (the object's own window) must give the same points."""
    pts, table = srs_mini
    monkeypatch.setenv("BBG_MSM_C", str(c))
    pip = bbg.Pippenger.from_points(pts)
    for frm, rng in ((0, 4096), (11, 700), (100, 64), (5, 9), (0, 1)):
        sc = inputs.fr_elements(2100 + c + rng, rng, coarse_fraction=0.2)
        exp = orc.jac_to_buffer(orc.pippenger(sc, pts[frm:frm + rng], stride=1))
        assert orc.jac_to_buffer(pip.pippenger_unsafe(sc, frm, rng)) == exp, (c, frm, rng)
        monkeypatch.setenv("BBG_MSM_CALL_WINDOW", "0")
        assert orc.jac_to_buffer(pip.pippenger_unsafe(sc, frm, rng)) == exp, (c, frm, rng)
        monkeypatch.delenv("BBG_MSM_CALL_WINDOW")
    batch = [inputs.fr_elements(2200 + i, 300) for i in range(3)]
    got = pip.pippenger_unsafe_batch(batch, 40, 300)
    for i in range(3):
        assert orc.jac_to_buffer(got[i]) == orc.jac_to_buffer(orc.pippenger(batch[i], pts[40:340], stride=1))


@pytest.mark.parametrize("parts,c", [(2, 14), (4, 15), (2, 16), (4, 14), (1, 14)])
def test_msm_bucket_range_parts(bbg, orc, srs_mini, parts, c, monkeypatch):
    """msm.cu "parts": one bucket set cut into contiguous bucket ranges, each with its own accumulate -> merge -> reduce
    chain on its own stream (range h adds h B / H times the plain sum of its buckets), summed at the end.  Same group
    element as the oracle for uniform, skewed (every digit in one bucket) and sub-range inputs; (4, 14) asks for more
    parts than the bucket count allows and must fall back to fewer."""
    pts, table = srs_mini
    monkeypatch.setenv("BBG_MSM_PARTS", str(parts))
    monkeypatch.setenv("BBG_MSM_C", str(c))
    pip = bbg.Pippenger.from_points(pts)
    n = 4000
    sc = inputs.fr_elements(1970 + parts + c, n, coarse_fraction=0.2)
    assert orc.jac_to_buffer(pip.pippenger_unsafe(sc, 0, n)) == orc.jac_to_buffer(orc.pippenger(sc, pts[:n], stride=1))
    assert orc.jac_to_buffer(pip.pippenger_unsafe(sc[:333], 55, 333)) == orc.jac_to_buffer(orc.pippenger(sc[:333], pts[55:388], stride=1))
    one = np.repeat(inputs.fr_elements(6, 1), n, axis=0)
    assert orc.jac_to_buffer(pip.pippenger_unsafe(one, 0, n)) == orc.jac_to_buffer(orc.pippenger(one, pts[:n], stride=1))
    # digits only in the top bucket range / only in the bottom one: small scalars (low windows, low buckets) and r - small
    small_m = orc.to_mont(po.FR, list(range(1, 501)))
    assert orc.jac_to_buffer(pip.pippenger_unsafe(small_m, 0, 500)) == orc.jac_to_buffer(orc.pippenger(small_m, pts[:500], stride=1))
    neg_m = orc.to_mont(po.FR, [po.FR_MODULUS - k for k in range(1, 501)]) if hasattr(po, "FR_MODULUS") else orc.field_op(po.FR, po.OP_NEG, small_m)
    assert orc.jac_to_buffer(pip.pippenger_unsafe(neg_m, 0, 500)) == orc.jac_to_buffer(orc.pippenger(neg_m, pts[:500], stride=1))
    back = pip.pippenger_unsafe(sc, 0, n)  # the same call again: workspaces and counters are reusable
    assert orc.jac_to_buffer(back) == orc.jac_to_buffer(orc.pippenger(sc, pts[:n], stride=1))


@pytest.mark.parametrize("lg,world", [(12, 2), (14, 4), (16, 8), (18, 8), (20, 2)])
def test_ntt_multi_gpu_data_path_simulated(bbg, orc, lg, world):
    """The multi-GPU four-step NTT (phase 0 / all-to-all / phase 1, csrc/ntt.cu + bbg/dist_ntt.py) with every rank run in
    turn on one device: same kernels, same index maths as the torchrun path; must equal the single-array transform."""
    import torch
    from bbg import dist_ntt
    n = 1 << lg
    x = inputs.fr_elements(3000 + lg, n, coarse_fraction=0.25)
    const = inputs.fr_elements(3100 + lg, 1)[0]
    xt = torch.from_numpy(x.view(np.int64)).cuda()
    for kind, gs in ((bbg.FFT, 0), (bbg.COSET_FFT, n // 4), (bbg.COSET_IFFT, 0), (bbg.IFFT_WITH_CONSTANT, 0)):
        got = dist_ntt.simulate(bbg, xt, kind, world, generator_size=gs, constant=const).cpu().numpy().view(np.uint64)
        exp = bbg.ntt(x.copy(), kind, generator_size=gs, constant=const)
        assert np.array_equal(canon(orc, got), canon(orc, exp)), (lg, world, kind)
    if lg <= 14:
        assert np.array_equal(canon(orc, exp), canon(orc, orc.ntt(po.NTT_IFFT_CONST, x, constant=const)))


@pytest.mark.parametrize("lg,world", [(12, 2), (14, 4), (16, 8), (18, 8), (20, 2)])
def test_ntt_fused_exchange_simulated(bbg, orc, lg, world):
    """bbg_ntt_dist_fused_dev: the pass before the exchange stores every element straight into the owner's receive
    buffer (NVLink peer memory in the torchrun path, buffers of the same device here) at the slot the all-to-all would
    have used; phase 1 on those buffers must reproduce the single-array transform bit for bit."""
    import torch
    from bbg import dist_ntt
    n = 1 << lg
    x = inputs.fr_elements(3300 + lg, n, coarse_fraction=0.25)
    const = inputs.fr_elements(3400 + lg, 1)[0]
    xt = torch.from_numpy(x.view(np.int64)).cuda()
    for kind, gs in ((bbg.FFT, 0), (bbg.COSET_FFT, n // 4), (bbg.COSET_IFFT, 0), (bbg.IFFT_WITH_CONSTANT, 0)):
        got = dist_ntt.simulate(bbg, xt, kind, world, generator_size=gs, constant=const, fused=True).cpu().numpy().view(np.uint64)
        exp = bbg.ntt(x.copy(), kind, generator_size=gs, constant=const)
        assert np.array_equal(canon(orc, got), canon(orc, exp)), (lg, world, kind)


@pytest.mark.parametrize("lg,world", [(12, 2), (14, 4), (16, 8), (17, 4), (20, 8)])
def test_ntt_natural_blocks_over_peer_memory_simulated(bbg, orc, lg, world):
    """bbg_ntt_dist_natural_dev: rank q holds the natural block x[q n / W, (q + 1) n / W) and ends with the natural block of
    X; the first pass loads from the owners' blocks, the last pass stores to the owners' blocks (peer memory under torchrun,
    buffers of one device here).  The concatenated output blocks must be the single-array transform."""
    import torch
    from bbg import dist_ntt
    n = 1 << lg
    x = inputs.fr_elements(3500 + lg, n, coarse_fraction=0.25)
    const = inputs.fr_elements(3600 + lg, 1)[0]
    xt = torch.from_numpy(x.view(np.int64)).cuda()
    for kind, gs in ((bbg.FFT, 0), (bbg.COSET_FFT, n // 4), (bbg.COSET_IFFT, 0), (bbg.IFFT_WITH_CONSTANT, 0)):
        got = dist_ntt.simulate(bbg, xt, kind, world, generator_size=gs, constant=const, fused="natural").cpu().numpy().view(np.uint64)
        exp = bbg.ntt(x.copy(), kind, generator_size=gs, constant=const)
        assert np.array_equal(canon(orc, got), canon(orc, exp)), (lg, world, kind)


def test_multi_process_nccl_paths(bbg):
    """Real N > 1 run (torchrun, NCCL): needs >= 2 visible GPUs, skipped on a single-GPU box."""
    import subprocess
    import sys
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "scripts", "dist_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == world
