"""Pins the plain-C oracle's L0 arithmetic against (a) the reference's own known-answer constants,
(b) golden vectors generated from the unmodified reference, (c) the reference .so when present."""
import numpy as np
import pytest

from helpers import unhex
from oracle import pyoracle as po


def F(*limbs):
    return np.array(limbs, dtype=np.uint64).reshape(1, 4)


# ---- the reference's own KATs, restated (bb/ecc/curves/bn254/fq.test.cpp:71-166)
def test_fq_mul_check_against_constants(orc):
    a = F(0x2523b6fa3956f038, 0x158aa08ecdd9ec1d, 0xf48216a4c74738d4, 0x2514cc93d6f0a1bf)
    b = F(0xb68aee5e4c8fc17c, 0xc5193de7f401d5e8, 0xb8777d4dde671db3, 0xe513e75c087b0bb)
    e = F(0x7ed4174114b521c4, 0x58f5bd1d4279fdc2, 0x6a73ac09ee843d41, 0x687a76ae9b3425c)
    assert np.array_equal(orc.reduce(po.FQ, orc.field_op(po.FQ, po.OP_MUL, a, b)), e)


def test_fq_mul_short_integers(orc):
    e = F(0x65991a6dc2f3a183, 0xe3ba1f83394a2d08, 0x8401df65a169db3f, 0x1727099643607bba)
    assert np.array_equal(orc.reduce(po.FQ, orc.field_op(po.FQ, po.OP_MUL, F(0xa, 0, 0, 0), F(0xb, 0, 0, 0))), e)


def test_fq_sqr_check_against_constants(orc):
    a = F(0x329596aa978981e8, 0x8542e6e254c2a5d0, 0xc5b687d82eadb178, 0x2d242aaf48f56b8a)
    e = F(0xbf4fb34e120b8b12, 0xf64d70efbf848328, 0xefbb6a533f2e7d89, 0x1de50f941425e4aa)
    assert np.array_equal(orc.reduce(po.FQ, orc.field_op(po.FQ, po.OP_SQR, a)), e)


def test_fq_add_check_against_constants(orc):
    a = F(0x7d2e20e82f73d3e8, 0x8e50616a7a9d419d, 0xcdc833531508914b, 0xd510253a2ce62c)
    b = F(0x2829438b071fd14e, 0xb03ef3f9ff9274e, 0x605b671f6dc7b209, 0x8701f9d971fbc9)
    e = F(0xa55764733693a536, 0x995450aa1a9668eb, 0x2e239a7282d04354, 0x15c121f139ee1f6)
    assert np.array_equal(orc.reduce(po.FQ, orc.field_op(po.FQ, po.OP_ADD, a, b)), e)


def test_fq_sub_check_against_constants(orc):
    a = F(0xd68d01812313fb7c, 0x2965d7ae7c6070a5, 0x08ef9af6d6ba9a48, 0x0cb8fe2108914f53)
    b = F(0x2cd2a2a37e9bf14a, 0xebc86ef589c530f6, 0x75124885b362b8fe, 0x1394324205c7a41d)
    e = F(0xe5daeaf47cf50779, 0xd51ed34a5b0d0a3c, 0x4c2d9827a4d939a6, 0x29891a51e3fb4b5f)
    assert np.array_equal(orc.reduce(po.FQ, orc.field_op(po.FQ, po.OP_SUB, a, b)), e)


def _mont(orc, *limbs):
    return orc.field_op(po.FQ, po.OP_TO_MONT, F(*limbs))[0]


def _jac(orc, x, y, z):
    return np.concatenate([_mont(orc, *x), _mont(orc, *y), _mont(orc, *z)])


def _same_point(orc, a, b):
    return orc.jac_to_buffer(a) == orc.jac_to_buffer(b)


# bb/ecc/curves/bn254/g1.test.cpp:39-120
def test_g1_mixed_add_check_against_constants(orc):
    lhs = _jac(orc, (0x92716caa6cac6d26, 0x1e6e234136736544, 0x1bb04588cde00af0, 0x9a2ac922d97e6f5),
               (0x9e693aeb52d79d2d, 0xf0c1895a61e5e975, 0x18cd7f5310ced70f, 0xac67920a22939ad),
               (0xfef593c9ce1df132, 0xe0486f801303c27d, 0x9bbd01ab881dc08e, 0x2a589badf38ec0f9))
    rhs = np.concatenate([_mont(orc, 0xa1ec5d1398660db8, 0x6be3e1f6fd5d8ab1, 0x69173397dd272e11, 0x12575bbfe1198886),
                          _mont(orc, 0xcfbfd4441138823e, 0xb5f817e28a1ef904, 0xefb7c5629dcc1c42, 0x1a9ed3d6f846230e)])
    exp = _jac(orc, (0x2a9d0201fccca20, 0x36f969b294f31776, 0xee5534422a6f646, 0x911dbc6b02310b6),
               (0x14c30aaeb4f135ef, 0x9c27c128ea2017a1, 0xf9b7d80c8315eabf, 0x35e628df8add760),
               (0xa43fe96673d10eb3, 0x88fbe6351753d410, 0x45c21cc9d99cb7d, 0x3018020aa6e9ede5))
    res = orc.g1_mixed_add(lhs, rhs)
    # the reference's formulas give this exact Jacobian representative
    assert np.array_equal(orc.reduce(po.FQ, res.reshape(3, 4)), orc.reduce(po.FQ, exp.reshape(3, 4)))


def test_g1_dbl_check_against_constants(orc):
    lhs = _jac(orc, (0x8d1703aa518d827f, 0xd19cc40779f54f63, 0xabc11ce30d02728c, 0x10938940de3cbeec),
               (0xcf1798994f1258b4, 0x36307a354ad90a25, 0xcd84adb348c63007, 0x6266b85241aff3f),
               (0xe213e18fd2df7044, 0xb2f42355982c5bc8, 0xf65cf5150a3a9da1, 0xc43bde08b03aca2))
    exp = _jac(orc, (0xd5c6473044b2e67c, 0x89b185ea20951f3a, 0x4ac597219cf47467, 0x2d00482f63b12c86),
               (0x4e7e6c06a87e4314, 0x906a877a71735161, 0xaa7b9893cc370d39, 0x62f206bef795a05),
               (0x8813bdca7b0b115a, 0x929104dffdfabd22, 0x3fff575136879112, 0x18a299c1f683bdca))
    res = orc.g1_dbl(orc.g1_dbl(orc.g1_dbl(lhs)))
    assert np.array_equal(orc.reduce(po.FQ, res.reshape(3, 4)), orc.reduce(po.FQ, exp.reshape(3, 4)))


def test_g1_add_check_against_constants(orc):
    lhs = _jac(orc, (0x184b38afc6e2e09a, 0x4965cd1c3687f635, 0x334da8e7539e71c4, 0xf708d16cfe6e14),
               (0x2a6ff6ffc739b3b6, 0x70761d618b513b9, 0xbf1645401de26ba1, 0x114a1616c164b980),
               (0x10143ade26bbd57a, 0x98cf4e1f6c214053, 0x6bfdc534f6b00006, 0x1875e5068ababf2c))
    rhs = _jac(orc, (0xafdb8a15c98bf74c, 0xac54df622a8d991a, 0xc6e5ae1f3dad4ec8, 0x1bd3fb4a59e19b52),
               (0x21b3bb529bec20c0, 0xaabd496406ffb8c1, 0xcd3526c26ac5bdcb, 0x187ada6b8693c184),
               (0xffcd440a228ed652, 0x8a795c8f234145f1, 0xd5279cdbabb05b95, 0xbdf19ba16fc607a))
    exp = _jac(orc, (0x18764da36aa4cd81, 0xd15388d1fea9f3d3, 0xeb7c437de4bbd748, 0x2f09b712adf6f18f),
               (0x50c5f3cab191498c, 0xe50aa3ce802ea3b5, 0xd9d6125b82ebeff8, 0x27e91ba0686e54fe),
               (0xe4b81ef75fedf95, 0xf608edef14913c75, 0xfd9e178143224c96, 0xa8ae44990c8accd))
    res = orc.g1_add(lhs, rhs)
    assert np.array_equal(orc.reduce(po.FQ, res.reshape(3, 4)), orc.reduce(po.FQ, exp.reshape(3, 4)))


# g1.test.cpp:284-299
def test_g1_group_exponentiation_check_against_constants(orc):
    a = orc.field_op(po.FR, po.OP_TO_MONT, F(0xb67299b792199cf0, 0xc1da7df1e7e12768, 0x692e427911532edf, 0x13dd85e87dc89978))[0]
    ex = _mont(orc, 0x9bf840faf1b4ba00, 0xe81b7260d068e663, 0x7610c9a658d2c443, 0x278307cd3d0cddb0)
    ey = _mont(orc, 0xf6ed5fb779ebecb, 0x414ca771acbe183c, 0xe3692cb56dfbdb67, 0x3d3c5ed19b080a3)
    res = orc.g1_to_affine(orc.g1_mul(orc.g1_one(), a))
    assert np.array_equal(orc.reduce(po.FQ, res.reshape(2, 4)), np.stack([ex, ey]))


# ---- golden vectors from the unmodified reference
@pytest.mark.parametrize("name,fid", [("fq", po.FQ), ("fr", po.FR)])
def test_field_golden(orc, golden, name, fid):
    v = golden["fields"][name]
    a, b = unhex(v["a_mont"]), unhex(v["b_mont"])
    assert np.array_equal(orc.to_mont(fid, [int(s, 16) for s in v["a_int"]]), a)
    assert np.array_equal(orc.field_op(fid, po.OP_MUL, a, b), unhex(v["mul_raw"]))  # raw coarse limbs match too
    assert np.array_equal(orc.field_op(fid, po.OP_SQR, a), unhex(v["sqr_raw"]))
    assert np.array_equal(orc.reduce(fid, orc.field_op(fid, po.OP_ADD, a, b)), unhex(v["add"]))
    assert np.array_equal(orc.reduce(fid, orc.field_op(fid, po.OP_SUB, a, b)), unhex(v["sub"]))
    assert np.array_equal(orc.reduce(fid, orc.field_op(fid, po.OP_NEG, a)), unhex(v["neg"]))
    assert np.array_equal(orc.field_op(fid, po.OP_FROM_MONT, a), unhex(v["from_mont"]))
    assert np.array_equal(orc.reduce(fid, orc.field_op(fid, po.OP_INVERT, a[v["invert_idx"]])), unhex(v["invert"]))


def test_constants_golden(orc, golden):
    c = golden["constants"]
    assert np.array_equal(orc.fr_root_of_unity(28), unhex(c["fr_root_of_unity_28"])[0])
    out = po.aligned_empty(4)
    orc.lib.orc_fr_coset_generator(po._p(out))
    assert np.array_equal(out, unhex(c["fr_coset_generator0"])[0])
    out0 = out.copy()
    orc.lib.orc_fq_beta(po._p(out))
    assert np.array_equal(out, unhex(c["fq_beta"])[0])
    one = orc.g1_one()
    assert np.array_equal(one[:4], unhex(c["g1_one_x"])[0]) and np.array_equal(one[4:], unhex(c["g1_one_y"])[0])
    for lg, hexs in golden["domains"].items():
        assert np.array_equal(orc.reduce(po.FR, orc.domain_constants(1 << int(lg))), orc.reduce(po.FR, unhex(hexs))), lg
    five = orc.to_mont(po.FR, [5])
    assert np.array_equal(orc.reduce(po.FR, out0.reshape(1, 4)), five)  # coset generator 0 is 5


def test_srs_mini_loader_golden(orc, golden, srs_mini):
    pts, table = srs_mini
    g = golden["srs_mini"]
    assert "%016x" % po.fnv1a64(pts) == g["fnv_points"]
    assert "%016x" % po.fnv1a64(table[: 2 * g["num_points"]]) == g["fnv_table"]
    assert orc.affine_to_buffer(pts[1]).hex() == g["point1_buffer"]
    assert orc.affine_to_buffer(pts[4095]).hex() == g["point4095_buffer"]
    assert all(orc.g1_on_curve(pts[i]) for i in (0, 1, 2, 4095))
    with pytest.raises(RuntimeError):  # srs too short (io.cpp:159-161)
        orc.read_transcript_g1(5000, __import__("inputs").SRS_MINI_DIR)


def test_g1_golden(orc, golden, srs_mini):
    pts, table = srs_mini
    g = golden["g1"]
    s = unhex(g["scalars"])
    jacs = [orc.g1_mul(pts[i + 1], s[i]) for i in range(8)]
    for i in range(8):
        assert orc.jac_to_buffer(jacs[i]).hex() == g["mul"][i]
        assert orc.jac_to_buffer(orc.g1_mixed_add(jacs[i], pts[100 + i])).hex() == g["mixed_add"][i]
        assert orc.jac_to_buffer(orc.g1_add(jacs[i], jacs[(i + 3) % 8])).hex() == g["add"][i]
        assert orc.jac_to_buffer(orc.g1_dbl(jacs[i])).hex() == g["dbl"][i]
    inf = orc.g1_infinity()
    e = g["edge"]
    p5 = orc.g1_mixed_add(inf, pts[5])
    assert orc.jac_to_buffer(p5).hex() == e["inf_plus_affine"]
    assert orc.jac_to_buffer(orc.g1_mixed_add(p5, pts[5])).hex() == e["p_plus_p_mixed"]
    neg = pts[5].copy()
    neg[4:] = orc.field_op(po.FQ, po.OP_NEG, pts[5][4:].reshape(1, 4))[0]
    assert orc.jac_to_buffer(orc.g1_mixed_add(p5, neg)).hex() == e["p_minus_p_mixed"] == e["inf_buffer"]
    assert orc.jac_to_buffer(inf).hex() == e["inf_buffer"]
    assert orc.jac_to_buffer(orc.g1_sum(np.stack(jacs))).hex() == e["sum8"]


# ---- live cross-check against the compiled reference (dev container / any box the .so travelled to)
@pytest.mark.ref
def test_oracle_matches_reference_live(orc, ref):
    import inputs
    for fid in (po.FQ, po.FR):
        a = inputs.fr_elements(900 + fid, 200, coarse_fraction=0.3)
        b = inputs.fr_elements(910 + fid, 200, coarse_fraction=0.3)
        for op in (po.OP_MUL, po.OP_SQR):
            assert np.array_equal(orc.field_op(fid, op, a, b), ref.field_op(fid, op, a, b))
        for op in (po.OP_ADD, po.OP_SUB, po.OP_NEG, po.OP_FROM_MONT, po.OP_TO_MONT):
            assert np.array_equal(orc.reduce(fid, orc.field_op(fid, op, a, b)), ref.reduce(fid, ref.field_op(fid, op, a, b)))
