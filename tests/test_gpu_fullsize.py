"""Parity at BASELINE.json's full sizes (MSM 2^20 over the shipped SRS, NTT family at 2^22).

Two kinds of evidence, both through the C-ABI:
  * against the UNMODIFIED reference compiled into oracle/_ref/libbbref.so (it travels to the GPU box): the
    reference's CPU pippenger_unsafe / fft / ifft / coset_fft on the same inputs, canonical encodings identical;
  * size-independent properties that need no checker at that size: MSM linearity
    (msm(a) + msm(b) == msm(a + b)), from/range splitting, ifft(fft(x)) == x, coset round trip, and linearity of the
    transform -- the field additions come from the oracle's scalar C port on a thin sample, the group additions
    from the device's own g1_sum, itself pinned by the small-size tests.
"""
import os

import numpy as np
import pytest

import inputs
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

FULL = 1 << 20


@pytest.fixture(scope="module")
def bbg():
    import bbg as _bbg
    _bbg.init(0)
    return _bbg


@pytest.fixture(scope="module")
def full_srs(bbg):
    if not os.path.exists(os.path.join(po.REF_SRS_DIR, "transcript00.dat")):
        pytest.skip("oracle/_ref/srs_db/transcript00.dat (the reference's 2^20-point SRS) did not travel")
    return bbg.Pippenger.from_path(po.REF_SRS_DIR, FULL)


def np_reduce_once(a):
    """reduce_once (field_impl.hpp:268-294) on raw limbs, vectorised: v in [0, 2r) -> v - r if v >= r."""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    mod = [np.uint64((inputs.FR_MODULUS >> (64 * j)) & 0xFFFFFFFFFFFFFFFF) for j in range(4)]
    ge = np.zeros(a.shape[0], dtype=bool)
    decided = np.zeros(a.shape[0], dtype=bool)
    for j in (3, 2, 1, 0):
        ge |= ~decided & (a[:, j] > mod[j])
        decided |= a[:, j] != mod[j]
    ge |= ~decided
    borrow = np.zeros(a.shape[0], dtype=np.uint64)
    red = np.empty_like(a)
    for j in range(4):
        d = a[:, j] - mod[j]
        b1 = (a[:, j] < mod[j]).astype(np.uint64)
        d2 = d - borrow
        b2 = (d < borrow).astype(np.uint64)
        red[:, j] = d2
        borrow = b1 + b2
    out = a.copy()
    out[ge] = red[ge]
    return out


def np_fr_add(a, b):
    """(a + b) mod r on raw limbs of canonical inputs (< r < 2^254, so no carry out of limb 3), vectorised."""
    out = np.empty_like(a)
    carry = np.zeros(a.shape[0], dtype=np.uint64)
    for j in range(4):
        s = a[:, j] + b[:, j]
        c1 = (s < a[:, j]).astype(np.uint64)
        s2 = s + carry
        c2 = (s2 < s).astype(np.uint64)
        out[:, j] = s2
        carry = c1 + c2
    return np_reduce_once(out)


def test_np_fr_add_matches_oracle(orc):
    a = inputs.fr_elements(1, 512)
    b = inputs.fr_elements(2, 512)
    exp = orc.reduce(po.FR, orc.field_op(po.FR, 1, a, b))
    assert np.array_equal(np_fr_add(a, b), exp)
    c = inputs.fr_elements(3, 512, coarse_fraction=0.5)
    assert np.array_equal(np_reduce_once(c), orc.reduce(po.FR, c))


def test_msm_2p20_matches_reference_cpu(bbg, orc, full_srs):
    """config #2: BN254 G1 Pippenger MSM 2^20, bit-exact vs the reference's CPU path."""
    if not po.Ref.available():
        pytest.skip("oracle/_ref/libbbref.so not built")
    ref = po.Ref()
    pts = ref.read_transcript_g1(FULL, po.REF_SRS_DIR)
    table = ref.point_table(pts)
    for seed, n in ((7, FULL), (8, FULL - 3), (9, (1 << 19) + 12345)):
        sc = inputs.fr_elements(seed, n, coarse_fraction=0.001)
        exp = ref.jac_to_buffer(ref.pippenger(sc, table, n=n, unsafe=True, copy=False))
        got = orc.jac_to_buffer(full_srs.pippenger_unsafe(sc, 0, n))
        assert got == exp, (seed, n)


def test_msm_2p20_linearity_and_ranges(bbg, orc, full_srs):
    a = inputs.fr_elements(21, FULL)
    b = inputs.fr_elements(22, FULL)
    ra = full_srs.pippenger_unsafe(a, 0, FULL)
    rb = full_srs.pippenger_unsafe(b, 0, FULL)
    rab = full_srs.pippenger_unsafe(np_fr_add(a, b), 0, FULL)
    lhs = orc.jac_to_buffer(bbg.g1_sum(np.stack([ra, rb])))
    assert lhs == orc.jac_to_buffer(rab)
    # from/range splitting (Pippenger::pippenger_unsafe(scalars, from, range) + g1_sum, c_bind.cpp:40-45)
    cuts = [0, 1, 333_333, 1 << 19, FULL - 1, FULL]
    parts = [full_srs.pippenger_unsafe(a[lo:hi], lo, hi - lo) for lo, hi in zip(cuts[:-1], cuts[1:])]
    assert orc.jac_to_buffer(bbg.g1_sum(np.stack(parts))) == orc.jac_to_buffer(ra)


def test_msm_pinned_scalars_piecewise_upload(bbg, orc, full_srs):
    """Pinned host scalars take the piecewise upload (copy stream + histogram chasing the pieces, api.cu
    msm_host_scalars); pageable ones the plain path.  Same bytes in, same point out -- also for ragged sizes whose last
    piece is short or empty."""
    for seed, n in ((31, FULL), (32, FULL - 1000), (33, (1 << 18) + 1), (34, (3 << 18) + 77)):
        sc = inputs.fr_elements(seed, n, coarse_fraction=0.01)
        pinned = bbg.pinned_empty((n, 4))
        pinned[...] = sc
        a = orc.jac_to_buffer(full_srs.pippenger_unsafe(sc, 0, n))
        b = orc.jac_to_buffer(full_srs.pippenger_unsafe(pinned, 0, n))
        assert a == b, (seed, n)
        bbg.pinned_free(pinned)


def test_msm_2p20_structured_scalars(bbg, orc, full_srs):
    """All-equal scalars put every window's digits in ONE bucket per window (worst-case skew at full size):
    msm(k, ..., k) == k * msm(1, ..., 1)."""
    one = orc.to_mont(po.FR, [1])[0]
    k_int = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
    k = orc.to_mont(po.FR, [k_int])[0]
    ones = np.tile(one, (FULL, 1))
    ks = np.tile(k, (FULL, 1))
    s1 = full_srs.pippenger_unsafe(ones, 0, FULL)
    sk = full_srs.pippenger_unsafe(ks, 0, FULL)
    exp = orc.g1_mul(orc.g1_to_affine(s1), k)
    assert orc.jac_to_buffer(sk) == orc.jac_to_buffer(exp)


def test_msm_2p24_linearity_and_sharding_synthetic_points(bbg, orc):
    """North-star sizes beyond the shipped SRS (2^22 ... 2^26): bases are the SRS followed by distinct synthetic points
    P_i + D_k built on the device (never replicate points), exactly as bench.py does.  No CPU checker runs
    at this size in seconds, so the evidence is structural: linearity in the scalars, and the reference's own
    from/range + g1_sum sharding (what the multi-GPU path does) against the single MSM."""
    import torch
    if not os.path.exists(os.path.join(po.REF_SRS_DIR, "transcript00.dat")):
        pytest.skip("full SRS did not travel")
    lg = 24
    n = 1 << lg
    srs = bbg.read_transcript_g1(FULL, po.REF_SRS_DIR)
    dev = torch.device("cuda", 0)
    srs_dev = torch.from_numpy(srs.view(np.int64)).to(dev)
    pts = torch.empty((n, 8), dtype=torch.int64, device=dev)
    for blk in range(n // FULL):
        if blk == 0:
            pts[:FULL] = srs_dev
        else:
            bbg.g1_add_affine(srs_dev, srs[blk], out_dev=pts[blk * FULL:(blk + 1) * FULL])
    torch.cuda.synchronize()
    pip = bbg.Pippenger.from_device_points(pts, n)
    del pts, srs_dev
    a = inputs.fr_elements(61, n)
    b = inputs.fr_elements(62, n)
    ra = pip.pippenger_unsafe(a, 0, n)
    rb = pip.pippenger_unsafe(b, 0, n)
    rab = pip.pippenger_unsafe(np_fr_add(a, b), 0, n)
    assert orc.jac_to_buffer(bbg.g1_sum(np.stack([ra, rb]))) == orc.jac_to_buffer(rab)
    cuts = [0, n // 8, n // 2 + 3, n - (1 << 20) - 1, n]
    parts = [pip.pippenger_unsafe(a[lo:hi], lo, hi - lo) for lo, hi in zip(cuts[:-1], cuts[1:])]
    assert orc.jac_to_buffer(bbg.g1_sum(np.stack(parts))) == orc.jac_to_buffer(ra)
    # the first 2^20 bases are the real SRS: that range must agree with the SRS-only object's result
    base = bbg.Pippenger.from_path(po.REF_SRS_DIR, FULL)
    assert orc.jac_to_buffer(pip.pippenger_unsafe(a[:FULL], 0, FULL)) == orc.jac_to_buffer(base.pippenger_unsafe(a[:FULL], 0, FULL))
    pip.close()
    base.close()


@pytest.mark.parametrize("lg", [20, 22])
def test_ntt_fullsize_matches_reference_cpu(bbg, orc, lg):
    """config #3: fr radix-2 NTT / iNTT / coset_fft at 2^22 vs the reference's CPU polynomial_arithmetic."""
    if not po.Ref.available():
        pytest.skip("oracle/_ref/libbbref.so not built")
    ref = po.Ref()
    n = 1 << lg
    x = inputs.fr_elements(300 + lg, n, coarse_fraction=0.01)
    for kind in (po.NTT_FFT, po.NTT_IFFT, po.NTT_COSET_FFT, po.NTT_COSET_IFFT):
        exp = np_reduce_once(ref.ntt(kind, x))
        got = np_reduce_once(bbg.ntt(x.copy(), kind))
        assert np.array_equal(got, exp), (lg, kind)


@pytest.mark.parametrize("lg", [22, 24])
def test_ntt_fullsize_round_trips_and_linearity(bbg, orc, lg):
    n = 1 << lg
    x = inputs.fr_elements(400 + lg, n)
    y = inputs.fr_elements(500 + lg, n)
    # outputs are any representative in [0, 2r) like the reference's: compare reduce_once'd limbs
    fx_raw = bbg.fft(x.copy())
    assert int(fx_raw[:, 3].max()) < 2 * inputs.FR_TOP + 2
    fx = np_reduce_once(fx_raw)
    assert np.array_equal(np_reduce_once(bbg.ifft(fx_raw.copy())), x)
    cx = bbg.coset_fft(x.copy())
    assert np.array_equal(np_reduce_once(bbg.coset_ifft(cx.copy())), x)
    # linearity: fft(x + y) == fft(x) + fft(y)
    fy = np_reduce_once(bbg.fft(y.copy()))
    assert np.array_equal(np_reduce_once(bbg.fft(np_fr_add(x, y))), np_fr_add(fx, fy))
    # a thin sample of the outputs against the definition X[k] = sum_i x[i] w^(ik), via the oracle's evaluate()
    w = orc.fr_root_of_unity(lg)
    for kidx in (0, 1, n // 2 + 5):
        z = pow_mont(orc, w, kidx)
        exp = orc.reduce(po.FR, orc.evaluate(x, z))
        assert np.array_equal(fx[kidx], exp.reshape(4)), kidx


def pow_mont(orc, base, e):
    acc = orc.to_mont(po.FR, [1])[0]
    b = np.array(base, dtype=np.uint64)
    while e:
        if e & 1:
            acc = orc.field_op(po.FR, 0, acc, b)[0]
        b = orc.field_op(po.FR, 0, b, b)[0]
        e >>= 1
    return acc
