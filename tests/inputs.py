"""Deterministic synthetic inputs shared by the golden generator, the CPU tests, the GPU parity
tests, smoke() and bench.py.  Explicit numpy PRNG -> limbs, so the oracle, the reference and the
CUDA path all see byte-identical arrays (SURVEY.md Appendix B caveat: the reference's own debug
RNG is compiler-dependent, so we never rely on it for inputs)."""
import os

import numpy as np

FR_TOP = 0x30644E72E131A029  # top limb of both BN254 moduli
FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SRS_MINI_DIR = os.path.join(GOLDEN_DIR, "srs_mini")
SRS_MINI_POINTS = 4096


def fr_elements(seed, n, coarse_fraction=0.0):
    """n uniform fr elements as raw Montgomery-form limbs (n, 4) uint64.

    Any value < r is the Montgomery form of exactly one field element, so drawing the limbs
    uniformly (top limb strictly below the modulus' top limb) is a uniform draw over fr minus a
    2^-62 sliver.  With coarse_fraction > 0 that share of the elements gets +r added, i.e. lands in
    the reference's legal non-canonical range [r, 2r) (callers do pass those).
    """
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] %= np.uint64(FR_TOP - 1)
    if coarse_fraction > 0.0 and n:
        pick = np.nonzero(rng.random(n) < coarse_fraction)[0]
        for i in pick.tolist():
            v = sum(int(a[i, j]) << (64 * j) for j in range(4)) + FR_MODULUS
            for j in range(4):
                a[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return a


def short_scalar_ints(seed, n):
    """The reference's short-input mix (scalar_multiplication.test.cpp:734-755): 128-, 64-, 3-bit and zero scalars."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        w = rng.integers(0, 1 << 64, size=2, dtype=np.uint64)
        k = i % 5
        if k == 0:
            out.append(int(w[0]) | (int(w[1]) << 64))
        elif k == 1:
            out.append(int(w[0]))
        elif k == 2:
            out.append(int(w[0]) & 7)
        elif k == 3:
            out.append(0)
        else:
            out.append(1)
    return out
