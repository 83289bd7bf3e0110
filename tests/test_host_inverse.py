"""Host-side unit test of csrc/inv.cuh: the branch-free Kaliski almost-inverse and its Montgomery fix-up table, compiled
with g++ from the very header the CUDA kernels include, checked against Python big-integer arithmetic."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle.pyoracle import FQ_MODULUS, FR_MODULUS

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("invhost") / "libinvhost.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "hostlib", "inv_host.cpp")])
    return ctypes.CDLL(so)


def limbs(v):
    return np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)


def unlimbs(a):
    return sum(int(x) << (32 * i) for i, x in enumerate(a))


@pytest.mark.parametrize("p", [FQ_MODULUS, FR_MODULUS])
def test_almost_inverse(lib, p):
    rng = random.Random(7)
    pl = limbs(p)
    cases = [1, 2, 3, p - 1, p - 2, (p + 1) // 2, 1 << 253, (1 << 253) + 1] + [rng.randrange(1, p) for _ in range(3000)]
    cases += [rng.randrange(1, 1 << b) for b in (8, 31, 32, 33, 64, 65, 128, 200) for _ in range(20)]
    ks = []
    for a in cases:
        out = np.zeros(8, dtype=np.uint32)
        k = lib.inv_almost(limbs(a).ctypes.data_as(ctypes.c_void_p), pl.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        assert 254 <= k <= 508
        r = unlimbs(out)
        assert r < p and r == pow(a, -1, p) * pow(2, k, p) % p, hex(a)
        ks.append(k)
    assert max(ks) <= 508 and min(ks) >= 254


@pytest.mark.parametrize("p", [FQ_MODULUS, FR_MODULUS])
def test_fix_table(lib, p):
    R = 1 << 256
    n = lib.inv_fix_entries()
    kmin = lib.inv_k_min()
    tab = np.zeros(n * 8, dtype=np.uint32)
    lib.inv_fix_table(limbs(p).ctypes.data_as(ctypes.c_void_p), limbs(R * R % p).ctypes.data_as(ctypes.c_void_p), tab.ctypes.data_as(ctypes.c_void_p))
    for i in range(n):
        k = kmin + i
        assert unlimbs(tab[8 * i:8 * i + 8]) == pow(R, 3, p) * pow(2, -k, p) % p
    # the whole pipeline: Montgomery form in, Montgomery form of the inverse out
    rng = random.Random(11)
    for _ in range(200):
        d = rng.randrange(1, p)
        a = d * R % p
        out = np.zeros(8, dtype=np.uint32)
        k = lib.inv_almost(limbs(a).ctypes.data_as(ctypes.c_void_p), limbs(p).ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        c = unlimbs(tab[8 * (k - kmin):8 * (k - kmin) + 8])
        mont = unlimbs(out) * c * pow(R, -1, p) % p  # montmul(out, C_k)
        assert mont == pow(d, -1, p) * R % p
