"""Host-side unit test of csrc/staging.hpp's CopyPool: every byte copied exactly once for sizes around the slice
boundaries, with 1..8 threads, repeatedly on the same pool (generation counter), and pool teardown joins cleanly."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    if not os.path.isdir(CUDA_INC):
        pytest.skip("CUDA headers not found")
    so = str(tmp_path_factory.mktemp("copypool") / "libcopypool.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-I", CUDA_INC, "-o", so,
                           os.path.join(HERE, "hostlib", "copypool_host.cpp")])
    l = ctypes.CDLL(so)
    l.pool_new.restype = ctypes.c_void_p
    l.pool_new.argtypes = [ctypes.c_uint]
    l.pool_delete.argtypes = [ctypes.c_void_p]
    l.pool_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    return l


@pytest.mark.parametrize("threads", [1, 2, 3, 8])
def test_parallel_copy_is_exact(lib, threads):
    pool = lib.pool_new(threads)
    rng = np.random.default_rng(threads)
    try:
        for size in (0, 1, 4095, 4096, 4097, (256 << 10) - 1, 256 << 10, (256 << 10) + 1, (1 << 20) + 12345, (4 << 20), (4 << 20) - 7):
            src = rng.integers(0, 256, size=size + 64, dtype=np.uint8)
            dst = np.full(size + 64, 0xAB, dtype=np.uint8)
            lib.pool_copy(pool, dst.ctypes.data + 32, src.ctypes.data + 32, size)  # unaligned on purpose
            assert np.array_equal(dst[32:32 + size], src[32:32 + size]), size
            assert (dst[:32] == 0xAB).all() and (dst[32 + size:] == 0xAB).all(), size  # nothing outside the range
    finally:
        lib.pool_delete(pool)


def test_many_small_jobs_reuse_the_pool(lib):
    pool = lib.pool_new(4)
    try:
        src = np.arange(300 << 10, dtype=np.uint8)
        for i in range(200):
            dst = np.zeros_like(src)
            lib.pool_copy(pool, dst.ctypes.data, src.ctypes.data, src.size)
            assert np.array_equal(dst, src), i
    finally:
        lib.pool_delete(pool)
