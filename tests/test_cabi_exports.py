"""The drop-in boundary on a box without a GPU: libbbg.so loads, exports every function include/bbg.h declares, the
ctypes binding's export list is the header's, and compute entry points fail LOUDLY (no CPU fallback) when there is no
CUDA device."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bbg.h")
LIB = os.path.join(ROOT, "aztec-2.0_b200", "libbbg.so")


def declared_functions():
    with open(HEADER) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)  # comments mention functions that are not declarations
    text = re.sub(r"//[^\n]*", " ", text)
    return sorted(set(re.findall(r"\b(bbg_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.fail("aztec-2.0_b200/libbbg.so is not built (run __graft_entry__.build())")
    return ctypes.CDLL(LIB)


def test_header_declares_the_expected_surface():
    names = declared_functions()
    assert len(names) >= 45
    for must in ("bbg_pippenger", "bbg_pippenger_unsafe", "bbg_new_pippenger", "bbg_delete_pippenger", "bbg_g1_sum",
                 "bbg_generate_pippenger_point_table", "bbg_read_transcript_g1", "bbg_ntt", "bbg_coset_fft_ext", "bbg_ifft",
                 "bbg_coset_fft_with_generator_shift", "bbg_new_evaluation_domain", "bbg_malloc", "bbg_free"):
        assert must in names, must


def test_library_exports_every_declared_function(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_dynamic_symbol_table_has_no_undeclared_bbg_entry_points():
    """Everything exported with the bbg_ prefix is part of the documented C-ABI (-fvisibility=hidden keeps the rest in)."""
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    exported = sorted({l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("bbg_")})
    assert exported == declared_functions()


def test_python_binding_export_list_matches_header():
    sys.path.insert(0, os.path.join(ROOT, "aztec-2.0_b200", "python"))
    import bbg
    assert sorted(bbg.EXPORTS) == declared_functions()


def _no_gpu(lib):
    lib.bbg_device_count.restype = ctypes.c_int
    return lib.bbg_device_count() <= 0


def test_compute_entry_points_fail_loudly_without_a_device(lib):
    if not _no_gpu(lib):
        pytest.skip("a CUDA device is present: the failure path is not reachable here")
    lib.bbg_last_error.restype = ctypes.c_char_p
    x = np.zeros((16, 4), dtype=np.uint64)
    rc = lib.bbg_ntt(ctypes.c_void_p(x.ctypes.data), ctypes.c_size_t(16), 0, ctypes.c_size_t(0), None)
    assert rc != 0, "bbg_ntt must not succeed without a device (there is no CPU fallback)"
    assert b"CUDA" in (lib.bbg_last_error() or b"") or b"device" in (lib.bbg_last_error() or b"")
    out = np.zeros(12, dtype=np.uint64)
    pts = np.zeros((4, 8), dtype=np.uint64)
    rc = lib.bbg_msm_points(ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(pts.ctypes.data), ctypes.c_size_t(4),
                            ctypes.c_void_p(out.ctypes.data))
    assert rc != 0
    assert not out.any(), "no result may be written by a failed call"
    lib.bbg_new_pippenger_from_points.restype = ctypes.c_void_p
    assert not lib.bbg_new_pippenger_from_points(ctypes.c_void_p(pts.ctypes.data), ctypes.c_size_t(4))


def test_host_side_domain_constants_and_layout_need_no_device():
    """Pure host code of the library (csrc/host_field.hpp): evaluation_domain constants (root, root^-1, n, n^-1, g, g^-1;
    bb/polynomials/evaluation_domain.cpp:57-76) against the oracle for every power of two up to fr's 2-adicity, the golden
    vectors, and the multi-GPU NTT layout contract."""
    sys.path.insert(0, os.path.join(ROOT, "aztec-2.0_b200", "python"))
    sys.path.insert(0, ROOT)
    import json
    import bbg
    from oracle import pyoracle as po
    orc = po.Oracle()
    for lg in range(1, 29):
        a = bbg.domain_constants(1 << lg)
        b = orc.domain_constants(1 << lg)
        assert np.array_equal(orc.reduce(po.FR, a), orc.reduce(po.FR, b)), lg
    with pytest.raises(bbg.BbgError):
        bbg.domain_constants(3)
    with open(os.path.join(ROOT, "tests", "golden", "vectors.json")) as f:
        golden = json.load(f)
    for lg in (12, 16, 20, 22, 24, 25, 26, 28):
        for world in (2, 4, 8):
            in_pos, out_pos = bbg.ntt_dist_layout(1 << lg, world)
            passes = 2 if lg <= 16 else (3 if lg <= 24 else 4)
            first = lg // passes + (1 if lg % passes else 0)
            last = lg // passes + (1 if (passes - 1) < lg % passes else 0)
            rb = world.bit_length() - 1
            assert (in_pos, out_pos) == (last - rb, first - rb), (lg, world)
    with pytest.raises(bbg.BbgError):
        bbg.ntt_dist_layout(1 << 12, 3)
    assert "fields" in golden  # the fixture the GPU parity tests read is present in the tree
