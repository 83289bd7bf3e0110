"""bbg -- thin ctypes binding over libbbg.so (include/bbg.h), used by the tests, smoke() and bench.py.

The product is the C-ABI library (hand-written sm_100a CUDA behind barretenberg's call signatures);
this module only marshals numpy arrays (host pointers) and torch CUDA tensors (device pointers) into
it.  Names mirror barretenberg's: ``Pippenger`` (bb/ecc/curves/bn254/scalar_multiplication/pippenger.hpp:35-52),
``pippenger`` / ``pippenger_unsafe`` (scalar_multiplication.hpp:139-148), ``fft`` / ``ifft`` / ``coset_fft`` /
``coset_ifft`` / ``*_with_constant`` / ``coset_fft_with_generator_shift`` (bb/polynomials/polynomial_arithmetic.hpp:23-39),
``g1_sum`` (c_bind.cpp:40-45).  Errors raise ``BbgError`` carrying the library's message, like the
reference's throw_or_abort.

There is NO CPU fallback: if libbbg.so is missing the import fails; without a CUDA device every
compute call raises BbgError(BBG_ERR_NO_DEVICE).

Layouts (numpy uint64): fr/fq (..., 4); g1 affine (..., 8); g1 Jacobian (..., 12) -- barretenberg's.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "libbbg.so"))

OK, ERR_CUDA, ERR_ARG, ERR_NO_DEVICE, ERR_SRS, ERR_IO = range(6)
(FFT, IFFT, COSET_FFT, COSET_IFFT, FFT_WITH_CONSTANT, IFFT_WITH_CONSTANT, COSET_FFT_WITH_CONSTANT,
 COSET_FFT_WITH_GENERATOR_SHIFT) = range(8)

EXPORTS = [
    "bbg_ntt_dist_layout", "bbg_ntt_dist_dev", "bbg_pippenger_bind_host_table", "bbg_set_auto_adopt", "bbg_bench_field_mul", "bbg_g1_add_affine_dev", "bbg_profile", "bbg_profile_read", "bbg_init", "bbg_shutdown", "bbg_last_error", "bbg_device_count", "bbg_kernel_launches", "bbg_last_device_ms",
    "bbg_malloc", "bbg_free", "bbg_new_pippenger", "bbg_new_pippenger_from_path", "bbg_new_pippenger_from_table",
    "bbg_new_pippenger_from_points", "bbg_new_pippenger_from_device_points", "bbg_delete_pippenger", "bbg_pippenger_num_points", "bbg_pippenger_get_point_table",
    "bbg_pippenger_device_points", "bbg_pippenger_window_bits", "bbg_pippenger_levels", "bbg_pippenger_unsafe", "bbg_pippenger_unsafe_dev", "bbg_pippenger", "bbg_msm_points",
    "bbg_msm_points_dev", "bbg_generate_pippenger_point_table", "bbg_g1_sum", "bbg_g1_sum_dev", "bbg_read_transcript_g1",
    "bbg_read_g1_elements_from_buffer", "bbg_ntt", "bbg_ntt_dev", "bbg_coset_fft_ext", "bbg_coset_fft_ext_dev",
    "bbg_new_evaluation_domain", "bbg_delete_evaluation_domain", "bbg_ifft", "bbg_coset_fft_with_generator_shift",
    "bbg_domain_constants", "bbg_field_op", "bbg_g1_op",
    "bbg_pippenger_unsafe_batch", "bbg_pippenger_unsafe_batch_dev", "bbg_pippenger_batch",
    "bbg_field_op_dev", "bbg_g1_normalize", "bbg_resident_mode", "bbg_ntt_ex", "bbg_wire_coset_fft", "bbg_turbo_quotient",
    "bbg_permutation_quotient", "bbg_divide_by_pseudo_vanishing_polynomial", "bbg_compute_lagrange_polynomial_fft",
    "bbg_permutation_grand_product", "bbg_evaluate", "bbg_compute_opening_polynomial", "bbg_poly_write", "bbg_linear_combination", "bbg_evaluate_batch", "bbg_wire_ifft", "bbg_stats_totals", "bbg_ntt_dist_fused_dev", "bbg_ntt_dist_natural_dev", "bbg_wire_ifft_batch",
    "bbg_peer_buffer_alloc", "bbg_peer_buffer_open", "bbg_peer_buffer_close", "bbg_peer_buffer_free", "bbg_resident_invalidate", "bbg_resident_flush", "bbg_resident_stats",
]


class BbgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbbg error %d: %s" % (code, msg))
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError("libbbg.so not built (run `make -C aztec-2.0_b200/csrc` or __graft_entry__.build()); "
                      "there is no CPU fallback")
lib = ctypes.CDLL(LIB_PATH)

_vp, _sz, _int = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
lib.bbg_last_error.restype = ctypes.c_char_p
lib.bbg_kernel_launches.restype = ctypes.c_uint64
lib.bbg_last_device_ms.restype = ctypes.c_double
lib.bbg_malloc.restype = _vp
lib.bbg_malloc.argtypes = [_sz]
lib.bbg_free.argtypes = [_vp]
for _n in ("bbg_new_pippenger", "bbg_new_pippenger_from_table", "bbg_new_pippenger_from_points",
           "bbg_new_pippenger_from_device_points"):
    getattr(lib, _n).restype = _vp
    getattr(lib, _n).argtypes = [_vp, _sz]
lib.bbg_new_pippenger_from_path.restype = _vp
lib.bbg_new_pippenger_from_path.argtypes = [ctypes.c_char_p, _sz]
lib.bbg_delete_pippenger.argtypes = [_vp]
lib.bbg_pippenger_num_points.restype = _sz
lib.bbg_pippenger_num_points.argtypes = [_vp]
lib.bbg_pippenger_get_point_table.argtypes = [_vp, _vp]
lib.bbg_pippenger_device_points.restype = _vp
lib.bbg_pippenger_device_points.argtypes = [_vp]
lib.bbg_pippenger_window_bits.restype = ctypes.c_uint
lib.bbg_pippenger_window_bits.argtypes = [_vp]
lib.bbg_pippenger_levels.restype = ctypes.c_uint
lib.bbg_pippenger_levels.argtypes = [_vp]
lib.bbg_pippenger_unsafe.argtypes = [_vp, _vp, _sz, _sz, _vp]
lib.bbg_pippenger_unsafe_dev.argtypes = [_vp, _vp, _sz, _sz, _vp, _vp]
lib.bbg_pippenger.argtypes = [_vp, _vp, _sz, _int, _vp]
lib.bbg_msm_points.argtypes = [_vp, _vp, _sz, _vp]
lib.bbg_msm_points_dev.argtypes = [_vp, _vp, _sz, _sz, _vp, _vp]
lib.bbg_generate_pippenger_point_table.argtypes = [_vp, _vp, _sz]
lib.bbg_g1_sum.argtypes = [_vp, _sz, _vp]
lib.bbg_g1_sum_dev.argtypes = [_vp, _sz, _vp, _vp]
lib.bbg_read_transcript_g1.argtypes = [_vp, _sz, ctypes.c_char_p]
lib.bbg_read_g1_elements_from_buffer.argtypes = [_vp, _vp, _sz]
lib.bbg_ntt.argtypes = [_vp, _sz, _int, _sz, _vp]
lib.bbg_ntt_dev.argtypes = [_vp, _sz, _int, _sz, _vp, _vp]
lib.bbg_coset_fft_ext.argtypes = [_vp, _sz, _sz]
lib.bbg_coset_fft_ext_dev.argtypes = [_vp, _sz, _sz, _vp]
lib.bbg_new_evaluation_domain.restype = _vp
lib.bbg_new_evaluation_domain.argtypes = [_sz]
lib.bbg_delete_evaluation_domain.argtypes = [_vp]
lib.bbg_ifft.argtypes = [_vp, _vp]
lib.bbg_coset_fft_with_generator_shift.argtypes = [_vp, _vp, _vp]
lib.bbg_domain_constants.argtypes = [_sz, _vp]
lib.bbg_field_op.argtypes = [_int, _int, _vp, _vp, _vp, _sz]
lib.bbg_g1_op.argtypes = [_int, _vp, _vp, _vp, _sz]
lib.bbg_init.argtypes = [_int]
lib.bbg_profile.argtypes = [_int]
lib.bbg_profile_read.argtypes = [_vp, _int]
lib.bbg_ntt_dist_layout.argtypes = [_sz, _int, _vp, _vp]
lib.bbg_ntt_dist_dev.argtypes = [_vp, _vp, _sz, _int, _sz, _vp, _int, _int, _int, _vp]
lib.bbg_pippenger_bind_host_table.argtypes = [_vp, _vp]
lib.bbg_set_auto_adopt.argtypes = [_int]
lib.bbg_bench_field_mul.argtypes = [_int, _int, _vp]
lib.bbg_g1_add_affine_dev.argtypes = [_vp, _vp, _sz, _vp, _vp]
lib.bbg_pippenger_unsafe_batch.argtypes = [_vp, _vp, _sz, _sz, _sz, _vp]
lib.bbg_pippenger_unsafe_batch_dev.argtypes = [_vp, _vp, _sz, _sz, _sz, _vp, _vp]
lib.bbg_pippenger_batch.argtypes = [_vp, _sz, _vp, _sz, _vp]
lib.bbg_field_op_dev.argtypes = [_int, _int, _vp, _vp, _vp, _sz, _vp]
lib.bbg_g1_normalize.argtypes = [_vp, _sz, _vp]
lib.bbg_ntt_ex.argtypes = [_vp, _sz, _int, _sz, _vp, ctypes.c_uint]
lib.bbg_wire_coset_fft.argtypes = [_vp, _vp, _sz, _sz, ctypes.c_uint]
lib.bbg_turbo_quotient.argtypes = [_int, _vp, _sz, _vp, _vp, _vp, ctypes.c_uint]
lib.bbg_permutation_quotient.argtypes = [_vp, _vp, ctypes.c_uint, _vp, _vp, _sz, ctypes.c_uint, _vp, _vp, _vp, _vp, _vp, ctypes.c_uint]
lib.bbg_divide_by_pseudo_vanishing_polynomial.argtypes = [_vp, _sz, _sz, ctypes.c_uint, ctypes.c_uint]
lib.bbg_compute_lagrange_polynomial_fft.argtypes = [_vp, _sz, _sz]
lib.bbg_permutation_grand_product.argtypes = [_vp, _vp, ctypes.c_uint, _sz, _vp, _vp, _vp, ctypes.c_uint]
lib.bbg_evaluate.argtypes = [_vp, _sz, _vp, _vp]
lib.bbg_compute_opening_polynomial.argtypes = [_vp, _vp, _vp, _sz, _sz, _vp, ctypes.c_uint]
lib.bbg_poly_write.argtypes = [_vp, _sz, _vp, _sz]
lib.bbg_stats_totals.argtypes = [_vp]
lib.bbg_ntt_dist_fused_dev.argtypes = [_vp, _vp, _vp, _sz, _int, _sz, _vp, _int, _int, _vp]
lib.bbg_ntt_dist_natural_dev.argtypes = [_vp, _vp, _vp, _vp, _sz, _int, _sz, _vp, _int, _int, _int, _vp]
lib.bbg_wire_ifft_batch.argtypes = [_vp, _sz, _vp, _sz, ctypes.c_uint]
lib.bbg_peer_buffer_alloc.argtypes = [_sz, _vp, _vp]
lib.bbg_peer_buffer_open.argtypes = [_vp, _vp]
lib.bbg_peer_buffer_close.argtypes = [_vp]
lib.bbg_peer_buffer_free.argtypes = [_vp]
lib.bbg_wire_ifft.argtypes = [_vp, _sz, _vp, ctypes.c_uint]
lib.bbg_evaluate_batch.argtypes = [_vp, _vp, _sz, _vp, _vp]
lib.bbg_linear_combination.argtypes = [_vp, _vp, _vp, _vp, _sz, _sz, ctypes.c_uint]
lib.bbg_resident_mode.argtypes = [_int]
lib.bbg_resident_invalidate.argtypes = [_vp, _sz]
lib.bbg_resident_flush.argtypes = [_vp, _sz]
lib.bbg_resident_stats.argtypes = [_vp]
NUM_PHASES = 13
PHASE_NAMES = ["msm_digits", "msm_scan", "msm_scatter", "msm_pairs", "msm_accumulate", "msm_fixup", "msm_reduce", "msm_combine",
               "ntt_tables", "ntt_pass0", "ntt_pass1", "ntt_pass2", "ntt_pass3"]


def _check(rc):
    if rc != OK:
        raise BbgError(rc, (lib.bbg_last_error() or b"").decode(errors="replace"))


def _np(a, width=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if width is not None:
        a = a.reshape(-1, width)
    return a


def _ptr(a):
    """host pointer of a numpy array, or device pointer of a torch CUDA tensor"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


def _stream_ptr(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    return stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)


def init(device=-1):
    _check(lib.bbg_init(device))


def shutdown():
    lib.bbg_shutdown()


def device_count():
    return int(lib.bbg_device_count())


def kernel_launches():
    return int(lib.bbg_kernel_launches())


def last_device_ms():
    return float(lib.bbg_last_device_ms())


def profile(enable):
    _check(lib.bbg_profile(1 if enable else 0))


def profile_read():
    """{phase name: milliseconds} of the last compute call (device time, CUDA events)."""
    ms = (ctypes.c_double * NUM_PHASES)()
    _check(lib.bbg_profile_read(ctypes.cast(ms, ctypes.c_void_p), NUM_PHASES))
    return {PHASE_NAMES[i]: float(ms[i]) for i in range(NUM_PHASES)}


def bench_field_mul(field=0, iters=2000):
    """Sustained register-resident Montgomery multiplies per second over the whole GPU (integer-pipe roofline)."""
    out = ctypes.c_double(0.0)
    _check(lib.bbg_bench_field_mul(field, iters, ctypes.cast(ctypes.pointer(out), ctypes.c_void_p)))
    return out.value


def g1_add_affine(points_dev, q_affine, out_dev=None, stream=None):
    """out[i] = affine(points[i] + q) on torch CUDA tensors holding 64-byte affine points (synthetic bases)."""
    out_dev = points_dev if out_dev is None else out_dev
    q = _np(q_affine, 8)
    n = points_dev.numel() * points_dev.element_size() // 64
    _check(lib.bbg_g1_add_affine_dev(points_dev.data_ptr(), out_dev.data_ptr(), n, q.ctypes.data, _stream_ptr(stream)))
    return out_dev


def pinned_empty(shape, dtype=np.uint64):
    """numpy array over page-locked host memory from bbg_malloc (bbmalloc, c_bind.cpp:11-19)."""
    shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    p = lib.bbg_malloc(max(nbytes, 1))
    if not p:
        raise BbgError(ERR_CUDA, (lib.bbg_last_error() or b"").decode())
    buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        lib.bbg_free(p)


class Pippenger:
    """bb/.../pippenger.hpp:35-52: owns the SRS (here: resident in HBM as n affine points)."""

    def __init__(self, handle, keep=None):
        if not handle:
            raise BbgError(ERR_SRS, (lib.bbg_last_error() or b"").decode(errors="replace"))
        self.h = handle
        self._keep = keep

    @classmethod
    def from_path(cls, srs_dir, num_points):
        return cls(lib.bbg_new_pippenger_from_path(str(srs_dir).encode(), num_points))

    @classmethod
    def from_raw(cls, raw_points, num_points):
        raw = np.ascontiguousarray(np.frombuffer(raw_points, dtype=np.uint8) if not isinstance(raw_points, np.ndarray) else raw_points)
        return cls(lib.bbg_new_pippenger(raw.ctypes.data, num_points))

    @classmethod
    def from_table(cls, table2n, num_points):
        t = _np(table2n, 8)
        return cls(lib.bbg_new_pippenger_from_table(t.ctypes.data, num_points), keep=t)

    @classmethod
    def from_device_points(cls, points_dev, num_points=None):
        """points_dev: torch CUDA tensor of 64-byte affine points (copied device-to-device into the object)."""
        import torch
        n = points_dev.numel() * points_dev.element_size() // 64 if num_points is None else num_points
        host = torch.empty(0)
        del host
        obj = cls(lib.bbg_new_pippenger_from_device_points(points_dev.data_ptr(), n))
        return obj

    @classmethod
    def from_points(cls, points, num_points=None):
        p = _np(points, 8)
        n = p.shape[0] if num_points is None else num_points
        return cls(lib.bbg_new_pippenger_from_points(p.ctypes.data, n))

    def get_num_points(self):
        return int(lib.bbg_pippenger_num_points(self.h))

    def window_bits(self):
        return int(lib.bbg_pippenger_window_bits(self.h))

    def levels(self):
        return int(lib.bbg_pippenger_levels(self.h))

    def get_point_table(self):
        out = np.zeros((2 * self.get_num_points(), 8), dtype=np.uint64)
        _check(lib.bbg_pippenger_get_point_table(self.h, out.ctypes.data))
        return out

    def device_points(self):
        return int(lib.bbg_pippenger_device_points(self.h) or 0)

    def pippenger_unsafe(self, scalars, from_=0, range_=None, stream=None):
        """MSM over monomials [from, from+range).  numpy scalars -> numpy Jacobian (12,);
        torch CUDA scalars -> torch CUDA uint8[96] (asynchronous on `stream`)."""
        if isinstance(scalars, np.ndarray) or not hasattr(scalars, "data_ptr"):
            s = _np(scalars, 4)
            n = s.shape[0] if range_ is None else range_
            out = np.zeros(12, dtype=np.uint64)
            _check(lib.bbg_pippenger_unsafe(self.h, s.ctypes.data, from_, n, out.ctypes.data))
            return out
        import torch
        n = scalars.numel() * scalars.element_size() // 32 if range_ is None else range_
        out = torch.empty(96, dtype=torch.uint8, device=scalars.device)
        _check(lib.bbg_pippenger_unsafe_dev(self.h, scalars.data_ptr(), from_, n, out.data_ptr(), _stream_ptr(stream)))
        return out

    def pippenger_unsafe_batch(self, scalar_arrays, from_=0, range_=None, stream=None):
        """Several MSMs over monomials [from, from+range) in one call (bbg_pippenger_unsafe_batch).  A list of numpy
        arrays -> numpy (k, 12); a list of torch CUDA tensors -> torch CUDA uint8 (k, 96), asynchronous on `stream`."""
        k = len(scalar_arrays)
        if k == 0:
            return np.zeros((0, 12), dtype=np.uint64)
        if isinstance(scalar_arrays[0], np.ndarray) or not hasattr(scalar_arrays[0], "data_ptr"):
            arrs = [_np(a, 4) for a in scalar_arrays]
            n = arrs[0].shape[0] if range_ is None else range_
            ptrs = (ctypes.c_void_p * k)(*[a.ctypes.data for a in arrs])
            out = np.zeros((k, 12), dtype=np.uint64)
            _check(lib.bbg_pippenger_unsafe_batch(self.h, ctypes.cast(ptrs, _vp), k, from_, n, out.ctypes.data))
            return out
        import torch
        n = scalar_arrays[0].numel() * scalar_arrays[0].element_size() // 32 if range_ is None else range_
        ptrs = (ctypes.c_void_p * k)(*[a.data_ptr() for a in scalar_arrays])
        out = torch.empty((k, 96), dtype=torch.uint8, device=scalar_arrays[0].device)
        _check(lib.bbg_pippenger_unsafe_batch_dev(self.h, ctypes.cast(ptrs, _vp), k, from_, n, out.data_ptr(), _stream_ptr(stream)))
        return out

    def close(self):
        if self.h:
            lib.bbg_delete_pippenger(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pippenger(scalars, points_table2n, num_points, handle_edge_cases=True):
    s, t = _np(scalars, 4), _np(points_table2n, 8)
    out = np.zeros(12, dtype=np.uint64)
    _check(lib.bbg_pippenger(s.ctypes.data if num_points else None, t.ctypes.data if num_points else None, num_points,
                             1 if handle_edge_cases else 0, out.ctypes.data))
    return out


def resident_mode(enable):
    """Resident polynomials on / off (include/bbg.h); enable < 0 queries."""
    rc = lib.bbg_resident_mode(int(enable))
    if enable < 0:
        return rc
    _check(rc)


def stats_totals():
    out = (ctypes.c_uint64 * 12)()
    _check(lib.bbg_stats_totals(ctypes.cast(out, _vp)))
    return {name: {"calls": int(out[3 * i]), "h2d": int(out[3 * i + 1]), "d2h": int(out[3 * i + 2])} for i, name in enumerate(("msm", "ntt", "srs", "poly"))}


def resident_stats():
    out = (ctypes.c_uint64 * 4)()
    _check(lib.bbg_resident_stats(ctypes.cast(out, _vp)))
    return {"hits": int(out[0]), "misses": int(out[1]), "h2d_bytes_saved": int(out[2]), "bytes_resident": int(out[3])}


def resident_flush(arr=None):
    _check(lib.bbg_resident_flush(None if arr is None else arr.ctypes.data, 0 if arr is None else arr.nbytes))


def resident_invalidate(arr=None):
    _check(lib.bbg_resident_invalidate(None if arr is None else arr.ctypes.data, 0 if arr is None else arr.nbytes))


def pippenger_batch(scalar_arrays, points_table2n, num_points):
    """bbg_pippenger_batch: k MSMs over an ADOPTED interleaved table addressed by its host pointer."""
    k = len(scalar_arrays)
    arrs = [_np(a, 4) for a in scalar_arrays]
    t = points_table2n
    ptrs = (ctypes.c_void_p * max(k, 1))(*[a.ctypes.data for a in arrs])
    out = np.zeros((k, 12), dtype=np.uint64)
    _check(lib.bbg_pippenger_batch(ctypes.cast(ptrs, _vp), k, t.ctypes.data, num_points, out.ctypes.data))
    return out


def pippenger_unsafe(scalars, points_table2n, num_points):
    return pippenger(scalars, points_table2n, num_points, handle_edge_cases=False)


def msm_points(scalars, points, num_points=None):
    s, p = _np(scalars, 4), _np(points, 8)
    n = s.shape[0] if num_points is None else num_points
    out = np.zeros(12, dtype=np.uint64)
    _check(lib.bbg_msm_points(s.ctypes.data if n else None, p.ctypes.data if n else None, n, out.ctypes.data))
    return out


def generate_pippenger_point_table(points):
    p = _np(points, 8)
    out = np.zeros((2 * p.shape[0], 8), dtype=np.uint64)
    _check(lib.bbg_generate_pippenger_point_table(p.ctypes.data, out.ctypes.data, p.shape[0]))
    return out


def g1_sum(elements):
    e = _np(elements, 12)
    out = np.zeros(12, dtype=np.uint64)
    _check(lib.bbg_g1_sum(e.ctypes.data, e.shape[0], out.ctypes.data))
    return out


def read_transcript_g1(degree, srs_dir):
    out = np.zeros((degree, 8), dtype=np.uint64)
    _check(lib.bbg_read_transcript_g1(out.ctypes.data, degree, str(srs_dir).encode()))
    return out


def read_g1_elements_from_buffer(raw):
    raw = np.ascontiguousarray(np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw)
    n = raw.size // 64
    out = np.zeros((n, 8), dtype=np.uint64)
    _check(lib.bbg_read_g1_elements_from_buffer(out.ctypes.data, raw.ctypes.data, n * 64))
    return out


def ntt(coeffs, kind, generator_size=0, constant=None, n=None, stream=None):
    """In place.  numpy array -> host entry point (returns the same array); torch CUDA tensor -> _dev."""
    k = None if constant is None else _np(constant, 4)
    kp = None if k is None else k.ctypes.data
    if isinstance(coeffs, np.ndarray):
        assert coeffs.dtype == np.uint64 and coeffs.flags["C_CONTIGUOUS"]
        size = coeffs.size // 4 if n is None else n
        _check(lib.bbg_ntt(coeffs.ctypes.data, size, kind, generator_size, kp))
        return coeffs
    size = coeffs.numel() * coeffs.element_size() // 32 if n is None else n
    _check(lib.bbg_ntt_dev(coeffs.data_ptr(), size, kind, generator_size, kp, _stream_ptr(stream)))
    return coeffs


def fft(c, **kw):
    return ntt(c, FFT, **kw)


def ifft(c, **kw):
    return ntt(c, IFFT, **kw)


def coset_fft(c, generator_size=0, **kw):
    return ntt(c, COSET_FFT, generator_size=generator_size, **kw)


def coset_ifft(c, **kw):
    return ntt(c, COSET_IFFT, **kw)


def fft_with_constant(c, value, **kw):
    return ntt(c, FFT_WITH_CONSTANT, constant=value, **kw)


def ifft_with_constant(c, value, **kw):
    return ntt(c, IFFT_WITH_CONSTANT, constant=value, **kw)


def coset_fft_with_constant(c, value, generator_size=0, **kw):
    return ntt(c, COSET_FFT_WITH_CONSTANT, generator_size=generator_size, constant=value, **kw)


def coset_fft_with_generator_shift(c, value, generator_size=0, **kw):
    return ntt(c, COSET_FFT_WITH_GENERATOR_SHIFT, generator_size=generator_size, constant=value, **kw)


def coset_fft_ext(coeffs, n, domain_extension, stream=None):
    """coset_fft(coeffs, small_domain, large_domain, ext): coeffs holds ext*n elements, first n = input."""
    if isinstance(coeffs, np.ndarray):
        assert coeffs.dtype == np.uint64 and coeffs.flags["C_CONTIGUOUS"] and coeffs.size >= 4 * n * domain_extension, \
            "coset_fft_ext writes ext*n elements into coeffs"
        _check(lib.bbg_coset_fft_ext(coeffs.ctypes.data, n, domain_extension))
    else:
        _check(lib.bbg_coset_fft_ext_dev(coeffs.data_ptr(), n, domain_extension, _stream_ptr(stream)))
    return coeffs


def ntt_dist_layout(n, world):
    """(in_pos, out_pos): rank r holds input indices whose bits [in_pos, in_pos + log2 world) equal r (packed), and
    ends with the output indices whose bits [out_pos, ...) equal r."""
    a, b = ctypes.c_uint(0), ctypes.c_uint(0)
    _check(lib.bbg_ntt_dist_layout(n, world, ctypes.cast(ctypes.pointer(a), _vp), ctypes.cast(ctypes.pointer(b), _vp)))
    return a.value, b.value


def ntt_dist_phase(src, dst, n, kind, rank, world, phase, generator_size=0, constant=None, stream=None):
    """One phase of the multi-GPU NTT on torch CUDA tensors (include/bbg.h bbg_ntt_dist_dev)."""
    k = None if constant is None else _np(constant, 4)
    _check(lib.bbg_ntt_dist_dev(src.data_ptr(), dst.data_ptr(), n, kind, generator_size, None if k is None else k.ctypes.data,
                                rank, world, phase, _stream_ptr(stream)))
    return dst


def peer_buffer_alloc(nbytes):
    """(device pointer, 64-byte IPC handle) of a buffer other ranks can map (bbg_peer_buffer_alloc)"""
    ptr = ctypes.c_void_p(0)
    handle = (ctypes.c_uint8 * 64)()
    _check(lib.bbg_peer_buffer_alloc(nbytes, ctypes.cast(ctypes.pointer(ptr), _vp), ctypes.cast(handle, _vp)))
    return int(ptr.value), bytes(handle)


def peer_buffer_open(handle):
    ptr = ctypes.c_void_p(0)
    buf = (ctypes.c_uint8 * 64).from_buffer_copy(handle)
    _check(lib.bbg_peer_buffer_open(ctypes.cast(buf, _vp), ctypes.cast(ctypes.pointer(ptr), _vp)))
    return int(ptr.value)


def peer_buffer_close(ptr):
    _check(lib.bbg_peer_buffer_close(ptr))


def peer_buffer_free(ptr):
    _check(lib.bbg_peer_buffer_free(ptr))


def ntt_dist_fused_phase0(src, work, peer_ptrs, n, kind, rank, world, generator_size=0, constant=None, stream=None):
    """phase 0 with the exchange fused into its last pass (bbg_ntt_dist_fused_dev); peer_ptrs: `world` raw device pointers"""
    k = None if constant is None else _np(constant, 4)
    tab = (ctypes.c_void_p * world)(*peer_ptrs)
    _check(lib.bbg_ntt_dist_fused_dev(src.data_ptr(), work.data_ptr(), ctypes.cast(tab, _vp), n, kind, generator_size,
                                      None if k is None else k.ctypes.data, rank, world, _stream_ptr(stream)))


def ntt_dist_natural_phase(peer_in, work, peer_recv, peer_out, n, kind, rank, world, phase, generator_size=0, constant=None, stream=None):
    """bbg_ntt_dist_natural_dev: natural blocks in and out over peer memory; the tables are lists of raw device pointers
    (None where the phase does not use them), `work` a torch tensor or None"""
    k = None if constant is None else _np(constant, 4)

    def tab(ptrs):
        return None if ptrs is None else ctypes.cast((ctypes.c_void_p * world)(*ptrs), _vp)
    tabs = [tab(peer_in), tab(peer_recv), tab(peer_out)]
    _check(lib.bbg_ntt_dist_natural_dev(tabs[0], None if work is None else work.data_ptr(), tabs[1], tabs[2], n, kind, generator_size,
                                        None if k is None else k.ctypes.data, rank, world, phase, _stream_ptr(stream)))


class RawCudaArray:
    """a (rows, 4) int64 view of raw device memory for torch.as_tensor (the __cuda_array_interface__ protocol)"""

    def __init__(self, ptr, rows):
        self.__cuda_array_interface__ = {"shape": (rows, 4), "typestr": "<i8", "data": (int(ptr), False), "version": 3, "strides": None}


def ntt_dist_phase_raw(src_ptr, dst, n, kind, rank, world, phase, generator_size=0, constant=None, stream=None):
    """bbg_ntt_dist_dev with a raw device pointer as the source (a peer-mapped receive buffer)"""
    k = None if constant is None else _np(constant, 4)
    _check(lib.bbg_ntt_dist_dev(src_ptr, dst.data_ptr(), n, kind, generator_size, None if k is None else k.ctypes.data,
                                rank, world, phase, _stream_ptr(stream)))
    return dst


def domain_constants(n):
    out = np.zeros((6, 4), dtype=np.uint64)
    _check(lib.bbg_domain_constants(n, out.ctypes.data))
    return out


def field_op(field, op, a, b=None):
    a = _np(a, 4)
    bb = None if b is None else _np(b, 4)
    out = np.zeros_like(a)
    _check(lib.bbg_field_op(field, op, a.ctypes.data, None if bb is None else bb.ctypes.data, out.ctypes.data, a.shape[0]))
    return out


# ---- quotient-stage pointwise kernels and scans (include/bbg.h); numpy (host-pointer) flavour
KEEP_ON_DEVICE, KEEP_IF_AHEAD = 1, 2
NUM_POLYNOMIALS = 36
WIDGET_TURBO_ARITHMETIC, WIDGET_TURBO_FIXED_BASE, WIDGET_TURBO_RANGE, WIDGET_TURBO_LOGIC = range(4)


def _ptr_table(arrays, count):
    """ctypes array of `count` host pointers; `arrays`: dict index -> numpy array, or a sequence"""
    items = arrays.items() if isinstance(arrays, dict) else enumerate(arrays)
    tab = [None] * count
    for k, a in items:
        tab[k] = a.ctypes.data
    return (ctypes.c_void_p * count)(*tab)


def _fr1(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64).reshape(4))


def turbo_quotient(kind, polys, n_large, alpha_base, alpha, quotient, flags=0):
    """quotient (n_large, 4) += the gate identity of widget `kind`; polys: dict PolynomialIndex -> (n_large, 4) uint64"""
    a0, a = _fr1(alpha_base), _fr1(alpha)
    _check(lib.bbg_turbo_quotient(kind, ctypes.cast(_ptr_table(polys, NUM_POLYNOMIALS), _vp), n_large, a0.ctypes.data, a.ctypes.data,
                                  quotient.ctypes.data, flags))
    return quotient


def permutation_quotient(wire_ffts, sigma_ffts, z_fft, lagrange_1, n_large, roots_cut, alpha_base, beta, gamma, delta, quotient, flags=0):
    w = len(wire_ffts)
    c = [_fr1(x) for x in (alpha_base, beta, gamma, delta)]
    _check(lib.bbg_permutation_quotient(ctypes.cast(_ptr_table(wire_ffts, 4), _vp), ctypes.cast(_ptr_table(sigma_ffts, 4), _vp), w,
                                        z_fft.ctypes.data, lagrange_1.ctypes.data, n_large, roots_cut, c[0].ctypes.data, c[1].ctypes.data,
                                        c[2].ctypes.data, c[3].ctypes.data, quotient.ctypes.data, flags))
    return quotient


def divide_by_pseudo_vanishing_polynomial(evals, n_small, roots_cut=4, flags=0):
    _check(lib.bbg_divide_by_pseudo_vanishing_polynomial(evals.ctypes.data, n_small, evals.size // 4, roots_cut, flags))
    return evals


def compute_lagrange_polynomial_fft(n_small, n_large):
    out = np.zeros((n_large, 4), dtype=np.uint64)
    _check(lib.bbg_compute_lagrange_polynomial_fft(out.ctypes.data, n_small, n_large))
    return out


def permutation_grand_product(wires, sigmas, n, beta, gamma, z=None, flags=0):
    z = np.zeros((n, 4), dtype=np.uint64) if z is None else z
    b, g = _fr1(beta), _fr1(gamma)
    _check(lib.bbg_permutation_grand_product(ctypes.cast(_ptr_table(wires, 4), _vp), ctypes.cast(_ptr_table(sigmas, 4), _vp), len(wires), n,
                                             b.ctypes.data, g.ctypes.data, z.ctypes.data, flags))
    return z


def evaluate(coeffs, z, n=None):
    c = _np(coeffs, 4)
    out = np.zeros(4, dtype=np.uint64)
    zz = _fr1(z)
    _check(lib.bbg_evaluate(c.ctypes.data, c.shape[0] if n is None else n, zz.ctypes.data, out.ctypes.data))
    return out


def evaluate_batch(polys, zs):
    """[sum_i polys[k][i] * zs[k]^i] as a (k, 4) array, one launch"""
    arrs = [_np(a, 4) for a in polys]
    k = len(arrs)
    tab = (ctypes.c_void_p * max(k, 1))(*[a.ctypes.data for a in arrs])
    ns = (ctypes.c_size_t * max(k, 1))(*[a.shape[0] for a in arrs])
    z = _np(zs, 4)
    out = np.zeros((k, 4), dtype=np.uint64)
    _check(lib.bbg_evaluate_batch(ctypes.cast(tab, _vp), ctypes.cast(ns, _vp), k, z.ctypes.data, out.ctypes.data))
    return out


def compute_opening_polynomial(src, z, n_eval=None, n=None, dest=None, flags=0):
    """(dest, F(z)); dest may be src itself (in place)"""
    s = src if isinstance(src, np.ndarray) and src.dtype == np.uint64 and src.flags["C_CONTIGUOUS"] else _np(src, 4)
    n_eval = s.size // 4 if n_eval is None else n_eval
    n = n_eval if n is None else n
    dest = np.zeros((n, 4), dtype=np.uint64) if dest is None else dest
    f = np.zeros(4, dtype=np.uint64)
    zz = _fr1(z)
    _check(lib.bbg_compute_opening_polynomial(s.ctypes.data, dest.ctypes.data, zz.ctypes.data, n_eval, n, f.ctypes.data, flags))
    return dest, f


def linear_combination(polys, scalars, n, base=None, dest=None, flags=0):
    """dest[i] = (base[i] or 0) + sum_k polys[k][i] * scalars[k]"""
    dest = np.zeros((n, 4), dtype=np.uint64) if dest is None else dest
    sc = _np(scalars, 4)
    tab = (ctypes.c_void_p * max(len(polys), 1))(*[a.ctypes.data for a in polys])
    _check(lib.bbg_linear_combination(dest.ctypes.data, None if base is None else base.ctypes.data, ctypes.cast(tab, _vp), sc.ctypes.data,
                                      len(polys), n, flags))
    return dest


def wire_coset_fft(wire, wire_fft, n, ext=4, flags=0):
    _check(lib.bbg_wire_coset_fft(wire.ctypes.data, wire_fft.ctypes.data, n, ext, flags))
    return wire_fft


def wire_ifft(wire, lagrange_copy=None, flags=0):
    _check(lib.bbg_wire_ifft(wire.ctypes.data, wire.size // 4, None if lagrange_copy is None else lagrange_copy.ctypes.data, flags))
    return wire


def wire_ifft_batch(wires, lagrange_copies=None, flags=0):
    """bbg_wire_ifft_batch: in-place iffts of several equally long columns (numpy arrays)"""
    k = len(wires)
    w = (ctypes.c_void_p * k)(*[a.ctypes.data for a in wires])
    lc = None
    if lagrange_copies is not None:
        lc = (ctypes.c_void_p * k)(*[None if a is None else a.ctypes.data for a in lagrange_copies])
    _check(lib.bbg_wire_ifft_batch(ctypes.cast(w, _vp), wires[0].size // 4, None if lc is None else ctypes.cast(lc, _vp), k, flags))
    return wires


def poly_write(host_array, elem_offset, values):
    v = _np(values, 4)
    _check(lib.bbg_poly_write(host_array.ctypes.data, elem_offset, v.ctypes.data, v.shape[0]))


def ntt_ex(coeffs, kind, generator_size=0, constant=None, flags=0):
    k = None if constant is None else _np(constant, 4)
    _check(lib.bbg_ntt_ex(coeffs.ctypes.data, coeffs.size // 4, kind, generator_size, None if k is None else k.ctypes.data, flags))
    return coeffs


def field_op_dev(field, op, a, b=None, out=None, stream=None):
    """bbg_field_op on torch CUDA tensors of 32-byte elements (op 7 = reduce_once: canonical form for comparisons)."""
    import torch
    out = torch.empty_like(a) if out is None else out
    n = a.numel() * a.element_size() // 32
    _check(lib.bbg_field_op_dev(field, op, a.data_ptr(), None if b is None else b.data_ptr(), out.data_ptr(), n, _stream_ptr(stream)))
    return out


def g1_normalize(elements):
    """g1::affine_element(element) for (n, 12) Jacobian elements -> (n, 8) canonical affine."""
    e = _np(elements, 12)
    out = np.zeros((e.shape[0], 8), dtype=np.uint64)
    _check(lib.bbg_g1_normalize(e.ctypes.data, e.shape[0], out.ctypes.data))
    return out


def g1_op(op, a, b=None):
    a = _np(a, 12)
    bb = None if b is None else _np(b, 8 if op == 0 else 12)
    out = np.zeros_like(a)
    _check(lib.bbg_g1_op(op, a.ctypes.data, None if bb is None else bb.ctypes.data, out.ctypes.data, a.shape[0]))
    return out
