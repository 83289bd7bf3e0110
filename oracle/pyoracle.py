"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings for the two checkers:

* ``Oracle``  -> oracle/libbn254_oracle.so   our plain-C restatement (bn254_oracle.c)
* ``Ref``     -> oracle/_ref/libbbref.so      the UNMODIFIED reference, compiled from
                                              /root/reference by oracle/Makefile (may be absent)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arms import
this module.  The product package (aztec-2.0_b200/) never does.

All field elements are numpy uint64 arrays of shape (..., 4) (little-endian limbs, Montgomery
form), affine points (..., 8), Jacobian points (..., 12) -- exactly the reference's memory layout.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libbn254_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libbbref.so")
REF_GPU_SO = os.path.join(HERE, "_ref", "libbbref_gpu.so")  # same reference objects, hot path -> bbg_shim.cpp -> libbbg.so
REF_SRS_DIR = os.path.join(HERE, "_ref", "srs_db")

FQ, FR = 0, 1
OP_MUL, OP_ADD, OP_SUB, OP_SQR, OP_TO_MONT, OP_FROM_MONT, OP_INVERT, OP_REDUCE, OP_NEG = range(9)
(NTT_FFT, NTT_IFFT, NTT_COSET_FFT, NTT_COSET_IFFT, NTT_FFT_CONST, NTT_IFFT_CONST,
 NTT_COSET_FFT_CONST, NTT_COSET_FFT_GEN_SHIFT) = range(8)

FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ_MODULUS = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


def aligned_empty(shape, dtype=np.uint64, align=64):
    """numpy array whose data pointer is `align`-byte aligned (the reference uses aligned AVX moves)."""
    shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    raw = np.zeros(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes].view(dtype).reshape(shape)


def aligned_copy(a, align=64):
    out = aligned_empty(a.shape, a.dtype, align)
    out[...] = a
    return out


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def int_to_limbs(v):
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def limbs_to_int(a):
    return sum(int(a[i]) << (64 * i) for i in range(4))


def random_field_ints(rng, n, modulus):
    """n uniform integers in [0, modulus) from an explicit numpy Generator (512 random bits mod p,
    like field::random_element, field_impl.hpp:505-515, but with our own reproducible stream)."""
    raw = rng.integers(0, 1 << 64, size=(n, 8), dtype=np.uint64)
    return [sum(int(raw[i, j]) << (64 * j) for j in range(8)) % modulus for i in range(n)]


def ints_to_array(vals):
    out = aligned_empty((len(vals), 4))
    for i, v in enumerate(vals):
        for j in range(4):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def fnv1a64(arr):
    """FNV-1a-64 over u64 words in order (SURVEY.md Appendix B convention)."""
    h = 0xCBF29CE484222325
    for v in np.ascontiguousarray(arr).reshape(-1).tolist():
        h ^= v
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


class Oracle:
    """Plain-C restatement (always available; built on demand)."""

    name = "port"

    def __init__(self):
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
                os.path.join(HERE, "bn254_oracle.c")):
            build_oracle()
        self.lib = ctypes.CDLL(ORACLE_SO)

    # -- fields
    def field_op(self, field, op, a, b=None):
        a = aligned_copy(np.asarray(a, dtype=np.uint64).reshape(-1, 4))
        bb = None if b is None else aligned_copy(np.asarray(b, dtype=np.uint64).reshape(-1, 4))
        out = aligned_empty(a.shape)
        for i in range(a.shape[0]):
            self.lib.orc_field_op(field, op, _p(a[i]), None if bb is None else _p(bb[i]), _p(out[i]))
        return out

    def to_mont(self, field, ints):
        return self.field_op(field, OP_TO_MONT, ints_to_array(ints))

    def from_mont_ints(self, field, arr):
        c = self.field_op(field, OP_FROM_MONT, arr)
        return [limbs_to_int(c[i]) for i in range(c.shape[0])]

    def reduce(self, field, arr):
        return self.field_op(field, OP_REDUCE, arr).reshape(np.asarray(arr).shape)

    def fr_root_of_unity(self, log2n):
        out = aligned_empty(4)
        self.lib.orc_fr_root_of_unity(ctypes.c_uint(log2n), _p(out))
        return out

    def domain_constants(self, n):
        out = aligned_empty((6, 4))
        self.lib.orc_domain_constants(ctypes.c_size_t(n), _p(out))
        return out

    # -- group
    def g1_one(self):
        out = aligned_empty(8)
        self.lib.orc_g1_one(_p(out))
        return out

    def g1_infinity(self):
        out = aligned_empty(12)
        self.lib.orc_g1_set_infinity(_p(out))
        return out

    def g1_mixed_add(self, jac, aff):
        jac, aff, out = aligned_copy(jac), aligned_copy(aff), aligned_empty(12)
        self.lib.orc_g1_mixed_add(_p(jac), _p(aff), _p(out))
        return out

    def g1_add(self, a, b):
        a, b, out = aligned_copy(a), aligned_copy(b), aligned_empty(12)
        self.lib.orc_g1_add(_p(a), _p(b), _p(out))
        return out

    def g1_dbl(self, a):
        a, out = aligned_copy(a), aligned_empty(12)
        self.lib.orc_g1_dbl(_p(a), _p(out))
        return out

    def g1_to_affine(self, jac):
        jac, out = aligned_copy(jac), aligned_empty(8)
        self.lib.orc_g1_to_affine(_p(jac), _p(out))
        return out

    def g1_mul(self, aff, scalar):
        aff, scalar, out = aligned_copy(aff), aligned_copy(scalar), aligned_empty(12)
        self.lib.orc_g1_mul(_p(aff), _p(scalar), _p(out))
        return out

    def g1_on_curve(self, aff):
        aff = aligned_copy(aff)
        return bool(self.lib.orc_g1_on_curve(_p(aff)))

    def g1_sum(self, jacs):
        jacs = aligned_copy(np.asarray(jacs, dtype=np.uint64).reshape(-1, 12))
        out = aligned_empty(12)
        self.lib.orc_g1_sum(_p(jacs), ctypes.c_size_t(jacs.shape[0]), _p(out))
        return out

    def affine_to_buffer(self, aff):
        aff = aligned_copy(aff)
        buf = np.zeros(64, dtype=np.uint8)
        self.lib.orc_g1_affine_to_buffer(_p(aff), _p(buf))
        return bytes(buf)

    def jac_to_buffer(self, jac):
        """The 64 bytes work_queue.hpp:231,239 feeds the transcript: the canonical MSM result."""
        jac = aligned_copy(np.asarray(jac, dtype=np.uint64).reshape(12))
        buf = np.zeros(64, dtype=np.uint8)
        self.lib.orc_g1_jac_to_buffer(_p(jac), _p(buf))
        return bytes(buf)

    # -- SRS
    def read_g1_elements_from_buffer(self, raw):
        raw = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw
        n = raw.size // 64
        out = aligned_empty((n, 8))
        self.lib.orc_read_g1_elements_from_buffer(_p(out), _p(np.ascontiguousarray(raw)), ctypes.c_size_t(n * 64))
        return out

    def read_transcript_g1(self, degree, srs_dir, slack=0):
        out = aligned_empty((degree + slack, 8))
        rc = self.lib.orc_read_transcript_g1(_p(out), ctypes.c_size_t(degree), srs_dir.encode())
        if rc != 0:
            raise RuntimeError("srs too short")
        return out

    def point_table(self, points):
        """2n-entry interleaved table [P0, phi(P0), P1, ...] (scalar_multiplication.cpp:104-112)."""
        n = points.shape[0]
        table = aligned_empty((2 * n + 256, 8))
        pts = aligned_copy(points)
        self.lib.orc_generate_pippenger_point_table(_p(pts), _p(table), ctypes.c_size_t(n))
        return table

    # -- MSM
    def pippenger(self, scalars, points, n=None, stride=2):
        scalars, points = aligned_copy(scalars), aligned_copy(points)
        n = scalars.shape[0] if n is None else n
        out = aligned_empty(12)
        self.lib.orc_pippenger(_p(scalars), _p(points), ctypes.c_size_t(n), ctypes.c_size_t(stride), _p(out))
        return out

    def naive_msm(self, scalars, points, n=None, stride=2):
        scalars, points = aligned_copy(scalars), aligned_copy(points)
        n = scalars.shape[0] if n is None else n
        out = aligned_empty(12)
        self.lib.orc_naive_msm(_p(scalars), _p(points), ctypes.c_size_t(n), ctypes.c_size_t(stride), _p(out))
        return out

    # -- NTT
    def ntt(self, kind, coeffs, generator_size=0, constant=None):
        c = aligned_copy(np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4))
        k = None if constant is None else aligned_copy(np.asarray(constant, dtype=np.uint64).reshape(4))
        self.lib.orc_ntt(kind, _p(c), ctypes.c_size_t(c.shape[0]), ctypes.c_size_t(generator_size), _p(k))
        return c

    def coset_fft_ext(self, coeffs, n, ext):
        c = aligned_empty((n * ext, 4))
        c[:n] = np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4)[:n]
        self.lib.orc_coset_fft_ext(_p(c), ctypes.c_size_t(n), ctypes.c_size_t(ext))
        return c

    def evaluate(self, coeffs, z):
        c = aligned_copy(np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4))
        z = aligned_copy(z)
        out = aligned_empty(4)
        self.lib.orc_evaluate(_p(c), _p(z), ctypes.c_size_t(c.shape[0]), _p(out))
        return out

    def set_num_threads(self, n):
        self.lib.orc_set_num_threads(int(n))


class Ref:
    """The unmodified reference (oracle/_ref/libbbref.so). `Ref.available()` says if it was built."""

    name = "reference"

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self, path=None):
        self.lib = L = ctypes.CDLL(path or REF_SO)
        L.ref_point_table_size.restype = ctypes.c_size_t
        L.ref_new_domain.restype = ctypes.c_void_p
        L.ref_new_runtime_state.restype = ctypes.c_void_p
        self._domains = {}

    def num_threads(self):
        return int(self.lib.ref_num_threads())

    def set_num_threads(self, n):
        self.lib.ref_set_num_threads(int(n))

    def field_op(self, field, op, a, b=None):
        a = aligned_copy(np.asarray(a, dtype=np.uint64).reshape(-1, 4))
        bb = None if b is None else aligned_copy(np.asarray(b, dtype=np.uint64).reshape(-1, 4))
        out = aligned_empty(a.shape)
        fn = self.lib.ref_fq_op if field == FQ else self.lib.ref_fr_op
        for i in range(a.shape[0]):
            fn(op, _p(a[i]), None if bb is None else _p(bb[i]), _p(out[i]))
        return out

    def reduce(self, field, arr):
        return self.field_op(field, OP_REDUCE, arr).reshape(np.asarray(arr).shape)

    def fr_constants(self):
        out = aligned_empty((3, 4))
        self.lib.ref_fr_constants(_p(out[0]), _p(out[1]), _p(out[2]))
        return out

    def fq_constants(self):
        out = aligned_empty((4, 4))
        self.lib.ref_fq_constants(_p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]))
        return out

    def debug_random_frs(self, n):
        out = aligned_empty((n, 4))
        self.lib.ref_debug_random_frs(_p(out), ctypes.c_size_t(n))
        return out

    def split_endo(self, k):
        k, out = aligned_copy(k), aligned_empty(4)
        self.lib.ref_fr_split_endo(_p(k), _p(out))
        return out

    def g1_infinity(self):
        out = aligned_empty(12)
        self.lib.ref_g1_set_infinity(_p(out))
        return out

    def g1_mixed_add(self, jac, aff):
        jac, aff, out = aligned_copy(jac), aligned_copy(aff), aligned_empty(12)
        self.lib.ref_g1_mixed_add(_p(jac), _p(aff), _p(out))
        return out

    def g1_add(self, a, b):
        a, b, out = aligned_copy(a), aligned_copy(b), aligned_empty(12)
        self.lib.ref_g1_add(_p(a), _p(b), _p(out))
        return out

    def g1_dbl(self, a):
        a, out = aligned_copy(a), aligned_empty(12)
        self.lib.ref_g1_dbl(_p(a), _p(out))
        return out

    def g1_to_affine(self, jac):
        jac, out = aligned_copy(jac), aligned_empty(8)
        self.lib.ref_g1_to_affine(_p(jac), _p(out))
        return out

    def g1_mul(self, aff, scalar):
        aff, scalar, out = aligned_copy(aff), aligned_copy(scalar), aligned_empty(12)
        self.lib.ref_g1_mul(_p(aff), _p(scalar), _p(out))
        return out

    def g1_on_curve(self, aff):
        aff = aligned_copy(aff)
        return bool(self.lib.ref_g1_on_curve(_p(aff)))

    def g1_hash_to_curve(self, seed):
        out = aligned_empty(8)
        self.lib.ref_g1_hash_to_curve(ctypes.c_uint64(seed), _p(out))
        return out

    def g1_sum(self, jacs):
        jacs = aligned_copy(np.asarray(jacs, dtype=np.uint64).reshape(-1, 12))
        out = aligned_empty(12)
        self.lib.ref_g1_sum(_p(jacs), ctypes.c_size_t(jacs.shape[0]), _p(out))
        return out

    def affine_to_buffer(self, aff):
        aff = aligned_copy(aff)
        buf = np.zeros(64, dtype=np.uint8)
        self.lib.ref_g1_affine_to_buffer(_p(aff), _p(buf))
        return bytes(buf)

    def jac_to_buffer(self, jac):
        jac = aligned_copy(np.asarray(jac, dtype=np.uint64).reshape(12))
        buf = np.zeros(64, dtype=np.uint8)
        self.lib.ref_g1_jac_to_buffer(_p(jac), _p(buf))
        return bytes(buf)

    def read_g1_elements_from_buffer(self, raw):
        raw = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw
        n = raw.size // 64
        out = aligned_empty((n, 8))
        self.lib.ref_read_g1_elements_from_buffer(_p(out), _p(np.ascontiguousarray(raw)), ctypes.c_size_t(n * 64))
        return out

    def read_transcript_g1(self, degree, srs_dir=REF_SRS_DIR, slack=0):
        out = aligned_empty((degree + slack, 8))
        rc = self.lib.ref_read_transcript_g1(_p(out), ctypes.c_size_t(degree), srs_dir.encode())
        if rc != 0:
            raise RuntimeError("srs too short")
        return out

    def point_table(self, points):
        n = points.shape[0]
        size = int(self.lib.ref_point_table_size(ctypes.c_size_t(n)))
        table = aligned_empty((max(size, 2 * n + 256), 8))
        table[:n] = points
        self.lib.ref_generate_pippenger_point_table(_p(table), _p(table), ctypes.c_size_t(n))
        return table

    def new_runtime_state(self, n):
        return ctypes.c_void_p(self.lib.ref_new_runtime_state(ctypes.c_size_t(n)))

    def delete_runtime_state(self, st):
        self.lib.ref_delete_runtime_state(st)

    def pippenger(self, scalars, table, n=None, unsafe=True, state=None, copy=True):
        """table MUST be a 2n interleaved table with slack (use point_table())."""
        if copy:
            scalars, table = aligned_copy(scalars), aligned_copy(table)
        n = scalars.shape[0] if n is None else n
        out = aligned_empty(12)
        rc = self.lib.ref_pippenger(_p(scalars), _p(table), ctypes.c_size_t(n), state, 1 if unsafe else 0, _p(out))
        if rc != 0:
            raise RuntimeError("reference pippenger threw")
        return out

    # Pippenger class (pippenger.hpp:35-52)
    def new_pippenger(self, srs_dir, num_points):
        self.lib.ref_new_pippenger_from_path.restype = ctypes.c_void_p
        h = self.lib.ref_new_pippenger_from_path(srs_dir.encode(), ctypes.c_size_t(num_points))
        if not h:
            raise RuntimeError("Pippenger constructor threw")
        return ctypes.c_void_p(h)

    def delete_pippenger(self, h):
        self.lib.ref_delete_pippenger(h)

    def pippenger_table(self, h, num_points):
        out = aligned_empty((2 * num_points, 8))
        self.lib.ref_pippenger_copy_table(h, _p(out))
        return out

    def pippenger_class_unsafe(self, h, scalars, from_, range_):
        scalars = aligned_copy(scalars)
        out = aligned_empty(12)
        rc = self.lib.ref_pippenger_class_unsafe(h, _p(scalars), ctypes.c_size_t(from_), ctypes.c_size_t(range_), _p(out))
        if rc != 0:
            raise RuntimeError("Pippenger::pippenger_unsafe threw")
        return out

    def naive_msm(self, scalars, points, n=None, stride=2):
        scalars, points = aligned_copy(scalars), aligned_copy(points)
        n = scalars.shape[0] if n is None else n
        out = aligned_empty(12)
        self.lib.ref_naive_msm(_p(scalars), _p(points), ctypes.c_size_t(n), ctypes.c_size_t(stride), _p(out))
        return out

    def domain(self, n, generator_size=0):
        key = (n, generator_size)
        if key not in self._domains:
            self._domains[key] = ctypes.c_void_p(
                self.lib.ref_new_domain(ctypes.c_size_t(n), ctypes.c_size_t(generator_size)))
        return self._domains[key]

    def domain_constants(self, n):
        out = aligned_empty((6, 4))
        self.lib.ref_domain_constants(self.domain(n), _p(out))
        return out

    def ntt(self, kind, coeffs, generator_size=0, constant=None, inplace=False):
        c = coeffs if inplace else aligned_copy(np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4))
        k = None if constant is None else aligned_copy(np.asarray(constant, dtype=np.uint64).reshape(4))
        self.lib.ref_ntt(self.domain(c.shape[0], generator_size), kind, _p(c), _p(k))
        return c

    def coset_fft_ext(self, coeffs, n, ext):
        c = aligned_empty((n * ext, 4))
        c[:n] = np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4)[:n]
        self.lib.ref_coset_fft_ext(self.domain(n), self.domain(n * ext), _p(c), ctypes.c_size_t(ext))
        return c

    def evaluate(self, coeffs, z):
        c = aligned_copy(np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4))
        z = aligned_copy(z)
        out = aligned_empty(4)
        self.lib.ref_evaluate(_p(c), _p(z), ctypes.c_size_t(c.shape[0]), _p(out))
        return out


def _ref_poly_methods():
    """quotient-stage functions of the compiled reference (oracle/ref_shim.cpp), attached to Ref below"""
    def divide_by_pseudo_vanishing_polynomial(self, evals, n_small, roots_cut=4):
        c = aligned_copy(np.asarray(evals, dtype=np.uint64).reshape(-1, 4))
        self.lib.ref_divide_by_pseudo_vanishing_polynomial(_p(c), self.domain(n_small), self.domain(c.shape[0]), ctypes.c_size_t(roots_cut))
        return c

    def compute_lagrange_polynomial_fft(self, n_small, n_large):
        c = aligned_empty((n_large, 4))
        self.lib.ref_compute_lagrange_polynomial_fft(_p(c), self.domain(n_small), self.domain(n_large))
        return c

    def compute_kate_opening_coefficients(self, src, z):
        s_ = aligned_copy(np.asarray(src, dtype=np.uint64).reshape(-1, 4))
        d = aligned_empty(s_.shape)
        f = aligned_empty(4)
        self.lib.ref_compute_kate_opening_coefficients(_p(s_), _p(d), _p(aligned_copy(z)), ctypes.c_size_t(s_.shape[0]), _p(f))
        return d, f

    def turbo_quotient(self, kind, polys, n_large, alpha_base, alpha, quotient):
        """polys: dict PolynomialIndex -> (n_large, 4) array; returns the accumulated quotient (reference templates)"""
        keep = {k: aligned_copy(np.asarray(v, dtype=np.uint64).reshape(-1, 4)) for k, v in polys.items()}
        table = (ctypes.c_void_p * 36)(*[keep[k].ctypes.data if k in keep else None for k in range(36)])
        q = aligned_copy(np.asarray(quotient, dtype=np.uint64).reshape(-1, 4))
        self.lib.ref_turbo_quotient(kind, table, ctypes.c_size_t(n_large), _p(aligned_copy(alpha_base)), _p(aligned_copy(alpha)), _p(q))
        return q
    return (divide_by_pseudo_vanishing_polynomial, compute_lagrange_polynomial_fft, compute_kate_opening_coefficients, turbo_quotient)


for _f in _ref_poly_methods():
    setattr(Ref, _f.__name__, _f)


def best_checker():
    """The reference itself when its .so travelled with the repo, else the plain-C restatement."""
    return Ref() if Ref.available() else Oracle()
