"""GPU tests for the round-2 entry points: batched MSMs (two streams / two workspaces), resident polynomials,
tiny and unregistered MSMs, cross-stream ordering of the `_dev` entry points and handle lifetime.

Same bar as tests/test_gpu_parity.py: bit-exact on canonical encodings against the oracle, through the C-ABI.
"""
import numpy as np
import pytest

import inputs
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bbg():
    import bbg as _bbg
    _bbg.init(0)
    return _bbg


def canon(orc, a):
    return np.array(orc.reduce(po.FR, np.asarray(a).reshape(-1, 4)))


# ------------------------------------------------------------------------------------------ batched MSM
# bb/plonk/proof_system/prover/work_queue.hpp:213-243 processes the prover's W_1..W_4 / T_1..T_4 commitments one by
# one; bbg_pippenger_unsafe_batch takes them together.  Every result must equal the single-call result.
@pytest.mark.parametrize("k", [1, 2, 3, 5])
def test_msm_batch_matches_single_calls(bbg, orc, srs_mini, k):
    pts, _ = srs_mini
    n = inputs.SRS_MINI_POINTS
    pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)
    arrays = [inputs.fr_elements(500 + i, n, coarse_fraction=0.2 if i % 2 else 0.0) for i in range(k)]
    got = pip.pippenger_unsafe_batch(arrays, 0, n)
    assert got.shape == (k, 12)
    for i in range(k):
        exp = orc.jac_to_buffer(orc.pippenger(arrays[i], pts, stride=1))
        assert orc.jac_to_buffer(got[i]) == exp, i
        assert orc.jac_to_buffer(pip.pippenger_unsafe(arrays[i], 0, n)) == exp, i


def test_msm_batch_subrange_and_device_pointers(bbg, orc, srs_mini):
    import torch
    pts, _ = srs_mini
    n = inputs.SRS_MINI_POINTS
    pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)
    frm, rng = 100, 3001  # Pippenger::pippenger_unsafe(scalars, from, range), ragged
    arrays = [inputs.fr_elements(600 + i, rng) for i in range(4)]
    exp = [orc.jac_to_buffer(orc.pippenger(a, pts[frm:frm + rng], stride=1)) for a in arrays]
    got = pip.pippenger_unsafe_batch(arrays, frm, rng)
    assert [orc.jac_to_buffer(g) for g in got] == exp
    dev = [torch.from_numpy(a.view(np.int64)).cuda() for a in arrays]
    out = pip.pippenger_unsafe_batch(dev, frm, rng)
    torch.cuda.synchronize()
    host = out.cpu().numpy().view(np.uint64).reshape(4, 12)
    assert [orc.jac_to_buffer(g) for g in host] == exp


def test_msm_batch_by_table_address(bbg, orc, srs_mini):
    pts, table = srs_mini
    n = 2048
    tab = np.ascontiguousarray(table[: 2 * n])
    pip = bbg.Pippenger.from_table(tab, n)
    arrays = [inputs.fr_elements(700 + i, n) for i in range(3)]
    got = bbg.pippenger_batch(arrays, tab, n)
    for i in range(3):
        assert orc.jac_to_buffer(got[i]) == orc.jac_to_buffer(orc.pippenger(arrays[i], pts[:n], stride=1))
    # a table the library has never seen is refused (the batch entry point never uploads bases)
    other = np.ascontiguousarray(table[: 2 * n]).copy()
    with pytest.raises(bbg.BbgError):
        bbg.pippenger_batch(arrays, other, n)
    pip.close()


def test_msm_batch_skewed_scalars_both_workspaces(bbg, orc, srs_mini):
    """all-equal scalars put every digit of a window in one bucket: the slot merge runs all of its levels (the
    cooperative k_msm_merge_rest loop), in both workspaces of a batch"""
    pts, _ = srs_mini
    n = inputs.SRS_MINI_POINTS
    pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)
    one = inputs.fr_elements(42, 1)
    same_a = np.repeat(one, n, axis=0)
    same_b = np.repeat(inputs.fr_elements(43, 1), n, axis=0)
    got = pip.pippenger_unsafe_batch([same_a, same_b, same_a], 0, n)
    ea = orc.jac_to_buffer(orc.pippenger(same_a, pts, stride=1))
    eb = orc.jac_to_buffer(orc.pippenger(same_b, pts, stride=1))
    assert [orc.jac_to_buffer(g) for g in got] == [ea, eb, ea]


@pytest.mark.parametrize("levels,c,k", [(0, 0, 9), (1, 9, 4), (2, 0, 6), (0, 12, 2)])
def test_msm_fused_batch_groups_and_level_splits(bbg, orc, srs_mini, levels, c, k, monkeypatch):
    """Small MSMs of one batch call are fused four at a time into ONE pass of the kernels (api.cu msm_batch, msm_device with
    an MsmBatch): vector m's digits go to bucket sets [m S, (m + 1) S).  Every way of splitting windows into fixed-base
    levels x bucket sets (S = 1, S > 1, no precomputed levels at all), full and partial groups, zero and one-point ranges,
    and a vector that is all zeros next to live ones must give the single-call results."""
    pts, _ = srs_mini
    if levels:
        monkeypatch.setenv("BBG_MSM_LEVELS", str(levels))
    if c:
        monkeypatch.setenv("BBG_MSM_C", str(c))
    pip = bbg.Pippenger.from_points(pts)
    n = 2500
    arrays = [inputs.fr_elements(900 + i, n, coarse_fraction=0.3 if i % 3 == 0 else 0.0) for i in range(k)]
    arrays[1] = np.zeros((n, 4), dtype=np.uint64)  # all-zero scalars: the point at infinity
    got = pip.pippenger_unsafe_batch(arrays, 7, n)
    for i in range(k):
        assert orc.jac_to_buffer(got[i]) == orc.jac_to_buffer(orc.pippenger(arrays[i], pts[7:7 + n], stride=1)), i
    one = pip.pippenger_unsafe_batch([a[:1] for a in arrays], 3, 1)
    for i in range(k):
        assert orc.jac_to_buffer(one[i]) == orc.jac_to_buffer(orc.pippenger(arrays[i][:1], pts[3:4], stride=1)), i
    none = pip.pippenger_unsafe_batch([a[:0] for a in arrays], 0, 0)
    inf = orc.jac_to_buffer(orc.pippenger(arrays[0][:0], pts[:0], stride=1))
    assert all(orc.jac_to_buffer(g) == inf for g in none)


def test_msm_fused_batch_can_be_disabled(bbg, orc, srs_mini):
    """BBG_MSM_FUSED_BATCH_MAX_LOG2 is read once per process: the un-fused batch path (one chain per MSM on four streams,
    what large MSMs use) is exercised in a child process on the same inputs."""
    import os
    import subprocess
    import sys
    code = (
        "import sys\n"
        "sys.path[:0] = ['.', 'tests', 'aztec-2.0_b200/python']\n"
        "import numpy as np, bbg, inputs\n"
        "from oracle import pyoracle as po\n"
        "bbg.init(0)\n"
        "orc = po.Oracle()\n"
        "n = inputs.SRS_MINI_POINTS\n"
        "pts = orc.read_transcript_g1(n, inputs.SRS_MINI_DIR)\n"
        "pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)\n"
        "arrays = [inputs.fr_elements(950 + i, n) for i in range(6)]\n"
        "got = pip.pippenger_unsafe_batch(arrays, 0, n)\n"
        "ok = all(orc.jac_to_buffer(got[i]) == orc.jac_to_buffer(orc.pippenger(arrays[i], pts, stride=1)) for i in range(6))\n"
        "print('OK' if ok else 'MISMATCH')\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BBG_MSM_FUSED_BATCH_MAX_LOG2="0")
    r = subprocess.run([sys.executable, "-c", code.replace("\\n", "\n")], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-500:] + r.stderr[-2000:]


# ------------------------------------------------------------------------------------------ tiny / unregistered MSMs
# bb/plonk/proof_system/verifier/verifier.cpp:164-170 calls pippenger with a few dozen points
@pytest.mark.parametrize("n", [1, 2, 7, 27, 33, 255, 256, 257])
def test_tiny_msm_unregistered_points(bbg, orc, srs_mini, n):
    pts, table = srs_mini
    sc = inputs.fr_elements(800 + n, n, coarse_fraction=0.25)
    exp = orc.jac_to_buffer(orc.pippenger(sc, pts[:n], stride=1))
    assert orc.jac_to_buffer(bbg.msm_points(sc, pts[:n])) == exp
    assert orc.jac_to_buffer(bbg.pippenger(sc, np.ascontiguousarray(table[: 2 * n]), n)) == exp


def test_tiny_msm_edge_cases(bbg, orc, srs_mini):
    pts, _ = srs_mini
    n = 27
    # zero scalars, repeated points, opposite points
    z = np.zeros((n, 4), dtype=np.uint64)
    assert orc.jac_to_buffer(bbg.msm_points(z, pts[:n])) == orc.jac_to_buffer(orc.pippenger(z, pts[:n], stride=1))
    rep = np.repeat(pts[5:6], n, axis=0)
    sc = inputs.fr_elements(901, n)
    assert orc.jac_to_buffer(bbg.msm_points(sc, rep)) == orc.jac_to_buffer(orc.pippenger(sc, rep, stride=1))


# ------------------------------------------------------------------------------------------ resident polynomials
def test_resident_chain_ifft_msm_fft(bbg, orc, srs_mini):
    """ifft -> commitment MSM -> coset FFT on the same host array (what the prover does to every wire): with residency
    on the second and third call find the device mirror; results are identical to the non-resident path."""
    pts, _ = srs_mini
    n = inputs.SRS_MINI_POINTS
    pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)
    x0 = inputs.fr_elements(910, n)
    # plain path
    a = x0.copy()
    bbg.ifft(a)
    msm_plain = orc.jac_to_buffer(pip.pippenger_unsafe(a, 0, n))
    fft_plain = bbg.coset_fft(a.copy())
    bbg.resident_mode(True)
    try:
        s0 = bbg.resident_stats()
        b = x0.copy()
        bbg.ifft(b)
        assert np.array_equal(canon(orc, b), canon(orc, a))
        assert orc.jac_to_buffer(pip.pippenger_unsafe(b, 0, n)) == msm_plain
        s1 = bbg.resident_stats()
        assert s1["hits"] >= s0["hits"] + 1 and s1["h2d_bytes_saved"] >= s0["h2d_bytes_saved"] + 32 * n
        c = b  # transformed in place again from the mirror
        bbg.coset_fft(c)
        assert np.array_equal(canon(orc, c), canon(orc, fft_plain))
        assert bbg.resident_stats()["hits"] >= s1["hits"] + 1
    finally:
        bbg.resident_mode(False)
    assert bbg.resident_stats()["bytes_resident"] == 0


def test_resident_detects_host_rewrite(bbg, orc, srs_mini):
    """the caller rewrites its array between two calls without telling the library: the fingerprint check must catch it
    (whole-array rewrite, the prover's blinding positions at the end, and the very first element)"""
    pts, _ = srs_mini
    n = inputs.SRS_MINI_POINTS
    pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)
    bbg.resident_mode(True)
    try:
        a = inputs.fr_elements(920, n)
        first = orc.jac_to_buffer(pip.pippenger_unsafe(a, 0, n))
        assert first == orc.jac_to_buffer(orc.pippenger(a, pts, stride=1))
        for how in ("all", "tail", "head"):
            if how == "all":
                a[...] = inputs.fr_elements(921, n)
            elif how == "tail":
                a[n - 3:] = inputs.fr_elements(922, 3)  # prover.cpp:181-183 writes its blinding scalars there
            else:
                a[0] = inputs.fr_elements(923, 1)[0]
            got = orc.jac_to_buffer(pip.pippenger_unsafe(a, 0, n))
            assert got == orc.jac_to_buffer(orc.pippenger(a, pts, stride=1)), how
    finally:
        bbg.resident_mode(False)


def test_resident_slice_of_array(bbg, orc, srs_mini):
    """the quotient polynomial is committed to in four slices of one 4n array (prover.cpp:84-135)"""
    pts, _ = srs_mini
    n = 1024
    pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)
    bbg.resident_mode(True)
    try:
        q = inputs.fr_elements(930, 4 * n)
        bbg.coset_ifft(q)  # mirrors all 4n elements
        s0 = bbg.resident_stats()
        for i in range(4):
            sl = q[i * n:(i + 1) * n]
            assert orc.jac_to_buffer(pip.pippenger_unsafe(sl, 0, n)) == orc.jac_to_buffer(orc.pippenger(np.ascontiguousarray(sl), pts[:n], stride=1))
        assert bbg.resident_stats()["hits"] == s0["hits"] + 4
    finally:
        bbg.resident_mode(False)


# ------------------------------------------------------------------------------------------ streams / lifetime
def test_dev_calls_on_different_streams_are_ordered(bbg, orc, srs_mini):
    """`_dev` calls on two streams share the workspaces; the library must order them"""
    import torch
    pts, _ = srs_mini
    n = inputs.SRS_MINI_POINTS
    pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)
    a = inputs.fr_elements(940, n)
    b = inputs.fr_elements(941, n)
    ea = orc.jac_to_buffer(orc.pippenger(a, pts, stride=1))
    eb = orc.jac_to_buffer(orc.pippenger(b, pts, stride=1))
    da, db = torch.from_numpy(a.view(np.int64)).cuda(), torch.from_numpy(b.view(np.int64)).cuda()
    x = torch.from_numpy(inputs.fr_elements(942, 1 << 14).view(np.int64)).cuda()
    x_ref = x.clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        ra = pip.pippenger_unsafe(da, 0, n, stream=s1)
        rb = pip.pippenger_unsafe(db, 0, n, stream=s2)
        bbg.ntt(x, bbg.FFT, stream=s1)   # first use of this size on s1 builds the twiddle table ...
        bbg.ntt(x, bbg.IFFT, stream=s2)  # ... which s2 reads straight away
        host = pip.pippenger_unsafe(a, 0, n)  # host-pointer call on the library's own stream in between
        torch.cuda.synchronize()
        assert orc.jac_to_buffer(ra.cpu().numpy().view(np.uint64)) == ea
        assert orc.jac_to_buffer(rb.cpu().numpy().view(np.uint64)) == eb
        assert orc.jac_to_buffer(host) == ea
        assert np.array_equal(canon(orc, x.cpu().numpy().view(np.uint64)), canon(orc, x_ref.cpu().numpy().view(np.uint64)))


def test_stale_pippenger_handle_after_delete_is_harmless(bbg):
    pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, 64)
    h = pip.h
    pip.close()
    bbg.lib.bbg_delete_pippenger(h)  # double delete of a stale handle: looked up, not found, ignored
    assert pip.h is None


def test_init_on_another_device_is_refused(bbg):
    if bbg.device_count() < 2:
        assert bbg.lib.bbg_init(0) == bbg.OK  # same device: idempotent
        return
    assert bbg.lib.bbg_init(1) == bbg.ERR_ARG


def test_profile_reports_every_phase(bbg):
    """13 phases (ntt_pass3 was dropped by a 12-entry read)"""
    assert bbg.NUM_PHASES == len(bbg.PHASE_NAMES) == 13
    import torch
    x = torch.from_numpy(inputs.fr_elements(950, 1 << 13).view(np.int64)).cuda()
    bbg.profile(True)
    bbg.ntt(x, bbg.FFT)
    ph = bbg.profile_read()
    bbg.profile(False)
    assert set(ph) == set(bbg.PHASE_NAMES) and ph["ntt_pass0"] > 0 and ph["ntt_pass1"] > 0


# ------------------------------------------------------------------------------------------ g1 normalise
# bb/ecc/groups/element_impl.hpp:51-68  g1::affine_element(element)
def test_g1_normalize_vs_oracle(bbg, orc, srs_mini):
    pts, _ = srs_mini
    jacs = []
    acc = orc.g1_infinity()
    for i in range(40):
        acc = orc.g1_mixed_add(acc, pts[i])
        jacs.append(np.array(acc).reshape(12))
    jacs.append(np.array(orc.g1_infinity()).reshape(12))
    jacs = np.stack(jacs)
    got = bbg.g1_normalize(jacs)
    for i in range(jacs.shape[0]):
        assert orc.affine_to_buffer(got[i]) == orc.jac_to_buffer(jacs[i]), i
    assert got[-1, 3] >> 63 == 1  # infinity keeps barretenberg's flag (bit 255 of x)


def test_field_op_dev_reduce(bbg, orc):
    import torch
    a = inputs.fr_elements(960, 4096, coarse_fraction=0.5)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    out = bbg.field_op_dev(po.FR, 7, d)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), orc.reduce(po.FR, a))


def test_stats_totals_count_pcie_bytes(bbg):
    """bbg_stats_totals: what the prover harness differences to report PCIe bytes per proof"""
    x = inputs.fr_elements(970, 1 << 12)
    t0 = bbg.stats_totals()
    bbg.fft(x)
    t1 = bbg.stats_totals()
    assert t1["ntt"]["calls"] == t0["ntt"]["calls"] + 1
    assert t1["ntt"]["h2d"] - t0["ntt"]["h2d"] == x.nbytes and t1["ntt"]["d2h"] - t0["ntt"]["d2h"] == x.nbytes
    bbg.resident_mode(True)
    try:
        bbg.ifft(x)  # uploads once, mirror kept
        t2 = bbg.stats_totals()
        bbg.fft(x)   # mirror hit: nothing goes up
        t3 = bbg.stats_totals()
        assert t3["ntt"]["h2d"] == t2["ntt"]["h2d"] and t3["ntt"]["d2h"] - t2["ntt"]["d2h"] == x.nbytes
    finally:
        bbg.resident_mode(False)


def test_ntt_persistent_prefetch_variant_matches_default(bbg):
    """k_ntt_pass<.., PERSIST> (persistent CTAs, cp.async prefetch of the next tile group; BBG_NTT_PERSIST=1, an
    experiment that measured slower and is off by default): the knob is read once per process, so the variant runs in a
    child process and its fft / coset_fft / coset_ifft of a 2^20 array must equal this process's default kernels bit for bit."""
    import hashlib
    import os
    import subprocess
    import sys
    n = 1 << 20
    x = inputs.fr_elements(4242, n)
    here = {}
    for kind in (bbg.FFT, bbg.COSET_FFT, bbg.COSET_IFFT):
        here[kind] = hashlib.sha256(bbg.ntt(x.copy(), kind).tobytes()).hexdigest()
    code = (
        "import sys, hashlib\n"
        "sys.path[:0] = ['.', 'tests', 'aztec-2.0_b200/python']\n"
        "import bbg, inputs\n"
        "bbg.init(0)\n"
        "x = inputs.fr_elements(4242, 1 << 20)\n"
        "for kind in (bbg.FFT, bbg.COSET_FFT, bbg.COSET_IFFT):\n"
        "    print(kind, hashlib.sha256(bbg.ntt(x.copy(), kind).tobytes()).hexdigest())\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BBG_NTT_PERSIST="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = dict((int(l.split()[0]), l.split()[1]) for l in r.stdout.strip().splitlines())
    assert got == here
