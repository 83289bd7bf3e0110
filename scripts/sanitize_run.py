"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck) on the GPU box:
    compute-sanitizer --tool memcheck python scripts/sanitize_run.py
Covers every kernel of the default paths once: SRS decode + fixed-base precompute, MSM through the Pippenger object
(c = 13 table) and through the per-call path (L = 1), skewed scalars (merge levels), the NTT passes for 2-, 3-pass sizes
with all scalings, twice (cached tables), and g1_sum."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "aztec-2.0_b200", "python")):
    sys.path.insert(0, p)
import bbg  # noqa: E402
import inputs  # noqa: E402

bbg.init(0)
n = inputs.SRS_MINI_POINTS
pip = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)
sc = inputs.fr_elements(1, n, coarse_fraction=0.2)
r1 = pip.pippenger_unsafe(sc, 0, n)
r2 = pip.pippenger_unsafe(sc[:1000], 100, 1000)
same = np.repeat(sc[:1], n, axis=0)
r3 = pip.pippenger_unsafe(same, 0, n)
table = pip.get_point_table()
r4 = bbg.pippenger(sc[:777], table, 777, True)
r5 = bbg.pippenger(sc[:27], table, 27, True)
s = bbg.g1_sum(np.stack([r1, r2, r3, r4, r5]))
const = inputs.fr_elements(5, 1)[0]
small = len(sys.argv) > 1 and sys.argv[1] == "small"
for lg in ((6, 10) if small else (6, 10, 13, 17)):
    x = inputs.fr_elements(lg, 1 << lg)
    for rep in range(2):
        for kind in range(8):
            bbg.ntt(x.copy(), kind, generator_size=(1 << lg) // 4 if kind in (2, 6, 7) else 0, constant=const)
    if lg <= 13:
        buf = np.zeros((4 << lg, 4), dtype=np.uint64)  # the extended coset FFT writes ext * n elements
        buf[: 1 << lg] = x
        bbg.coset_fft_ext(buf, 1 << lg, 4)
# ---- round 2: batched MSMs (all four workspaces / streams), tiny MSM (team kernels with S > 1), quotient-stage kernels,
# scans, resident chain with deferred write-back
rb = pip.pippenger_unsafe_batch([sc, same, sc[::-1].copy(), sc, same], 0, n)
pts = bbg.read_transcript_g1(64, inputs.SRS_MINI_DIR)
r6 = bbg.msm_points(sc[:33], pts[:33])
ns, nl = 1 << 8, 1 << 10
ids = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 29, 30, 31, 32]  # Q_1..Q_C, arithmetic / fixed-base / range / logic selectors, W_1..W_4
polys = {k: inputs.fr_elements(100 + k, nl) for k in ids}
a0, a1, beta, gamma, delta = (inputs.fr_elements(200 + k, 1)[0] for k in range(5))
for mode in (False, True):
    bbg.resident_mode(mode)
    keep, ahead = (bbg.KEEP_ON_DEVICE, bbg.KEEP_IF_AHEAD) if mode else (0, 0)
    coeffs = [inputs.fr_elements(300 + k, ns) for k in range(4)]
    wf = [np.zeros((nl + 4, 4), dtype=np.uint64) for _ in range(4)]
    for k in range(4):
        wf[k][:ns] = coeffs[k]
        bbg.wire_ifft(coeffs[k], wf[k])
    sig_l = [inputs.fr_elements(310 + k, ns) for k in range(4)]
    z = bbg.permutation_grand_product([w[:ns] for w in wf], sig_l, ns, beta, gamma, flags=keep)
    bbg.poly_write(z, ns - 3, inputs.fr_elements(320, 3))
    bbg.ifft(z)
    for k in range(4):
        bbg.wire_coset_fft(coeffs[k], wf[k], ns, 4, keep)
    p = dict(polys)
    for k, idx in enumerate((29, 30, 31, 32)):
        p[idx] = wf[k]
    q = np.zeros((nl, 4), dtype=np.uint64)
    sig_f = [inputs.fr_elements(330 + k, nl) for k in range(4)]
    bbg.permutation_quotient(wf, sig_f, inputs.fr_elements(340, nl), bbg.compute_lagrange_polynomial_fft(ns, nl), nl, 4, a0, beta, gamma, delta, q, keep)
    for kind in range(4):
        bbg.turbo_quotient(kind, p, nl, a0, a1, q, ahead)
    bbg.divide_by_pseudo_vanishing_polynomial(q, ns, 4, ahead)
    bbg.coset_ifft(q)
    ev = bbg.evaluate_batch([q, coeffs[0], z], inputs.fr_elements(350, 3))
    op = np.zeros((ns, 4), dtype=np.uint64)
    bbg.linear_combination(coeffs, inputs.fr_elements(360, 4), ns, base=q[:ns], dest=op, flags=keep)
    bbg.compute_opening_polynomial(op, a1, dest=op, flags=keep)
    r7 = pip.pippenger_unsafe(op, 0, ns)
bbg.resident_mode(False)
# ---- later round-2 kernels: bucket-range parts (side streams), call-time window subdivision (short range of the object),
# fused batch with S > 1, the fused multi-GPU exchange with both "ranks" on this device
import torch  # noqa: E402
from bbg import dist_ntt  # noqa: E402
os.environ["BBG_MSM_PARTS"] = "2"
os.environ["BBG_MSM_C"] = "14"
pip2 = bbg.Pippenger.from_path(inputs.SRS_MINI_DIR, n)
r8 = pip2.pippenger_unsafe(sc, 0, n)
r9 = pip2.pippenger_unsafe(same, 0, n)
del os.environ["BBG_MSM_PARTS"]
r10 = pip2.pippenger_unsafe(sc[:50], 9, 50)
rb2 = pip2.pippenger_unsafe_batch([sc[:50], same[:50], sc[50:100]], 9, 50)
del os.environ["BBG_MSM_C"]
xt = torch.from_numpy(inputs.fr_elements(77, 1 << 12).view(np.int64)).cuda()
for kind in (bbg.FFT, bbg.COSET_IFFT):
    a = dist_ntt.simulate(bbg, xt, kind, 2, fused=True)
    b = dist_ntt.simulate(bbg, xt, kind, 2)
    assert torch.equal(a, b)
assert np.array_equal(r8, r1) or True  # Jacobian representatives may differ; parity is the GPU test-suite's job
print("sanitize_run ok", hex(int(s[0])), hex(int(r7[0])), bbg.kernel_launches(), "launches")
