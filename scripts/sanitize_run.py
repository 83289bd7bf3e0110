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
print("sanitize_run ok", hex(int(s[0])), bbg.kernel_launches(), "launches")
