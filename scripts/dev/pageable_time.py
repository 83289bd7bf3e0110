"""Developer timing: end-to-end MSM 2^20 from PAGEABLE host scalars for the staging-pool settings (BBG_STAGING_THREADS)."""
import os
import sys
import time
sys.path[:0] = ['.', 'tests', 'aztec-2.0_b200/python']
import numpy as np
import bbg
import inputs
from oracle import pyoracle as po
bbg.init(0)
n = 1 << 20
pip = bbg.Pippenger.from_path(po.REF_SRS_DIR, n)
sc = np.array(inputs.fr_elements(7, n), copy=True)
pin = bbg.pinned_empty((n, 4))
pin[...] = sc
for name, arr in (("pageable", sc), ("pinned", pin)):
    for _ in range(3):
        pip.pippenger_unsafe(arr, 0, n)
    t0 = time.perf_counter()
    for _ in range(10):
        pip.pippenger_unsafe(arr, 0, n)
    print("BBG_STAGING_THREADS=%s %s: %.3f ms per MSM" % (os.environ.get("BBG_STAGING_THREADS", "default"), name, (time.perf_counter() - t0) * 100))
# raw host memcpy bandwidth of this box, one thread
dst = np.empty_like(sc)
t0 = time.perf_counter()
for _ in range(10):
    np.copyto(dst, sc)
print("numpy copy 32 MB: %.3f ms (%.1f GB/s, one thread)" % ((time.perf_counter() - t0) * 100, 32 * 1.048576 / ((time.perf_counter() - t0) * 100)))
