"""Dev probe (GPU box): where does a host-pointer NTT call spend its time?  pageable vs pinned host buffers vs
device-resident, per size; prints wall ms and the library's own kernel-only ms (bbg_last_device_ms)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "aztec-2.0_b200", "python")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bbg  # noqa: E402
import inputs  # noqa: E402

bbg.init(0)
for lg in (16, 18, 20, 22):
    n = 1 << lg
    x = inputs.fr_elements(lg, n)
    pinned = bbg.pinned_empty((n, 4))
    pinned[...] = x
    dev = torch.from_numpy(x.view(np.int64)).cuda()
    for name, buf in (("pageable", x.copy()), ("pinned", pinned)):
        for _ in range(3):
            bbg.ifft(buf)
        t0 = time.perf_counter()
        dms = 0.0
        reps = 10
        for _ in range(reps):
            bbg.ifft(buf)
            dms += bbg.last_device_ms()
        wall = (time.perf_counter() - t0) / reps * 1e3
        print("2^%d %-9s wall %.3f ms  kernel-only %.3f ms" % (lg, name, wall, dms / reps))
    for _ in range(3):
        bbg.ifft(dev)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        bbg.ifft(dev)
    b.record()
    torch.cuda.synchronize()
    print("2^%d device    %.3f ms per transform (10 back to back)" % (lg, a.elapsed_time(b) / 10))
    t0 = time.perf_counter()
    for _ in range(10):
        bbg.ifft(dev)
        torch.cuda.synchronize()
    print("2^%d device    %.3f ms per transform (synchronised after each)" % (lg, (time.perf_counter() - t0) / 10 * 1e3))
