"""N > 1 host logic on CPU: world_size-2 gloo processes run the MSM sharding (range split + partial all-gather +
g1 sum) with the oracle standing in for the device kernels."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    sys.path.insert(0, os.path.join(ROOT, "aztec-2.0_b200", "python", "bbg"))
    import sharded
    for n in (0, 1, 7, 8, 9, 4096, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                lo, cnt = sharded.shard_range(n, r, world)
                assert lo == pos and cnt >= 0
                pos += cnt
            assert pos == n
    with pytest.raises(ValueError):
        sharded.shard_range(10, 2, 2)


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "aztec-2.0_b200", "python", "bbg")):
        sys.path.insert(0, p)
    import inputs
    import sharded
    from oracle import pyoracle as po
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    orc = po.Oracle()
    pts = orc.read_transcript_g1(inputs.SRS_MINI_POINTS, inputs.SRS_MINI_DIR)
    sc = inputs.fr_elements(4242, n)
    lo, cnt = sharded.shard_range(n, rank, world)
    res = sharded.msm_sharded(lambda s, f, r: orc.pippenger(s, pts[f:f + r], n=r, stride=1) if r else orc.g1_infinity(),
                              lambda parts: orc.g1_sum(parts), sc[lo:lo + cnt], n, rank, world)
    with open(os.path.join(out_dir, "r%d.hex" % rank), "w") as f:
        f.write(orc.jac_to_buffer(res).hex())
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 1])
def test_msm_sharded_gloo_world2(tmp_path, n):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import inputs
    from oracle import pyoracle as po
    port = 29500 + (os.getpid() % 2000) + n % 7
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    orc = po.Oracle()
    pts = orc.read_transcript_g1(inputs.SRS_MINI_POINTS, inputs.SRS_MINI_DIR)
    exp = orc.jac_to_buffer(orc.pippenger(inputs.fr_elements(4242, n), pts[:n], n=n, stride=1)).hex()
    got = [open(os.path.join(str(tmp_path), "r%d.hex" % r)).read() for r in range(2)]
    assert got[0] == got[1] == exp
