"""torchrun --nproc-per-node N scripts/dist_check.py : parity of the real multi-process paths (NCCL).
 * MSM sharded by point range + 96-byte partial all-gather + g1_sum  == oracle MSM over the whole range
 * four-step NTT with all_to_all_single                                  == the single-GPU transform of the same array
 * natural contiguous blocks in and out with every movement a peer load / store of the NTT passes (bbg_ntt_dist_natural_dev)
 * the same with the exchange fused into the pass before it (peer stores over NVLink, CUDA IPC buffers), three
   transforms back to back per size so that both receive buffers of the double-buffering are used
Prints one line per rank; exits non-zero on mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "aztec-2.0_b200", "python")):
    sys.path.insert(0, p)
os.environ["NCCL_DEBUG"] = "WARN"
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bbg  # noqa: E402
import inputs  # noqa: E402
from bbg import dist_ntt, sharded  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    bbg.init(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    orc = po.Oracle()
    ok = True

    # ---- MSM
    n = 4000
    pts = orc.read_transcript_g1(inputs.SRS_MINI_POINTS, inputs.SRS_MINI_DIR)
    sc = inputs.fr_elements(77, n, coarse_fraction=0.1)
    pip = bbg.Pippenger.from_points(pts)
    lo, cnt = sharded.shard_range(n, rank, world)
    res = sharded.msm_sharded(lambda s, f, r: pip.pippenger_unsafe(s, f, r), bbg.g1_sum, sc[lo:lo + cnt], n, rank, world, device=dev)
    exp = orc.jac_to_buffer(orc.pippenger(sc, pts[:n], n=n, stride=1))
    ok &= orc.jac_to_buffer(res) == exp

    # ---- NTT
    for lg in (14, 20):
        nn = 1 << lg
        x = inputs.fr_elements(88 + lg, nn, coarse_fraction=0.25)
        in_pos, out_pos = bbg.ntt_dist_layout(nn, world)
        xch = dist_ntt.FusedExchange(bbg, nn, rank, world, natural=True) if world <= 8 else None
        for kind, gs in ((bbg.FFT, 0), (bbg.COSET_FFT, nn // 4), (bbg.COSET_IFFT, 0)):
            shard = torch.from_numpy(np.ascontiguousarray(dist_ntt.extract_shard(x, in_pos, world, rank)).view(np.int64)).to(dev)
            out = dist_ntt.ntt_sharded(bbg, shard, nn, kind, rank, world, generator_size=gs).cpu().numpy().view(np.uint64)
            full = bbg.ntt(x.copy(), kind, generator_size=gs)
            want = np.ascontiguousarray(dist_ntt.extract_shard(full, out_pos, world, rank))
            ok &= bool(np.array_equal(orc.reduce(po.FR, out), orc.reduce(po.FR, want)))
            if xch is not None:
                outs = [dist_ntt.ntt_sharded_fused(bbg, shard, nn, kind, rank, world, xch, generator_size=gs) for _ in range(3)]
                for o in outs:
                    fused_ok = bool(np.array_equal(orc.reduce(po.FR, o.cpu().numpy().view(np.uint64)), orc.reduce(po.FR, want)))
                    if not fused_ok:
                        print("rank %d: fused exchange mismatch lg=%d kind=%d" % (rank, lg, kind), flush=True)
                    ok &= fused_ok
                # natural blocks in / out, every movement a peer load or store of the passes themselves; twice in a row
                mm = nn // world
                block = torch.from_numpy(np.ascontiguousarray(x[rank * mm:(rank + 1) * mm]).view(np.int64)).to(dev)
                for rep in range(2):
                    o = dist_ntt.ntt_natural_fused(bbg, block, nn, kind, rank, world, xch, generator_size=gs).clone()
                    nat_ok = bool(np.array_equal(orc.reduce(po.FR, o.cpu().numpy().view(np.uint64)), orc.reduce(po.FR, full[rank * mm:(rank + 1) * mm])))
                    if not nat_ok:
                        print("rank %d: natural-block transform mismatch lg=%d kind=%d rep=%d" % (rank, lg, kind, rep), flush=True)
                    ok &= nat_ok
        if xch is not None:
            xch.close()
    print("rank %d/%d: %s" % (rank, world, "OK" if ok else "MISMATCH"), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


main()
