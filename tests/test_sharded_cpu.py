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


# ------------------------------------------------------------------------------------------ four-step NTT plumbing
def _cpu_phase(orc, po, lg_a, lg_b, rb):
    """CPU stand-in for bbg_ntt_dist_dev with the SAME external contract (input shard = indices whose bits
    [lg_b - rb, lg_b) equal the rank, packed; chunk r' of the intermediate buffer goes to rank r'; output shard = indices
    whose bits [lg_a - rb, lg_a) equal the rank, packed), implemented as the textbook two-factor decomposition
    N = A * B,  X[ka + A kb] = sum_b w_B^(b kb) [ w_N^(b ka) sum_a x[a B + b] w_A^(a ka) ]  with the oracle's field ops."""
    import numpy as np
    A, B, R = 1 << lg_a, 1 << lg_b, 1 << rb
    Bl, Al = B // R, A // R  # local b values (phase 0) and local ka values (phase 1) per rank
    w_n = orc.fr_root_of_unity(lg_a + lg_b)
    one = orc.to_mont(po.FR, [1])[0]

    def phase(src, dst, n, kind, rank, world, which, generator_size, constant):
        assert kind == po.NTT_FFT and n == A * B and world == R
        x = src.numpy().view(np.uint64).reshape(-1, 4)
        out = np.zeros_like(x)
        if which == 0:
            cols = x.reshape(A, Bl, 4)  # packed position = a * Bl + (b - rank * Bl)
            for bl in range(Bl):
                b = rank * Bl + bl
                y = np.array(orc.reduce(po.FR, orc.ntt(po.NTT_FFT, cols[:, bl].copy())))  # over a -> ka
                step = one.copy()  # w_N^(b ka), ka = 0, 1, ...
                wb = one.copy()
                e, base = b, np.array(w_n, dtype=np.uint64)
                while e:  # wb = w_N^b
                    if e & 1:
                        wb = orc.field_op(po.FR, po.OP_MUL, wb, base)[0]
                    base = orc.field_op(po.FR, po.OP_MUL, base, base)[0]
                    e >>= 1
                for ka in range(A):
                    # intermediate layout: ka major, local b minor => the chunk of destination rank r' (ka's top bits) is contiguous
                    out[ka * Bl + bl] = orc.field_op(po.FR, po.OP_MUL, y[ka], step)[0]
                    step = orc.field_op(po.FR, po.OP_MUL, step, wb)[0]
        else:
            # received: for every source rank s, (Al local ka) x (Bl b values of s)
            rec = x.reshape(R, Al, Bl, 4)
            for kal in range(Al):
                row = np.concatenate([rec[s, kal] for s in range(R)], axis=0)  # all B values of b, in order
                xk = np.array(orc.reduce(po.FR, orc.ntt(po.NTT_FFT, row.copy())))  # over b -> kb
                for kb in range(B):
                    out[kb * Al + kal] = xk[kb]  # k = (rank * Al + kal) + A * kb, packed in index order
        dst.copy_(torch_from(out))

    def torch_from(a):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int64))

    return phase


def _ntt_worker(rank, world, port, lg, out_dir):
    import torch
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "aztec-2.0_b200", "python"), os.path.join(ROOT, "aztec-2.0_b200", "python", "bbg")):
        sys.path.insert(0, p)
    import inputs
    import dist_ntt
    from oracle import pyoracle as po
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    orc = po.Oracle()
    n = 1 << lg
    rb = world.bit_length() - 1
    lg_a = lg_b = lg // 2  # ntt_device's two-pass factorisation for lg <= 16 (equal digits when lg is even)
    in_pos, out_pos = lg_b - rb, lg_a - rb
    x = inputs.fr_elements(777, n)
    shard = torch.from_numpy(np.ascontiguousarray(dist_ntt.extract_shard(x, in_pos, world, rank)).view(np.int64))
    out = dist_ntt.ntt_sharded(None, shard, n, po.NTT_FFT, rank, world, phase_fn=_cpu_phase(orc, po, lg_a, lg_b, rb))
    np.save(os.path.join(out_dir, "ntt_r%d.npy" % rank), out.numpy().view(np.uint64))
    dist.destroy_process_group()


def test_ntt_four_step_plumbing_gloo_world2(tmp_path):
    """dist_ntt.ntt_sharded (phase 0 -> ONE all_to_all_single -> phase 1) and the shard layout helpers under gloo with two
    CPU processes; the device phases are replaced by a textbook two-factor stand-in that honours the same layout contract
    as bbg_ntt_dist_dev (the real kernels are parity-tested on the GPU: test_ntt_multi_gpu_data_path_simulated)."""
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "aztec-2.0_b200", "python"), os.path.join(ROOT, "aztec-2.0_b200", "python", "bbg")):
        sys.path.insert(0, p)
    import inputs
    import dist_ntt
    from oracle import pyoracle as po
    lg, world = 12, 2
    n = 1 << lg
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_ntt_worker, args=(world, port, lg, str(tmp_path)), nprocs=world, join=True)
    orc = po.Oracle()
    x = inputs.fr_elements(777, n)
    exp = np.array(orc.reduce(po.FR, orc.ntt(po.NTT_FFT, x)))
    # the layout the library itself reports for this size (pure host code) must be the one the stand-in used
    import bbg
    assert bbg.ntt_dist_layout(n, world) == (lg // 2 - 1, lg // 2 - 1)
    out = np.zeros((n, 4), dtype=np.uint64)
    for r in range(world):
        shard = np.load(os.path.join(str(tmp_path), "ntt_r%d.npy" % r)).reshape(-1, 4)
        dist_ntt.insert_shard(out, shard, lg // 2 - 1, world, r)
    assert np.array_equal(out, exp)


def _ntt_natural_worker(rank, world, port, lg, out_dir):
    import torch
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "aztec-2.0_b200", "python"), os.path.join(ROOT, "aztec-2.0_b200", "python", "bbg")):
        sys.path.insert(0, p)
    import inputs
    import dist_ntt
    from oracle import pyoracle as po
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    orc = po.Oracle()
    n = 1 << lg
    rb = world.bit_length() - 1
    lg_a = lg_b = lg // 2
    m = n // world
    x = inputs.fr_elements(778, n)
    block = torch.from_numpy(np.ascontiguousarray(x[rank * m:(rank + 1) * m]).view(np.int64))  # NATURAL contiguous block
    out = dist_ntt.ntt_sharded_natural(None, block, n, po.NTT_FFT, rank, world, phase_fn=_cpu_phase(orc, po, lg_a, lg_b, rb),
                                       layout=(lg_b - rb, lg_a - rb))
    np.save(os.path.join(out_dir, "nat_r%d.npy" % rank), out.numpy().view(np.uint64))
    dist.destroy_process_group()


def test_ntt_natural_block_in_block_out_gloo_world2(tmp_path):
    """input and output both block-distributed in natural order (ntt_sharded_natural = re-distribution
    all-to-all + four-step core + re-distribution all-to-all), two CPU processes under gloo"""
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "aztec-2.0_b200", "python")):
        sys.path.insert(0, p)
    import inputs
    from oracle import pyoracle as po
    lg, world = 12, 2
    n = 1 << lg
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_ntt_natural_worker, args=(world, port, lg, str(tmp_path)), nprocs=world, join=True)
    orc = po.Oracle()
    exp = np.array(orc.reduce(po.FR, orc.ntt(po.NTT_FFT, inputs.fr_elements(778, n))))
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "nat_r%d.npy" % r)).reshape(-1, 4) for r in range(world)])
    assert np.array_equal(got, exp)
