"""Developer benchmark: per-phase device times of the MSM and NTT at a few sizes (not the graded bench)."""
import sys, os, time
sys.path[:0] = ['.', 'tests', 'aztec-2.0_b200/python']
import numpy as np, torch, bbg, inputs
from oracle import pyoracle as po

def main():
    bbg.init(0)
    logs = [int(x) for x in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['16', '20']) if x]
    ntt_logs = [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else ['16', '20', '22', '24']) if x]
    srs_dir = po.REF_SRS_DIR if os.path.exists(os.path.join(po.REF_SRS_DIR, 'transcript00.dat')) else inputs.SRS_MINI_DIR
    nmax = 1 << max(logs) if logs else 0
    if srs_dir == inputs.SRS_MINI_DIR:
        nmax = min(nmax, 4096)
    if logs:
        t0 = time.time(); pip = bbg.Pippenger.from_path(srs_dir, nmax); print('srs load+decode %.3fs' % (time.time() - t0))
    plain = os.environ.get('DEVBENCH_PLAIN', '0') == '1'  # no per-phase events: mean of 10 calls, L2 flushed in between
    bbg.profile(not plain)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for lg in logs:
        n = min(1 << lg, nmax)
        sc = torch.from_numpy(inputs.fr_elements(7, n).view(np.int64)).cuda()
        if plain:
            tot = 0.0
            for it in range(13):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush.zero_()
                e0.record(); out = pip.pippenger_unsafe(sc, 0, n); e1.record(); torch.cuda.synchronize()
                if it >= 3: tot += e0.elapsed_time(e1)
            print('MSM 2^%d: %.3f ms (mean of 10, no phase events)' % (lg, tot / 10))
            continue
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = pip.pippenger_unsafe(sc, 0, n); e1.record(); torch.cuda.synchronize()
        ph = bbg.profile_read()
        print('MSM 2^%d: %.3f ms  %.1f Mpts/s | ' % (lg, e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e3),
              ' '.join('%s=%.3f' % (k[4:], v) for k, v in ph.items() if k.startswith('msm')))
    for lg in ntt_logs:
        n = 1 << lg
        x = torch.from_numpy(inputs.fr_elements(8, n).view(np.int64)).cuda()
        for kind in (bbg.FFT, bbg.COSET_FFT, bbg.COSET_IFFT):
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); bbg.ntt(x, kind); e1.record(); torch.cuda.synchronize()
            ph = bbg.profile_read()
            ms = e0.elapsed_time(e1)
            print('NTT kind %d 2^%d: %.3f ms  %.1f Melem/s  %.1f GB/s(64n) | ' % (kind, lg, ms, n / ms / 1e3, 64 * n / ms / 1e6),
                  ' '.join('%s=%.3f' % (k[4:], v) for k, v in ph.items() if k.startswith('ntt')))
main()
