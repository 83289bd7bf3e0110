#!/usr/bin/env python
"""bench.py -- BN254 G1 MSM points/s (headline) and fr NTT elements/s on B200, beside the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n 20] [--ntt-log-n 22]

One "step" = one pass of the hot path over one batch of synthetic input: one MSM of 2^log_n points per rank
(BASELINE.json configs[1]: 2^20 on one B200).  For N > 1 (torchrun, one rank per GPU) the MSM is sharded by
contiguous point range exactly like Pippenger::pippenger_unsafe(scalars, from, range) + g1_sum
(bb/ecc/curves/bn254/scalar_multiplication/pippenger.cpp:27-31, c_bind.cpp:40-45): every rank owns 2^log_n bases
and scalars ("weak" scaling: n_total = N * 2^log_n), the only exchange is an all-gather of the 96-byte partial
sums followed by a one-warp g1 reduction on every rank.  The NTT family (fft / ifft / coset_fft at
2^ntt_log_n) is measured in the same run and reported under "ntt" (per-rank replicas when N > 1).

Printed by rank 0: ONE JSON line (contract in the task statement): value = device-resident throughput,
e2e = the same through the host-pointer C-ABI call (H2D of the scalars and D2H of the result inside the timed
region), roofline = dominant kernel vs the measured HBM peak, cpu_baseline = the unmodified reference
(oracle/_ref/libbbref.so) or the plain-C oracle timed on this box's host cores.

`--impl reference` times only the CPU reference on the same workload (rank 0; other ranks exit).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "aztec-2.0_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20, help="MSM points per GPU = 2^log_n")
    ap.add_argument("--ntt-log-n", type=int, default=22)
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (dev runs)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def srs_dir_and_capacity():
    """The reference's shipped 2^20-point transcript travels in oracle/_ref/srs_db (git-ignored); the committed
    4096-point excerpt is the fallback."""
    import inputs
    ref_srs_dir = os.path.join(ROOT, "oracle", "_ref", "srs_db")  # a DATA file of the reference; no oracle code is imported here
    full = os.path.join(ref_srs_dir, "transcript00.dat")
    if os.path.exists(full):
        return ref_srs_dir, 1 << 20
    return inputs.SRS_MINI_DIR, inputs.SRS_MINI_POINTS


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  The timed region of
    this bench is tens of milliseconds, far below nvidia-smi's polling period, so the samples come from NVML directly
    (nvidia_ml_py) on a background thread every ~1 ms; `nvidia-smi -lms` is only the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.thread = None
        self.path = "/tmp/bbg_clocks_%d_%d.csv" % (os.getpid(), index)
        self.sm, self.mx, self.reasons = [], [], set()

    def _nvml_loop(self):
        import pynvml as nv
        h = self.handle
        bits = []
        for name, const in (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                            ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap")):
            if hasattr(nv, const):
                bits.append((name, getattr(nv, const)))
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it is a plain index list
            phys = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            self.handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.mx = [float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))]
            self.stop_flag = False
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out = {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                       "samples": len(self.sm), "source": "nvml"}
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            with open(self.path) as f:
                for line in f:
                    c = [x.strip() for x in line.split(",")]
                    if len(c) < 9:
                        continue
                    try:
                        sm.append(float(c[1]))
                        mx.append(float(c[2]))
                    except ValueError:
                        continue
                    for k, name in enumerate(names):
                        if c[5 + k].lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}
        return out


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_checker():
    """The CPU arm: the compiled reference when it travelled (all host threads), else the scalar plain-C port.
    torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm must still use the whole host, so the
    thread count is set explicitly: the largest power of two <= the cores this process may run on (SURVEY.md 8d;
    barretenberg's own thread split assumes a power of two, bb/common/max_threads.hpp)."""
    from oracle import pyoracle as po
    if po.Ref.available():
        ref = po.Ref()
        try:
            avail = len(os.sched_getaffinity(0))
        except Exception:
            avail = os.cpu_count() or 1
        want = int(os.environ.get("BBG_CPU_THREADS", "0")) or (1 << (max(1, avail).bit_length() - 1))
        ref.set_num_threads(want)
        return ref, "reference"
    return po.Oracle(), "port"


def cpu_msm_setup(chk, kind, log_n):
    """(scalars, table/points, n, sample description) for the CPU arm.  The compiled reference runs the full
    workload; the scalar plain-C port (1 thread) runs a bounded 2^14-point sample of it."""
    import inputs
    from oracle import pyoracle as po
    srs_dir, cap = srs_dir_and_capacity()
    if kind == "reference":
        n = min(1 << log_n, cap)
        pts = chk.read_transcript_g1(n, srs_dir)
        table = chk.point_table(pts)
        sc = po.aligned_copy(inputs.fr_elements(1000, n))
        state = chk.new_runtime_state(n)
        run = lambda: chk.pippenger(sc, table, n=n, unsafe=True, state=state, copy=False)  # noqa: E731
        sample = "full workload: pippenger_unsafe over %d SRS points, runtime state pre-built" % n
        if n < (1 << log_n):
            sample = "bounded sample: pippenger_unsafe over %d SRS points (the shipped SRS holds 2^20)" % n
        return run, n, sample
    n = min(1 << min(log_n, 14), cap)
    pts = chk.read_transcript_g1(n, srs_dir if cap >= n else inputs.SRS_MINI_DIR)
    sc = inputs.fr_elements(1000, n)
    run = lambda: chk.pippenger(sc, pts, n=n, stride=1)  # noqa: E731
    return run, n, "bounded sample: %d points of the 2^%d workload (scalar plain-C port of the reference algorithm)" % (n, log_n)


def cpu_ntt_setup(chk, kind, log_n):
    import inputs
    from oracle import pyoracle as po
    lg = log_n if kind == "reference" else min(log_n, 16)
    n = 1 << lg
    x = po.aligned_copy(inputs.fr_elements(2000, n))
    if kind == "reference":
        chk.domain(n)  # evaluation_domain + compute_lookup_table outside the timed region
        run = lambda k: chk.ntt(k, x, inplace=True)  # noqa: E731
        sample = "full workload: fft, ifft, coset_fft in place on 2^%d elements, lookup tables pre-built" % lg
    else:
        run = lambda k: chk.ntt(k, x)  # noqa: E731
        sample = "bounded sample: 2^%d of 2^%d elements (scalar plain-C port)" % (lg, log_n)
    return run, n, sample


def time_cpu(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return ts


def host_cores(chk, kind):
    if kind == "reference":
        return chk.num_threads()
    return 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    chk, kind = cpu_checker()
    cores = host_cores(chk, kind)
    run, n, sample = cpu_msm_setup(chk, kind, args.log_n)
    ts = time_cpu(run, max(1, min(args.warmup, 2)), max(1, args.steps))
    sec = sum(ts) / len(ts)
    value = n / sec
    line = {
        "impl": "reference", "metric": "bn254_g1_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": len(ts), "warmup": max(1, min(args.warmup, 2)), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 limbs (254-bit Montgomery, x86-64 ADX/BMI2 asm)", "data": "synthetic",
        "config": {"workload": "BN254 G1 Pippenger MSM 2^%d (pippenger_unsafe, CPU, OpenMP)" % args.log_n, "points": n},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_ntt:
        nrun, nn, nsample = cpu_ntt_setup(chk, kind, args.ntt_log_n)
        per = {}
        for name, k in (("fft", 0), ("ifft", 1), ("coset_fft", 2)):
            t = time_cpu(lambda: nrun(k), 1, max(1, min(args.steps, 5)))
            per[name] = nn / (sum(t) / len(t))
        line["ntt"] = {"metric": "bn254_fr_ntt_elements_per_s", "value": statistics.mean(per.values()), "unit": "elements/s",
                       "per_kind": per, "log_n": args.ntt_log_n, "sample": nsample, "cores": cores, "kind": kind}
    emit(line)


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import bbg
    import inputs  # tests/inputs.py: seeded numpy generators only (the oracle is NOT imported by this arm)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    torch.cuda.set_device(local)
    bbg.init(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hbm_gbs, peak_src = peaks()
    K, W = args.steps, max(args.warmup, 3)
    n = 1 << args.log_n

    # ---- bases: the reference's SRS (rank 0's range); further ranks / sizes get distinct synthetic points
    #      P_i + D built on the device (never replicate points)
    srs_dir, cap = srs_dir_and_capacity()
    base_n = min(n, cap)
    srs = bbg.read_transcript_g1(base_n, srs_dir)  # decoded on the device by the library, like io::read_transcript_g1
    pts = torch.empty((n, 8), dtype=torch.int64, device=dev)
    srs_dev = torch.from_numpy(srs.view(np.int64)).to(dev)
    for blk in range((n + base_n - 1) // base_n):
        lo, hi = blk * base_n, min((blk + 1) * base_n, n)
        shift = rank * ((n + base_n - 1) // base_n) + blk
        if shift == 0:
            pts[lo:hi] = srs_dev[: hi - lo]
        else:
            bbg.g1_add_affine(srs_dev[: hi - lo].contiguous(), srs[shift % base_n], out_dev=pts[lo:hi])
    torch.cuda.synchronize()
    pip = bbg.Pippenger.from_device_points(pts, n)
    del pts, srs_dev

    sc_host = bbg.pinned_empty((n, 4))
    sc_host[...] = inputs.fr_elements(1000 + rank, n)
    sc_dev = torch.from_numpy(sc_host.view(np.int64)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gathered = torch.empty((world, 96), dtype=torch.uint8, device=dev)
    total = torch.empty(96, dtype=torch.uint8, device=dev)
    sync_token = torch.zeros(1, dtype=torch.float32, device=dev)

    def msm_step():
        part = pip.pippenger_unsafe(sc_dev, 0, n)  # async on torch's current stream
        if world > 1:
            dist.all_gather_into_tensor(gathered, part)
            bbg._check(bbg.lib.bbg_g1_sum_dev(gathered.data_ptr(), world, total.data_ptr(), torch.cuda.current_stream().cuda_stream))
            return total
        return part

    # ---- integer-pipe roofline measured live (fq multiplies / s)
    fq_muls = bbg.bench_field_mul(0, 2000)

    for _ in range(W):
        msm_step()
        flush.zero_()
    bbg.profile(True)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    launches0 = bbg.kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    phase_ms = {}
    for k in range(K):
        if world > 1:
            # align the ranks before every timed step (outside the event pair, like the L2 flush): without it the
            # host-side work between steps (flush launch, profile read-back) lets ranks drift, and the drift would be
            # booked as all-gather time by whichever rank arrives first
            dist.all_reduce(sync_token)
        ev[k][0].record()
        msm_step()
        ev[k][1].record()
        for name, ms in bbg.profile_read().items():  # synchronises on this step's last kernel
            phase_ms[name] = phase_ms.get(name, 0.0) + ms
        flush.zero_()  # L2 flush between timed iterations, outside the event pairs
    barrier()
    launches = bbg.kernel_launches() - launches0
    clk = clocks.stop()
    bbg.profile(False)
    ms_total = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    ms_per_step = ms_total / K
    value = world * n / (ms_per_step * 1e-3)

    # ---- e2e: host scalars -> C-ABI -> host result, every step (H2D 32 n bytes, D2H 96 bytes)
    def e2e_step():
        part = pip.pippenger_unsafe(sc_host, 0, n)  # numpy in / numpy out: bbg_pippenger_unsafe (host pointers)
        if world > 1:
            t = torch.from_numpy(part.view(np.uint8)).to(dev)
            dist.all_gather_into_tensor(gathered, t)
            return bbg.g1_sum(gathered.cpu().numpy().view(np.uint64).reshape(world, 12))
        return part
    for _ in range(2):
        result = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        result = e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / K
    e2e_value = world * n / e2e_s

    # ---- roofline of the dominant kernel (k_msm_accumulate): algorithmic bytes 96 n per launch
    acc_ms = phase_ms.get("msm_accumulate", 0.0) / K
    alg_bytes = 96.0 * n
    achieved = alg_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("k_msm_accumulate_2^%d" % args.log_n)
        except Exception:
            traffic = None
    Wn = {k: v / K for k, v in phase_ms.items() if k.startswith("msm")}
    c_bits = pip.window_bits()
    windows = (255 + c_bits - 1) // c_bits
    acc_muls = 10.0 * windows * n  # 8M + 2S per mixed addition, one per non-zero digit (upper bound)
    line = {
        "metric": "bn254_g1_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (254-bit Montgomery)",
        "data": "synthetic",
        "config": {"workload": "BN254 G1 Pippenger MSM 2^%d points per GPU (pippenger_unsafe), uniform fr scalars, SRS bases" % args.log_n,
                   "points_per_gpu": n, "points_total": world * n, "sharding": "contiguous point ranges + 96 B partial-sum all-gather" if world > 1 else "none",
                   "l2": "256 MiB flush written between timed iterations", "timing": "CUDA events per step on the launching stream, max over ranks"},
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 96, "ms_per_step": e2e_s * 1e3},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "k_msm_accumulate", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": achieved / hbm_gbs if hbm_gbs else None, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": acc_ms,
                     "note": "MSM is integer-pipe bound, not HBM bound (SURVEY.md 8d); see int_pipe"},
        "int_pipe": {"unit": "G fq-mul/s", "peak": fq_muls / 1e9, "peak_source": "bbg_bench_field_mul measured in this run",
                     "achieved": acc_muls / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else None,
                     "frac": (acc_muls / (acc_ms * 1e-3)) / fq_muls if acc_ms > 0 else None, "kernel": "k_msm_accumulate",
                     "model": "10 fq mul per mixed add x %d windows (c = %d, %d fixed-base levels) x n" % (windows, c_bits, pip.levels())},
        "phases_ms": Wn,
        "result_x_limb0": int(np.asarray(result).view(np.uint64)[0]),
    }

    # ---- NTT family at 2^ntt_log_n
    if not args.no_ntt:
        line["ntt"] = bench_ntt(args, bbg, torch, dev, inputs, K, W, hbm_gbs, peak_src, fq_muls, barrier, max_over_ranks, world, rank)

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu:
        chk, kind = cpu_checker()
        run, cn, sample = cpu_msm_setup(chk, kind, args.log_n)
        ts = time_cpu(run, 1, 5 if kind == "reference" else 1)
        cb = {"value": cn / (sum(ts) / len(ts)), "unit": "points/s", "cores": host_cores(chk, kind), "kind": kind, "sample": sample}
        if not args.no_ntt:
            nrun, nn, nsample = cpu_ntt_setup(chk, kind, args.ntt_log_n)
            per = {}
            for name, k in (("fft", 0), ("ifft", 1), ("coset_fft", 2)):
                t = time_cpu(lambda: nrun(k), 1, 3)
                per[name] = nn / (sum(t) / len(t))
            cb["ntt"] = {"value": statistics.mean(per.values()), "unit": "elements/s", "per_kind": per, "sample": nsample}
        line["cpu_baseline"] = cb
    elif rank == 0:
        line["cpu_baseline"] = None

    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def bench_ntt_sharded(args, bbg, torch, dev, inputs, K, W, hbm_gbs, peak_src, fq_muls, barrier, max_over_ranks, world, rank):
    """N > 1: ONE transform of world * 2^ntt_log_n elements, four-step with an NCCL all-to-all; every rank
    holds 2^ntt_log_n elements (weak scaling)."""
    import torch.distributed as dist
    from bbg import dist_ntt
    lg_local = args.ntt_log_n
    rb = world.bit_length() - 1
    if (1 << rb) != world:
        return {"unavailable": "the four-step NTT needs a power-of-two number of ranks"}
    n = (1 << lg_local) * world
    local = torch.from_numpy(inputs.fr_elements(2000 + rank, 1 << lg_local).view(np.int64)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    per = {}
    launches0 = bbg.kernel_launches()
    for name, kind in (("fft", bbg.FFT), ("ifft", bbg.IFFT), ("coset_fft", bbg.COSET_FFT)):
        for _ in range(W):
            dist_ntt.ntt_sharded(bbg, local, n, kind, rank, world)
        barrier()
        ms = 0.0
        for _ in range(K):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            dist.barrier()
            a.record()
            dist_ntt.ntt_sharded(bbg, local, n, kind, rank, world)
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        barrier()
        per[name] = {"ms": max_over_ranks(ms) / K}
        per[name]["elements_per_s"] = n / (per[name]["ms"] * 1e-3)
    launches = bbg.kernel_launches() - launches0
    mean_ms = statistics.mean(v["ms"] for v in per.values())
    return {
        "metric": "bn254_fr_ntt_elements_per_s", "value": n / (mean_ms * 1e-3), "unit": "elements/s", "log_n": lg_local + rb,
        "per_kind": per, "ms_per_transform": mean_ms, "gpu_launches": launches,
        "scaling": "weak: one 2^%d-point transform, 2^%d elements per GPU, four-step passes + one NCCL all-to-all of %d B per GPU"
                   % (lg_local + rb, lg_local, (32 << lg_local) * (world - 1) // world),
    }


def bench_ntt(args, bbg, torch, dev, inputs, K, W, hbm_gbs, peak_src, fq_muls, barrier, max_over_ranks, world, rank=0):
    if world > 1:
        return bench_ntt_sharded(args, bbg, torch, dev, inputs, K, W, hbm_gbs, peak_src, fq_muls, barrier, max_over_ranks, world, rank)
    lg = args.ntt_log_n
    n = 1 << lg
    x_host = bbg.pinned_empty((n, 4))
    x_host[...] = inputs.fr_elements(2000, n)
    x = torch.from_numpy(x_host.view(np.int64)).to(dev)
    kinds = (("fft", bbg.FFT), ("ifft", bbg.IFFT), ("coset_fft", bbg.COSET_FFT))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    per, pass_ms = {}, []
    launches0 = bbg.kernel_launches()
    bbg.profile(True)
    for name, kind in kinds:
        for _ in range(W):
            bbg.ntt(x, kind)
        barrier()
        ms = 0.0
        for _ in range(K):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            a.record()
            bbg.ntt(x, kind)
            b.record()
            ph = bbg.profile_read()
            pass_ms += [v for k2, v in ph.items() if k2.startswith("ntt_pass") and v > 0]
            ms += a.elapsed_time(b)
        barrier()
        per[name] = {"ms": max_over_ranks(ms) / K}
        per[name]["elements_per_s"] = world * n / (per[name]["ms"] * 1e-3)
    ntt_traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            ntt_traffic = json.load(f).get("k_ntt_pass_2^%d" % lg)
    except Exception:
        ntt_traffic = None
    bbg.profile(False)
    launches = bbg.kernel_launches() - launches0
    # e2e: host buffer in place through bbg_ntt (H2D + D2H of 32 n bytes each inside the call)
    bbg.ntt(x_host, bbg.FFT)
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, K // 2)):
        bbg.ntt(x_host, bbg.FFT)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / max(1, K // 2)
    mean_ms = statistics.mean(v["ms"] for v in per.values())
    kern_ms = statistics.mean(pass_ms) if pass_ms else None
    passes = 2 if lg <= 16 else (3 if lg <= 24 else 4)
    muls = (lg / 2.0 + passes - 1) * n
    return {
        "metric": "bn254_fr_ntt_elements_per_s", "value": world * n / (mean_ms * 1e-3), "unit": "elements/s", "log_n": lg,
        "per_kind": per, "ms_per_transform": mean_ms, "scaling": "replicas" if world > 1 else "single GPU",
        "e2e": {"value": world * n / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n, "ms_per_step": e2e_s * 1e3},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_ntt_pass", "achieved": (64.0 * n / (kern_ms * 1e-3) / 1e9) if kern_ms else None,
                     "peak": hbm_gbs, "unit": "GB/s", "frac": (64.0 * n / (kern_ms * 1e-3) / 1e9 / hbm_gbs) if kern_ms else None,
                     "traffic": ntt_traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": 64.0 * n, "kernel_ms": kern_ms,
                     "transform_frac": 64.0 * n / (mean_ms * 1e-3) / 1e9 / hbm_gbs,
                     "note": "each pass reads and writes the array once (64 n B); a transform is %d passes; the passes are integer-pipe bound" % passes},
        "int_pipe": {"unit": "G fr-mul/s", "peak": fq_muls / 1e9, "achieved": muls / (mean_ms * 1e-3) / 1e9,
                     "frac": muls / (mean_ms * 1e-3) / fq_muls, "model": "(log2 n / 2 + passes - 1) fr mul per element"},
    }


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL's version banner, library chatter) was
    re-routed to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
