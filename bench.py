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
    ap.add_argument("--strong-log-n", type=int, default=26, help="total points of the strong-scaling record (0 disables)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the size sweep and the extra configurations (N = 1)")
    ap.add_argument("--sweep-sizes", default="16,18,20,22,24,26")
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
    W = max(args.warmup, 3)  # the same warm-up rule as our arm
    ts = time_cpu(run, W, max(1, args.steps))
    sec = sum(ts) / len(ts)
    value = n / sec
    line = {
        "impl": "reference", "metric": "bn254_g1_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": len(ts), "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 limbs (254-bit Montgomery, x86-64 ADX/BMI2 asm)", "data": "synthetic",
        "config": {"workload": "BN254 G1 Pippenger MSM 2^%d points per GPU (pippenger_unsafe), uniform fr scalars, SRS bases" % args.log_n,
                   "points_per_gpu": n, "points_total": args.gpus * n},
        "measurement": {"arm": "the unmodified reference's CPU path (OpenMP, x86-64 ADX/BMI2 asm) on this box's host cores; one 2^%d MSM per step whatever --gpus is" % args.log_n},
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
FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
BLOCK_LOG = 20  # synthetic bases and scalars are defined per block of 2^20 points, independent of how they are sharded


def geometric_scalars(n, seed=5489):
    """config #1's scalar distribution (bb/plonk/pippenger_bench/main.cpp:52-58): element = random, accumulator = element,
    then accumulator *= element for every scalar; Montgomery-form limbs."""
    rng = np.random.default_rng(seed)
    e = int.from_bytes(rng.bytes(32), "little") % FR_MODULUS
    acc, out = e, np.empty((n, 4), dtype=np.uint64)
    R = 1 << 256
    for i in range(n):
        acc = acc * e % FR_MODULUS
        m = acc * R % FR_MODULUS
        out[i] = [(m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]
    return out


class Bases:
    """Global, sharding-independent base set: block 0 = the reference's SRS (monomial 0 = the generator), block b >= 1 =
    SRS + D_b with D_b = SRS point b: distinct points, never replicated.  A rank builds any range."""

    def __init__(self, bbg, torch, dev):
        self.bbg, self.torch, self.dev = bbg, torch, dev
        self.srs_dir, self.cap = srs_dir_and_capacity()
        self.block = min(1 << BLOCK_LOG, self.cap)
        self.srs = bbg.read_transcript_g1(self.block, self.srs_dir)  # decoded on the device by the library
        self.srs_dev = torch.from_numpy(self.srs.view(np.int64)).to(dev)

    def build(self, first, count):
        """device tensor (count, 8) int64 = global points [first, first + count)"""
        torch, bbg = self.torch, self.bbg
        pts = torch.empty((count, 8), dtype=torch.int64, device=self.dev)
        done = 0
        while done < count:
            g = first + done
            blk, off = divmod(g, self.block)
            take = min(self.block - off, count - done)
            src = self.srs_dev[off:off + take]
            if blk == 0:
                pts[done:done + take] = src
            else:
                bbg.g1_add_affine(src.contiguous(), self.srs[blk % self.block], out_dev=pts[done:done + take])
            done += take
        torch.cuda.synchronize()
        return pts

    def pippenger(self, first, count):
        pts = self.build(first, count)
        pip = self.bbg.Pippenger.from_device_points(pts, count)
        del pts
        return pip


def block_scalars(inputs, first, count, out=None, seed0=1000):
    """global scalars [first, first + count): block b is fr_elements(seed0 + b) (sharding-independent)"""
    out = np.empty((count, 4), dtype=np.uint64) if out is None else out
    blk = 1 << BLOCK_LOG
    done = 0
    while done < count:
        g = first + done
        b, off = divmod(g, blk)
        take = min(blk - off, count - done)
        out[done:done + take] = inputs.fr_elements(seed0 + b, blk)[off:off + take]
        done += take
    return out


class Env:
    pass


def time_msm(env, pip, n, sc_dev, sc_host, K, W, profile=False, e2e_steps=None):
    """device-timed (value) and end-to-end (host scalars through the C-ABI, partials combined on the device) MSM timing.
    Returns dict with ms_per_step, e2e_ms, phases, launches, result (numpy Jacobian of the combined result)."""
    bbg, torch, dist, dev, world = env.bbg, env.torch, env.dist, env.dev, env.world
    gathered = torch.empty((world, 96), dtype=torch.uint8, device=dev)
    total = torch.empty(96, dtype=torch.uint8, device=dev)

    def combine(part_dev):
        if world == 1:
            return part_dev
        dist.all_gather_into_tensor(gathered, part_dev)
        bbg._check(bbg.lib.bbg_g1_sum_dev(gathered.data_ptr(), world, total.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return total

    def step():
        return combine(pip.pippenger_unsafe(sc_dev, 0, n))  # async on torch's current stream

    for _ in range(W):
        step()
        env.flush.zero_()
    if profile:
        bbg.profile(True)
    env.barrier()
    launches0 = bbg.kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    phase_ms = {}
    for k in range(K):
        if world > 1:
            dist.all_reduce(env.sync_token)  # align the ranks before every timed step, outside the event pair
        ev[k][0].record()
        step()
        ev[k][1].record()
        if profile:
            for name, ms in bbg.profile_read().items():  # synchronises on this step's last kernel
                phase_ms[name] = phase_ms.get(name, 0.0) + ms
        env.flush.zero_()  # L2 flush between timed iterations, outside the event pairs
    env.barrier()
    launches = bbg.kernel_launches() - launches0
    if profile:
        bbg.profile(False)
    ms = env.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / K

    # e2e: host scalars -> bbg_pippenger_unsafe (H2D inside) -> 96-byte partial -> device-side combine -> host
    def e2e_step():
        part = pip.pippenger_unsafe(sc_host, 0, n)  # numpy in / numpy out: the host-pointer C-ABI call
        if world == 1:
            return part
        t = torch.from_numpy(part.view(np.uint8)).to(dev, non_blocking=True)
        return combine(t).cpu().numpy().view(np.uint64)

    EK = K if e2e_steps is None else e2e_steps
    for _ in range(2):
        result = e2e_step()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(EK):
        result = e2e_step()
    env.barrier()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0) / EK
    return {"ms": ms, "e2e_ms": e2e_s * 1e3, "phases": {k: v / K for k, v in phase_ms.items() if k.startswith("msm")},
            "launches": launches, "result": np.asarray(result, dtype=np.uint64).reshape(12)}


def msm_record(env, t, n_total, n_local, pip, hbm_gbs):
    """value / e2e / roofline of one timed MSM configuration"""
    acc_ms = t["phases"].get("msm_accumulate", 0.0)
    rec = {"points_total": n_total, "points_per_gpu": n_local, "ms_per_step": t["ms"], "value": n_total / (t["ms"] * 1e-3),
           "e2e": {"value": n_total / (t["e2e_ms"] * 1e-3), "unit": "points/s", "ms_per_step": t["e2e_ms"],
                   "h2d_bytes_per_step": 32 * n_local, "d2h_bytes_per_step": 96},
           "window_bits": pip.window_bits(), "levels": pip.levels()}
    if acc_ms > 0:
        ach = 96.0 * n_local / (acc_ms * 1e-3) / 1e9
        rec["roofline"] = {"bound": "hbm", "kernel": "k_msm_accumulate", "achieved": ach, "peak": hbm_gbs, "unit": "GB/s",
                           "frac": ach / hbm_gbs, "kernel_ms": acc_ms}
        windows = (255 + pip.window_bits() - 1) // pip.window_bits()
        rec["int_pipe_frac"] = (10.0 * windows * n_local / (acc_ms * 1e-3)) / env.fq_muls
    return rec


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

    env = Env()
    env.bbg, env.torch, env.dist, env.dev, env.world, env.rank = bbg, torch, dist, dev, world, rank

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

    env.barrier, env.max_over_ranks = barrier, max_over_ranks
    env.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    env.sync_token = torch.zeros(1, dtype=torch.float32, device=dev)
    hbm_gbs, peak_src = peaks()
    K, W = args.steps, max(args.warmup, 3)
    n = 1 << args.log_n
    bases = Bases(bbg, torch, dev)

    # ---- integer-pipe roofline measured live (fq multiplies / s)
    env.fq_muls = fq_muls = bbg.bench_field_mul(0, 2000)

    # ================= headline: weak scaling, 2^log_n points per GPU (BASELINE configs[1] at N = 1) =================
    t_setup = time.perf_counter()
    pip = bases.pippenger(rank * n, n)  # rank r owns global points [r n, (r + 1) n)
    setup_s = time.perf_counter() - t_setup
    sc_host = bbg.pinned_empty((n, 4))
    block_scalars(inputs, rank * n, n, out=sc_host)
    sc_dev = torch.from_numpy(sc_host.view(np.int64)).to(dev)
    clocks = ClockSampler(local)
    clocks.start()
    t = time_msm(env, pip, n, sc_dev, sc_host, K, W, profile=True)
    clk = clocks.stop()
    ms_per_step = t["ms"]
    value = world * n / (ms_per_step * 1e-3)
    result = t["result"]
    acc_ms = t["phases"].get("msm_accumulate", 0.0)
    alg_bytes = 96.0 * n
    achieved = alg_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj.get("k_msm_accumulate_2^%d" % args.log_n)
            traffic_src = tj.get("source")
        except Exception:
            traffic = None
    c_bits = pip.window_bits()
    windows = (255 + c_bits - 1) // c_bits
    acc_muls = 10.0 * windows * n  # 8M + 2S per mixed addition, one per non-zero digit (upper bound)
    level_bytes = pip.levels() * n * 64
    line = {
        "metric": "bn254_g1_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (254-bit Montgomery)",
        "data": "synthetic",
        "config": {"workload": "BN254 G1 Pippenger MSM 2^%d points per GPU (pippenger_unsafe), uniform fr scalars, SRS bases" % args.log_n,
                   "points_per_gpu": n, "points_total": world * n},
        "measurement": {"sharding": "contiguous point ranges + 96 B partial-sum all-gather + device-side g1 sum" if world > 1 else "none",
                        "l2": "256 MiB flush written between timed iterations", "timing": "CUDA events per step on the launching stream, max over ranks",
                        "fixed_base_precompute": "%d levels 2^(%d l) P_i of the %d bases resident in HBM (%.0f MiB), built once per Pippenger object in %.0f ms (outside the timed region, like the reference's endomorphism table + runtime state)"
                                                 % (pip.levels(), c_bits, n, level_bytes / 2**20, setup_s * 1e3)},
        "e2e": {"value": world * n / (t["e2e_ms"] * 1e-3), "unit": "points/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 96,
                "ms_per_step": t["e2e_ms"], "host_memory": "pinned (bbg_malloc = the reference's bbmalloc hook)"},
        "gpu_launches": t["launches"],
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "k_msm_accumulate", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": achieved / hbm_gbs if hbm_gbs else None, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": acc_ms,
                     "note": "MSM is integer-pipe bound, not HBM bound (SURVEY.md 8d); see int_pipe"},
        "int_pipe": {"unit": "G fq-mul/s", "peak": fq_muls / 1e9, "peak_source": "bbg_bench_field_mul measured in this run",
                     "achieved": acc_muls / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else None,
                     "frac": (acc_muls / (acc_ms * 1e-3)) / fq_muls if acc_ms > 0 else None, "kernel": "k_msm_accumulate",
                     "model": "10 fq mul per mixed add x %d windows (c = %d, %d fixed-base levels) x n" % (windows, c_bits, pip.levels())},
        "phases_ms": t["phases"],
        "result_affine": bbg.g1_normalize(result.reshape(1, 12))[0].tolist() if rank == 0 else None,
    }

    # ---- e2e from PAGEABLE host memory (what barretenberg's aligned_alloc buffers are, bb/common/mem.hpp:26-44)
    if world == 1:
        pageable = np.array(sc_host, copy=True)
        for _ in range(2):
            pip.pippenger_unsafe(pageable, 0, n)
        t0 = time.perf_counter()
        for _ in range(K):
            pip.pippenger_unsafe(pageable, 0, n)
        pg = (time.perf_counter() - t0) / K
        line["e2e_pageable"] = {"value": n / pg, "unit": "points/s", "ms_per_step": pg * 1e3, "h2d_bytes_per_step": 32 * n,
                                "d2h_bytes_per_step": 96, "host_memory": "pageable numpy array"}
        # batched: 4 MSMs per call (the prover's W_1..W_4), device-resident scalars, four streams / four workspaces
        devs = [sc_dev, sc_dev.clone(), sc_dev.clone(), sc_dev.clone()]
        for _ in range(2):
            pip.pippenger_unsafe_batch(devs, 0, n)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(max(1, K // 2)):
            pip.pippenger_unsafe_batch(devs, 0, n)
        b.record()
        torch.cuda.synchronize()
        bms = a.elapsed_time(b) / max(1, K // 2) / 4
        line["batched"] = {"msms_per_call": 4, "ms_per_msm": bms, "value": n / (bms * 1e-3), "unit": "points/s",
                           "note": "bbg_pippenger_unsafe_batch_dev: chain i on stream i mod 4 with its own workspace, tails overlap the next accumulation (MSMs of <= 2^17 points are fused four to a chain)"}
        del devs

    # ================= parity of the N > 1 path, outside every timed region =================
    if world > 1:
        line["parity"] = parity_msm(env, bases, inputs, pip, sc_host, n, result)

    # ================= config #5: strong scaling, 2^strong_log_n points in total =================
    if args.strong_log_n > 0:
        del pip
        line["strong_2p%d" % args.strong_log_n] = strong_scaling(args, env, bases, inputs, hbm_gbs, K, W)
    else:
        del pip

    # ---- NTT family at 2^ntt_log_n
    if not args.no_ntt:
        line["ntt"] = bench_ntt(args, bbg, torch, dev, inputs, K, W, hbm_gbs, peak_src, fq_muls, barrier, max_over_ranks, world, rank)

    # ---- size sweep + the other configurations the metric names (N = 1)
    if world == 1 and not args.no_sweep:
        line["sweep"] = sweep(args, env, bases, inputs, hbm_gbs)
        line["msm_no_precompute"] = msm_no_precompute(env, bases, inputs, 1 << min(args.log_n, 20), K)
        line["config1_geometric_2p16"] = config1(env, bases, K)

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
                tt = time_cpu(lambda: nrun(k), 1, 3)
                per[name] = nn / (sum(tt) / len(tt))
            cb["ntt"] = {"value": statistics.mean(per.values()), "unit": "elements/s", "per_kind": per, "sample": nsample}
        line["cpu_baseline"] = cb
        if not args.no_sweep:
            line["prover"] = prover_record()
    elif rank == 0:
        line["cpu_baseline"] = None

    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def parity_msm(env, bases, inputs, pip, sc_host, n, combined):
    """N > 1, outside the timed region: the sharded total (all-gather of partials + g1 sum on every rank) must equal
    the g1 sum of the from/range partials of EVERY rank's range recomputed on rank 0 alone (pippenger.cpp:27-31 +
    c_bind.cpp:40-45), compared as canonical affine points; and every rank must hold the same total."""
    bbg, torch, dist, dev, world, rank = env.bbg, env.torch, env.dist, env.dev, env.world, env.rank
    mine = bbg.g1_normalize(np.asarray(combined).reshape(1, 12))[0]
    ok = True
    if rank == 0:
        parts = []
        for r in range(world):
            if r == 0:
                parts.append(pip.pippenger_unsafe(np.array(sc_host, copy=True), 0, n))
            else:
                p_r = bases.pippenger(r * n, n)
                parts.append(p_r.pippenger_unsafe(block_scalars(inputs, r * n, n), 0, n))
                p_r.close()
        expect = bbg.g1_normalize(bbg.g1_sum(np.stack(parts)).reshape(1, 12))[0]
        ok = bool(np.array_equal(expect, mine))
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # every rank holds the same combined point
    t = torch.from_numpy(mine.view(np.int64).copy()).to(dev)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    return {"msm": bool(flag.item() == 1) and same,
            "msm_check": "rank 0 recomputed all %d ranges alone (from/range partials + g1_sum) == sharded total, canonical affine; all ranks agree" % world}


def strong_scaling(args, env, bases, inputs, hbm_gbs, K, W):
    """BASELINE configs[4]: ONE MSM of 2^strong_log_n points sharded over the ranks by contiguous point range (2^26 / N
    per GPU), bases = SRS prefix + distinct synthetic points, combined by all-gather + device-side g1 sum."""
    bbg, torch, dist, dev, world, rank = env.bbg, env.torch, env.dist, env.dev, env.world, env.rank
    n_total = 1 << args.strong_log_n
    if n_total % world:
        return {"unavailable": "world size does not divide the point count"}
    n = n_total // world
    first = rank * n
    t0 = time.perf_counter()
    pip = bases.pippenger(first, n)
    build_s = time.perf_counter() - t0
    sc_host = bbg.pinned_empty((n, 4))
    block_scalars(inputs, first, n, out=sc_host, seed0=3000)
    sc_dev = torch.from_numpy(sc_host.view(np.int64)).to(dev)
    steps = max(3, min(K, 5))
    t = time_msm(env, pip, n, sc_dev, sc_host, steps, W, profile=True, e2e_steps=steps)
    rec = msm_record(env, t, n_total, n, pip, hbm_gbs)
    rec["steps"] = steps
    rec["phases_ms"] = t["phases"]
    rec["scaling"] = "strong"
    rec["build_s"] = build_s
    rec["result_affine"] = bbg.g1_normalize(t["result"].reshape(1, 12))[0].tolist()
    rec["workload"] = ("one BN254 G1 MSM of 2^%d points sharded over %d GPU(s) by contiguous point range; the result is the same point for every N "
                       "(compare result_affine across the N = 1, 2, 4, 8 lines)" % (args.strong_log_n, world))
    if world > 1:
        # parity: rank 0 recomputes the LAST rank's range alone and compares with the partial that rank produced
        last = world - 1
        part = pip.pippenger_unsafe(np.array(sc_host, copy=True), 0, n)
        g = torch.empty((world, 96), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(g, torch.from_numpy(part.view(np.uint8)).to(dev))
        parts = g.cpu().numpy().view(np.uint64).reshape(world, 12)
        ok = True
        if rank == 0:
            pip.close()
            p_l = bases.pippenger(last * n, n)
            mine = p_l.pippenger_unsafe(block_scalars(inputs, last * n, n, seed0=3000), 0, n)
            p_l.close()
            a = bbg.g1_normalize(np.stack([mine, parts[last]]))
            total = bbg.g1_normalize(bbg.g1_sum(parts).reshape(1, 12))[0]
            ok = bool(np.array_equal(a[0], a[1])) and bool(np.array_equal(total, np.array(rec["result_affine"], dtype=np.uint64)))
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        rec["parity"] = bool(flag.item() == 1)
        rec["parity_check"] = "rank 0 recomputed rank %d's point range alone == that rank's partial; g1_sum of the gathered partials == combined result" % last
    pip.close()
    bbg.pinned_free(sc_host)
    return rec


def sweep(args, env, bases, inputs, hbm_gbs):
    """MSM and NTT at 2^16 ... 2^26 on one GPU: device-timed value, e2e through the host-pointer C-ABI, roofline each."""
    bbg, torch, dev = env.bbg, env.torch, env.dev
    out = {"msm": {}, "ntt": {}}
    sizes = [int(x) for x in args.sweep_sizes.split(",") if x]
    for lg in sizes:
        n = 1 << lg
        if lg == args.strong_log_n:
            out["msm"]["2^%d" % lg] = {"see": "strong_2p%d (at N = 1 that record IS the single-GPU 2^%d MSM)" % (lg, lg)}
            continue
        try:
            pip = bases.pippenger(0, n)
            sc_host = bbg.pinned_empty((n, 4))
            block_scalars(inputs, 0, n, out=sc_host, seed0=3000)
            sc_dev = torch.from_numpy(sc_host.view(np.int64)).to(dev)
            steps = 10 if lg <= 20 else (5 if lg <= 24 else 3)
            t = time_msm(env, pip, n, sc_dev, sc_host, steps, 3, profile=True, e2e_steps=steps)
            rec = msm_record(env, t, n, n, pip, hbm_gbs)
            rec["phases_ms"] = t["phases"]
            out["msm"]["2^%d" % lg] = rec
            pip.close()
            bbg.pinned_free(sc_host)
            del sc_dev
        except Exception as e:  # a size that does not fit must not take the headline down with it
            out["msm"]["2^%d" % lg] = {"error": str(e)[:200]}
    blk = inputs.fr_elements(2000, 1 << 16)
    for lg in sizes:
        n = 1 << lg
        try:
            x_host = bbg.pinned_empty((n, 4))
            x_host.reshape(-1, 1 << 16, 4)[...] = blk  # timing input: a 2^16-element random block tiled
            x = torch.from_numpy(x_host.view(np.int64)).to(dev)
            steps = 10 if lg <= 22 else 4
            per = {}
            for name, kind in (("fft", bbg.FFT), ("ifft", bbg.IFFT), ("coset_fft", bbg.COSET_FFT)):
                for _ in range(3):
                    bbg.ntt(x, kind)
                torch.cuda.synchronize()
                ms = 0.0
                for _ in range(steps):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    env.flush.zero_()
                    a.record()
                    bbg.ntt(x, kind)
                    b.record()
                    torch.cuda.synchronize()
                    ms += a.elapsed_time(b)
                per[name] = ms / steps
            bbg.ntt(x_host, bbg.FFT)
            t0 = time.perf_counter()
            for _ in range(max(2, steps // 2)):
                bbg.ntt(x_host, bbg.FFT)
            e2e_s = (time.perf_counter() - t0) / max(2, steps // 2)
            mean_ms = statistics.mean(per.values())
            out["ntt"]["2^%d" % lg] = {
                "ms": per, "value": n / (mean_ms * 1e-3), "unit": "elements/s",
                "e2e": {"value": n / e2e_s, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n},
                "roofline": {"bound": "hbm", "achieved": 64.0 * n / (mean_ms * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                             "frac": 64.0 * n / (mean_ms * 1e-3) / 1e9 / hbm_gbs, "note": "whole transform: 64 n algorithmic bytes / transform time"},
                "int_pipe_frac": ntt_muls_per_element(lg) * n / (mean_ms * 1e-3) / env.fq_muls}
            bbg.pinned_free(x_host)
            del x
        except Exception as e:
            out["ntt"]["2^%d" % lg] = {"error": str(e)[:200]}
    return out


def msm_no_precompute(env, bases, inputs, n, K):
    """The L = 1 path: bases the library has never seen (MemReferenceString users, bbg_pippenger with an unregistered table,
    bbg_msm_points): W bucket sets and the 255-doubling window Horner instead of fixed-base levels."""
    bbg, torch, dev = env.bbg, env.torch, env.dev
    pts = bases.build(0, n)
    sc = torch.from_numpy(inputs.fr_elements(1000, n).view(np.int64)).to(dev)
    out = torch.empty(96, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def step():
        bbg._check(bbg.lib.bbg_msm_points_dev(sc.data_ptr(), pts.data_ptr(), 1, n, out.data_ptr(), st))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    pts_h = pts.cpu().numpy().view(np.uint64)
    sc_h = sc.cpu().numpy().view(np.uint64)
    bbg.msm_points(sc_h, pts_h)
    t0 = time.perf_counter()
    for _ in range(max(2, K // 2)):
        bbg.msm_points(sc_h, pts_h)
    e2e = (time.perf_counter() - t0) / max(2, K // 2)
    return {"points": n, "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "points/s",
            "e2e": {"value": n / e2e, "ms_per_step": e2e * 1e3, "h2d_bytes_per_step": 96 * n, "d2h_bytes_per_step": 96,
                    "note": "bbg_msm_points: bases AND scalars uploaded from pageable memory every call"}}


def config1(env, bases, K):
    """BASELINE configs[0] on the GPU: the reference bench's own workload (bb/plonk/pippenger_bench/main.cpp:40-75):
    pippenger_unsafe over the first 2^16 SRS points with geometric-sequence scalars."""
    bbg, torch, dev = env.bbg, env.torch, env.dev
    n = min(1 << 16, bases.block)
    pip = bases.pippenger(0, n)
    sc = geometric_scalars(n)
    sc_dev = torch.from_numpy(sc.view(np.int64)).to(dev)
    for _ in range(3):
        pip.pippenger_unsafe(sc_dev, 0, n)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        pip.pippenger_unsafe(sc_dev, 0, n)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    pip.pippenger_unsafe(sc, 0, n)
    t0 = time.perf_counter()
    for _ in range(K):
        r = pip.pippenger_unsafe(sc, 0, n)
    e2e = (time.perf_counter() - t0) / K
    pip.close()
    return {"points": n, "scalars": "geometric sequence acc *= e (pippenger_bench/main.cpp:52-58)", "ms_per_step": ms, "value": n / (ms * 1e-3),
            "unit": "points/s", "e2e": {"value": n / e2e, "ms_per_step": e2e * 1e3, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 96,
                                       "host_memory": "pageable"},
            "result_affine_x0": int(bbg.g1_normalize(r.reshape(1, 12))[0][0])}


def prover_record(reps_gpu=8, reps_cpu=3):
    """BASELINE configs[3]: the reference's TurboPLONK join-split prover (n = 2^16) with its hot path resolved to libbbg.
    The CALLER is the unmodified reference prover, built as a test harness under oracle/_ref (oracle/js_harness.cpp, three link
    flavours); the thing measured is this repository's library and shims underneath it.  Proof bytes are compared with the
    all-CPU build in the same run."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    srs = os.path.join(ref, "srs_db")
    bins = {k: os.path.join(ref, k) for k in ("js_prover_gpu", "js_prover_gpu_l1", "js_prover_cpu")}
    if not all(os.path.exists(b) for b in bins.values()) or not os.path.exists(os.path.join(srs, "transcript00.dat")):
        return {"unavailable": "oracle/_ref/js_prover_* not built (needs the reference tree at build time)"}

    def run(binary, reps, env=None):
        e = dict(os.environ)
        e.update(env or {})
        p = subprocess.run([bins[binary], srs, str(reps)], capture_output=True, text=True, env=e, timeout=600)
        if p.returncode != 0:
            return None, p.stderr[-300:]
        d = json.loads(p.stdout.strip().splitlines()[-1])
        stats = {}
        for line in p.stderr.splitlines():
            if line.startswith('{"bbg_stats"'):
                st = json.loads(line)
                stats[st["bbg_stats"]] = st
        return d, stats
    out = {"circuit": "rollup::proofs::join_split (63 398 gates, n = 2^16), noop transaction, deterministic engine"}
    ref_proof = None
    for name, reps, env in (("js_prover_cpu", reps_cpu, {}), ("js_prover_gpu_l1", reps_gpu, {"BBG_STATS": "1"}), ("js_prover_gpu", reps_gpu, {"BBG_STATS": "1"})):
        d, stats = run(name, reps, env)
        if d is None:
            out[name] = {"error": stats}
            continue
        times = [p["construct_proof_s"] for p in d["proofs"]]
        # the first proofs of a process pay context creation, module load, first-touch of every mirror and clock ramp-up (tens to
        # hundreds of ms on a cold box): the steady state is the median of the second half of the run
        steady = sorted(times[len(times) // 2:] if len(times) > 3 else times)
        rec = {"construct_proof_ms_all": [round(t * 1e3, 3) for t in times], "construct_proof_ms": steady[len(steady) // 2] * 1e3,
               "keygen_s": d["keygen_s"], "keygen_warm_s": d.get("keygen_warm_s"), "cuda_init_s": d.get("cuda_init_s"), "verified": d["verified"],
               "kernel_launches": d["gpu_kernel_launches"]}
        if "h2d_bytes" in d["proofs"][-1] and name != "js_prover_cpu":
            rec["pcie_bytes_per_proof"] = {"h2d": d["proofs"][-1]["h2d_bytes"], "d2h": d["proofs"][-1]["d2h_bytes"]}
        if name == "js_prover_cpu":
            ref_proof = d["first_proof"]
        else:
            rec["first_proof_identical_to_cpu"] = (ref_proof is not None and d["first_proof"] == ref_proof)
        out[name] = rec
    try:
        out["speedup_vs_cpu"] = out["js_prover_cpu"]["construct_proof_ms"] / out["js_prover_gpu"]["construct_proof_ms"]
        l1, full = out["js_prover_gpu_l1"]["pcie_bytes_per_proof"], out["js_prover_gpu"]["pcie_bytes_per_proof"]
        out["pcie_bytes_per_proof_l1_over_resident"] = (l1["h2d"] + l1["d2h"]) / max(1, full["h2d"] + full["d2h"])
    except Exception:
        pass
    return out


def ntt_passes(lg):
    return 2 if lg <= 16 else (3 if lg <= 24 else 4)


def ntt_muls_per_element(lg, loge=None):
    """Exact fr multiplies per element of one fft: every radix-2 stage multiplies half of the elements, except in the
    last radix-E round of each pass (row bits LOGE-1..0), where the butterflies whose twiddle is 1 are skipped: stage s of
    that round skips a 2^-s share (ntt.cu radix_round, `tail`); plus one inter-pass twiddle per element per non-final pass."""
    P = ntt_passes(lg)
    loge = (1 if lg <= 16 else 2) if loge is None else loge
    skipped = sum(2.0 ** -s for s in range(loge))  # in units of "stages"
    return (lg - P * skipped) / 2.0 + (P - 1)


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
    # the same transforms with the exchange FUSED into the pass before it (peer stores over NVLink, no NCCL all-to-all)
    fused = None
    if world <= 8:
        try:
            xch = dist_ntt.FusedExchange(bbg, n, rank, world, natural=True)
            fused = {}
            for name, kind in (("fft", bbg.FFT), ("ifft", bbg.IFFT), ("coset_fft", bbg.COSET_FFT)):
                for _ in range(W):
                    dist_ntt.ntt_sharded_fused(bbg, local, n, kind, rank, world, xch)
                barrier()
                ms = 0.0
                for _ in range(K):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    flush.zero_()
                    dist.barrier()
                    a.record()
                    dist_ntt.ntt_sharded_fused(bbg, local, n, kind, rank, world, xch)
                    b.record()
                    torch.cuda.synchronize()
                    ms += a.elapsed_time(b)
                barrier()
                fused[name] = {"ms": max_over_ranks(ms) / K}
            # natural contiguous blocks in and out: all-to-all re-distributions (3 NCCL all-to-alls) against peer
            # loads / stores issued by the passes themselves (no collective but three one-word all-reduces)
            natural = {}
            for name, fn in (("nccl_3_all_to_alls", lambda: dist_ntt.ntt_sharded_natural(bbg, local, n, bbg.FFT, rank, world)),
                             ("peer_memory", lambda: dist_ntt.ntt_natural_fused(bbg, xch.in_view, n, bbg.FFT, rank, world, xch))):
                xch.in_view.copy_(local)
                for _ in range(W):
                    fn()
                barrier()
                ms = 0.0
                for _ in range(K):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    flush.zero_()
                    dist.barrier()
                    a.record()
                    fn()
                    b.record()
                    torch.cuda.synchronize()
                    ms += a.elapsed_time(b)
                barrier()
                natural[name] = {"fft_ms": max_over_ranks(ms) / K}
        except Exception as e:  # IPC peer mapping unavailable on this box: the NCCL path above is the record
            fused = {"unavailable": str(e)[:200]}
            xch = None
    # parity (outside the timed region): this rank's output shard == the matching slice of a single-GPU transform of the
    # gathered input, canonical (reduce_once'd) limbs, for a forward and a coset transform
    in_pos, out_pos = bbg.ntt_dist_layout(n, world)
    gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=dev)
    dist.all_gather_into_tensor(gathered, local)
    full_in = torch.empty((n, 4), dtype=local.dtype, device=dev)
    for r in range(world):
        dist_ntt.insert_shard(full_in, gathered[r], in_pos, world, r)
    del gathered
    ok = fused_ok = natural_ok = True
    m_blk = n // world
    for kind in (bbg.FFT, bbg.COSET_IFFT):
        full = full_in.clone()
        bbg.ntt(full, kind)
        if fused is not None and "unavailable" not in fused:
            got3 = dist_ntt.ntt_natural_fused(bbg, full_in[rank * m_blk:(rank + 1) * m_blk].contiguous(), n, kind, rank, world, xch)
            natural_ok = natural_ok and bool(torch.equal(bbg.field_op_dev(1, 7, got3.clone()), bbg.field_op_dev(1, 7, full[rank * m_blk:(rank + 1) * m_blk].contiguous())))
        want = dist_ntt.extract_shard(full, out_pos, world, rank).contiguous()
        got = dist_ntt.ntt_sharded(bbg, local, n, kind, rank, world)
        ok = ok and bool(torch.equal(bbg.field_op_dev(1, 7, got), bbg.field_op_dev(1, 7, want)))
        if fused is not None and "unavailable" not in fused:
            got2 = dist_ntt.ntt_sharded_fused(bbg, local, n, kind, rank, world, xch)
            fused_ok = fused_ok and bool(torch.equal(bbg.field_op_dev(1, 7, got2), bbg.field_op_dev(1, 7, want)))
            del got2
        del full, want, got
    flag = torch.tensor([1 if ok else 0, 1 if fused_ok else 0, 1 if natural_ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if fused is not None and "unavailable" not in fused:
        xch.close()
        fused = {"per_kind": fused, "ms_per_transform": statistics.mean(v["ms"] for v in fused.values()),
                 "parity": bool(flag[1].item() == 1),
                 "natural_blocks": dict(natural, parity=bool(flag[2].item() == 1),
                                        what="fft with natural contiguous blocks in and out (SURVEY 8e): three NCCL all-to-alls vs peer loads / stores by the passes"),
                 "how": "the pass before the exchange stores into the owners' receive buffers over NVLink peer memory (CUDA IPC); "
                        "a one-word all-reduce orders the last pass after every rank's stores"}
    mean_ms = statistics.mean(v["ms"] for v in per.values())
    return {
        "metric": "bn254_fr_ntt_elements_per_s", "value": n / (mean_ms * 1e-3), "unit": "elements/s", "log_n": lg_local + rb,
        "per_kind": per, "ms_per_transform": mean_ms, "gpu_launches": launches, "parity": bool(flag[0].item() == 1), "fused_exchange": fused,
        "parity_check": "every rank's output shard == the same slice of a single-GPU transform of the gathered input (fft, coset_ifft), canonical limbs",
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
    # e2e: host buffer in place through bbg_ntt (H2D + D2H of 32 n bytes each inside the call); pinned, then pageable
    def host_e2e(buf):
        bbg.ntt(buf, bbg.FFT)
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, K // 2)):
            bbg.ntt(buf, bbg.FFT)
        barrier()
        return max_over_ranks(time.perf_counter() - t0) / max(1, K // 2)
    e2e_s = host_e2e(x_host)
    e2e_pg = host_e2e(np.array(x_host, copy=True))
    mean_ms = statistics.mean(v["ms"] for v in per.values())
    kern_ms = statistics.mean(pass_ms) if pass_ms else None
    passes = ntt_passes(lg)
    muls = ntt_muls_per_element(lg) * n
    return {
        "metric": "bn254_fr_ntt_elements_per_s", "value": world * n / (mean_ms * 1e-3), "unit": "elements/s", "log_n": lg,
        "per_kind": per, "ms_per_transform": mean_ms, "scaling": "replicas" if world > 1 else "single GPU",
        "e2e": {"value": world * n / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n, "ms_per_step": e2e_s * 1e3,
                "host_memory": "pinned (bbg_malloc)"},
        "e2e_pageable": {"value": world * n / e2e_pg, "unit": "elements/s", "ms_per_step": e2e_pg * 1e3, "host_memory": "pageable numpy array"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_ntt_pass", "achieved": (64.0 * n / (kern_ms * 1e-3) / 1e9) if kern_ms else None,
                     "peak": hbm_gbs, "unit": "GB/s", "frac": (64.0 * n / (kern_ms * 1e-3) / 1e9 / hbm_gbs) if kern_ms else None,
                     "traffic": ntt_traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": 64.0 * n, "kernel_ms": kern_ms,
                     "transform_frac": 64.0 * n / (mean_ms * 1e-3) / 1e9 / hbm_gbs,
                     "note": "each pass reads and writes the array once (64 n B); a transform is %d passes; the passes are integer-pipe bound" % passes},
        "int_pipe": {"unit": "G fr-mul/s", "peak": fq_muls / 1e9, "achieved": muls / (mean_ms * 1e-3) / 1e9,
                     "frac": muls / (mean_ms * 1e-3) / fq_muls,
                     "model": "%.2f fr mul per element of an fft: radix-2 stages at 1/2 mul per element minus the twiddle-free butterflies of each pass's last round, plus %d inter-pass twiddles (ntt_muls_per_element in bench.py)"
                              % (ntt_muls_per_element(lg), passes - 1)},
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
