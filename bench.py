#!/usr/bin/env python
"""bench.py -- atom-frames/s of the time-correlation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload fft|windowed|helfand|helfand_direct] [--atoms A] [--frames T]

Default workload = BASELINE.json configs[3]: VelocityAutocorr fft=True,
100,000 atoms x 10,000 frames FP64 on one B200 (the largest single-GPU
configuration; per-GPU work is the same at every N: weak scaling, N=8 is
800,000 atoms).  One "step" = one pass of the hot path over the whole batch.

  value  atom-frames/s with the series already resident in HBM (ta_vacf_fft:
         kernel K1 + the atom-sum + NCCL all-reduce + D2H of the T-vector),
         device-timed with CUDA events, max over ranks.
  e2e    the same metric through the public class
         VelocityAutocorr(universe.atoms, fft=True).run(): pinned host float32
         trajectory -> H2D -> K0 transposition -> K1 -> timeseries on the host,
         every step, wall-clock, max over ranks.
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".

`--impl reference` times the oracle's restatement of the reference's CPU path
(numpy + pocketfft, one process per host core, atoms split between them) on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line.  Libraries write there too (NCCL prints "NCCL version ..." to stdout when
# NCCL_DEBUG=VERSION is set in the environment, NCCL_DEBUG_FILE notwithstanding), so file descriptor 1 is pointed at stderr
# for the whole run and the JSON line goes to the saved descriptor.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: str):
    if _REAL_STDOUT is None:
        print(line, flush=True)
        return
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (line + "\n").encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic work per atom-frame (xyz, FP64)
BYTES_PER_AF = {"fft": 32.0, "windowed": 32.0, "helfand": 56.0, "helfand_direct": 56.0}
FP64_NOMINAL_TFLOPS = 37.0   # 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz (no measured FP64 peak is provided)

DEFAULTS = {          # workload -> (atoms per GPU, frames)   BASELINE.json configs[3], [1], [2]
    "fft": (100_000, 10_000),
    "windowed": (1_000, 2_000),
    "helfand": (10_000, 5_000),          # ViscosityHelfand as shipped: S1 - 2 S2 by FFT + exact refinement (K1 + K5 + K6)
    "helfand_direct": (10_000, 5_000),   # ViscosityHelfand(fft=False): the direct O(T^2) lag sums (K3)
}


def flops_per_af(workload, T, D=3):
    if workload in ("fft", "helfand"):
        L = 1 << int(np.ceil(np.log2(2 * T - 1)))
        return (D + 1) * 2.5 * L * np.log2(L) / T
    if workload == "windowed":
        return D * (T + 1.0)
    return 1.5 * D * (T - 1.0)


# ---------------------------------------------------------------- helpers
class ClockSampler:
    """Samples SM clock, power and throttle reasons during the timed region (NVML, else nvidia-smi)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.dev, self.samples, self._stop, self._th = device_index, [], threading.Event(), None

    # NVML clock-event reason bits (nvml.h nvmlClocksEventReason*)
    REASON_BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def _sample_nvml(self, nv, h):
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = int(get(h))
        flags = {name: ("Active" if bits & bit else "Not Active") for name, bit in self.REASON_BITS}
        return [str(sm), str(mx), f"{pw:.2f}", flags["hw_slowdown"], flags["hw_thermal_slowdown"],
                flags["sw_thermal_slowdown"], flags["sw_power_cap"]]

    def _run(self):
        # NVML in-process (a sample every 10 ms: the timed region of the default run lasts a quarter of a second);
        # nvidia-smi as a subprocess (one sample takes ~0.1 s) when NVML is not importable
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.dev)
            self._sample_nvml(nv, h)
        except Exception:
            nv = None
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.samples.append(self._sample_nvml(nv, h))
                    self._stop.wait(0.01)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "reasons": sorted(reasons),
                "samples": len(self.samples)}


def fill_random_f32(arr, seed, threads=None):
    """Seeded N(0,1) float32 fill, chunked over threads (numpy releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    flat = arr.reshape(-1)
    threads = threads or min(32, os.cpu_count() or 1)
    chunk = 1 << 24
    jobs = [(i, min(i + chunk, flat.size)) for i in range(0, flat.size, chunk)]

    def work(j):
        lo, hi = jobs[j]
        np.random.Generator(np.random.Philox(key=seed, counter=[j, 0, 0, 0])).standard_normal(
            hi - lo, dtype=np.float32, out=flat[lo:hi])

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(len(jobs))))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def k1_kernel_name(plan):
    """Name of the K1 kernel the library picked, from the plan it reports (ta_fft_plan_info)."""
    rad = plan["radices"]
    if rad[1:] == [16, 16]:
        return f"k1f_fft_acf<{rad[0]},{plan['threads']}> (K1 three-pass radix-16 path, bulk series prefetch)"
    if rad[1:] == [8, 8, 8]:
        return f"k1e_fft_acf<{rad[0]}> (K1 four-pass radix-8 path)"
    return "k1_fft_acf<double> (K1 general mixed-radix kernel)"


def load_profile_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu summary."""
    path = os.path.join(ROOT, "profiles", "roofline_latest.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return d.get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ---------------------------------------------------------------- reference arm (CPU)
def _ref_worker(args):
    workload, T, natoms, seed = args
    import oracle

    rng = np.random.default_rng(seed)
    vel = rng.standard_normal((T, natoms, 3), dtype=np.float32).astype(np.float64)
    t0 = time.perf_counter()
    if workload == "helfand_direct":
        workload = "helfand"          # the reference has one Helfand route: the O(T^2) lag loop
    if workload == "fft":
        oracle.vacf_fft(vel)
    elif workload == "windowed":
        oracle.vacf_windowed(vel)
    else:
        pos = np.cumsum(vel, axis=0)
        lags = np.unique(np.linspace(1, T - 1, 5).astype(int))
        oracle.helfand_msd(vel, pos, np.ones(natoms), np.full(T, 8000.0), 300.0, lags=lags)
        # cost of a lag is proportional to (T - lag): extrapolate to all lags
        frac = float(np.sum(T - lags)) / float(np.sum(T - np.arange(1, T)))
        return (time.perf_counter() - t0) / frac
    return time.perf_counter() - t0


def cpu_sample_size(workload, T):
    """atoms per worker so that one sample is a few seconds of numpy work."""
    if workload == "fft":
        return max(8, int(4e7 / (T * np.log2(T) * 6)))
    if workload == "helfand_direct":
        return 64
    if workload == "windowed":
        return max(1, int(1.2e9 / (T * T * 3 * 8)))
    return 64


def run_reference(args, rank, world):
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor

    workload = args.workload
    A, T = DEFAULTS[workload]
    A = args.atoms or A
    T = args.frames or T
    cores = os.cpu_count() or 1
    per = cpu_sample_size(workload, T)
    times = []
    with ProcessPoolExecutor(cores) as ex:
        for step in range(args.warmup + args.steps):
            # all workers run concurrently; the step ends when the slowest is done
            # (workers time only the correlation itself, not their input generation)
            dts = list(ex.map(_ref_worker, [(workload, T, per, 1000 * step + w) for w in range(cores)]))
            if step >= args.warmup:
                times.append(max(dts))
    ms = 1e3 * float(np.mean(times))
    sample_af = cores * per * T
    value = sample_af / (ms / 1e3)
    sample = (f"{cores} processes x {per} atoms x {T} frames per step (oracle restatement of the reference's "
              f"numpy/pocketfft path; throughput is linear in atoms)")
    if workload in ("helfand", "helfand_direct"):
        sample += "; 5 sampled lags extrapolated by sum(T-lag)"
    line = {
        "impl": "reference", "metric": "atom-frames/s", "value": value, "unit": "atom-frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(workload, A, T, args.gpus),
        "cpu_baseline": {"value": value, "unit": "atom-frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "atom-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


def workload_config(workload, A, T, gpus):
    names = {"fft": "VelocityAutocorr fft=True dim_type=xyz (BASELINE.json configs[3])",
             "windowed": "VelocityAutocorr fft=False dim_type=xyz (BASELINE.json configs[1])",
             "helfand": "ViscosityHelfand dim_type=xyz (BASELINE.json configs[2])",
             "helfand_direct": "ViscosityHelfand dim_type=xyz fft=False, direct lag sums (BASELINE.json configs[2])"}
    return {"workload": names[workload], "atoms_per_gpu": A, "atoms_total": A * gpus, "frames": T,
            "dims": 3, "sharding": f"atoms x{gpus}", "l2": "inputs larger than L2 (no flush needed)"
            if A * T * 24 > 2 * 126e6 else "L2 flushed between steps"}


def bind_near_gpu(device_index):
    """Pin this process (and the threads it starts later) to the CPUs of the GPU's NUMA node, before the host
    trajectory is allocated: first touch then places the pages next to the PCIe root the copies leave from.
    What `numactl --cpunodebind` would do for a one-process-per-GPU launch; returns a description for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:           # nvml prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return {"node": None, "why": "single NUMA node"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"node": node, "why": "no allowed CPU on that node"}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception as e:  # no nvml / sysfs: run unbound
        return {"node": None, "why": f"{type(e).__name__}: {e}"[:80]}


# ---------------------------------------------------------------- B200 arm
def run_b200(args, rank, world, local_rank):
    from transport_analysis_b200 import _lib
    from transport_analysis_b200.synthetic import make_universe
    from transport_analysis_b200.velocityautocorr import VelocityAutocorr
    from transport_analysis_b200.viscosity import ViscosityHelfand

    numa = bind_near_gpu(local_rank) if not args.no_numa_bind else {"node": None, "why": "--no-numa-bind"}
    dist = None
    if world > 1:
        import torch.distributed as dist  # plumbing only: rendezvous, barrier, max-over-ranks

        dist.init_process_group("gloo", rank=rank, world_size=world)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    workload = args.workload
    A, T = DEFAULTS[workload]
    A = args.atoms or A
    T = args.frames or T
    helf = workload in ("helfand", "helfand_direct")

    # ---- NCCL communicator shared by the ranks (the library owns it)
    nccl_id = None
    if world > 1:
        box = [_lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]

    # ---- synthetic trajectory of this rank's atoms in pinned host memory (float32, frame-major)
    t_setup = time.perf_counter()
    vel = np.empty((T, A, 3), dtype=np.float32)
    fill_random_f32(vel, seed=1234 + rank)
    _lib.host_register(vel)
    pos = None
    if helf:
        pos = np.empty((T, A, 3), dtype=np.float32)
        fill_random_f32(pos, seed=4321 + rank)
        pos *= 10.0
        _lib.host_register(pos)
    masses = np.full(A, 15.999) if helf else None
    u = make_universe(pos, vel, masses=masses, dimensions=[20.0, 20.0, 20.0, 90.0, 90.0, 90.0] if helf else None)
    setup_s = time.perf_counter() - t_setup

    # ---- the public-API object; in multi-rank mode it runs on this rank's context
    if world > 1:
        ctx = _lib.Context([local_rank], rank=rank, nranks=world, nccl_id=nccl_id)
        dev_arg = ctx
    else:
        ctx = None
        dev_arg = [local_rank]
    if helf:
        ana = ViscosityHelfand(u.atoms, devices=dev_arg, fft=("auto" if workload == "helfand" else False))
    else:
        ana = VelocityAutocorr(u.atoms, fft=(workload == "fft"), devices=dev_arg)

    h2d = vel.nbytes * (2 if helf else 1)
    d2h = 8 * T

    # ---- e2e leg: full run() through the class, host buffers, every step
    e2e_times = []
    sampler = ClockSampler(local_rank)
    for step in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        ana.run()
        dt = time.perf_counter() - t0
        barrier()
        if step >= args.warmup:
            e2e_times.append(max_over_ranks(dt))
    ctx = ana._ctx
    ts_e2e = np.array(ana.results.timeseries)

    # ---- device-resident leg: the series are in HBM, time the compute call
    def compute():
        if workload == "fft":
            return ctx.vacf_fft()
        if workload == "windowed":
            return ctx.vacf_windowed()
        return ctx.helfand(ana._volumes, ana.boltzmann, ana.temp_avg, fft=(workload == "helfand"))

    need_flush = A * T * 24 <= 2 * 126e6      # inputs do not exceed L2: flush it between steps
    for _ in range(args.warmup):
        compute()
    launches0 = ctx.launch_count()
    dev_ms, k_ms = [], []
    with sampler:
        for _ in range(args.steps):
            if need_flush:
                ctx.flush_l2()
            barrier()
            ctx.timer_begin()
            ts = compute()
            ms = ctx.timer_end()
            barrier()
            dev_ms.append(max_over_ranks(ms))
            k_ms.append(ctx.last_kernel_ms())
    launches = ctx.launch_count() - launches0
    # same per-particle values; the particle sum is associated per launch (one launch here, one per staging chunk in run())
    assert np.allclose(ts, ts_e2e, rtol=1e-12, atol=1e-14 * np.abs(ts_e2e).max()), "device-resident and end-to-end results differ"

    ms_step = float(np.mean(dev_ms))
    af_total = float(A) * T * world
    value = af_total / (ms_step / 1e3)
    e2e_ms = float(np.mean(e2e_times)) * 1e3
    e2e_value = af_total / (e2e_ms / 1e3)
    kernel_ms = float(np.mean(k_ms))

    if rank == 0:
        peaks, peak_kind = load_peaks()
        bytes_launch = BYTES_PER_AF[workload] * A * T
        achieved = bytes_launch / (kernel_ms / 1e3) / 1e9
        fl = flops_per_af(workload, T) * A * T
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
                "traffic": load_profile_traffic(workload),
                "kernel": {"fft": k1_kernel_name(ctx.fft_plan_info()) if workload == "fft" else "",
                           "windowed": "k_windowed<double,PRODUCT> (K2)",
                           "helfand": "k1f_fft_acf (K1; K5 flags and K6 exact refinement follow)",
                           "helfand_direct": "k_windowed<double,SQDIFF> (K3)"}[workload],
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": bytes_launch,
                "kernel_share_of_step": kernel_ms / ms_step,
                "fp64": {"algorithmic_flop_per_launch": fl, "achieved_tflops": fl / (kernel_ms / 1e3) / 1e12,
                         "peak_tflops_nominal": FP64_NOMINAL_TFLOPS,
                         "frac": fl / (kernel_ms / 1e3) / 1e12 / FP64_NOMINAL_TFLOPS,
                         "note": "this kernel is FP64-pipe / shared-memory bound, not HBM bound (SURVEY.md 8d)"}}
        cpu = cpu_baseline(workload, T)
        line = {
            "metric": "atom-frames/s", "value": value, "unit": "atom-frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(workload, A, T, world), host_numa_binding=numa),
            "e2e": {"value": e2e_value, "unit": "atom-frames/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": {"fft": "VelocityAutocorr(ag, fft=True).run()", "windowed": "VelocityAutocorr(ag, fft=False).run()",
                            "helfand": "ViscosityHelfand(ag).run()", "helfand_direct": "ViscosityHelfand(ag, fft=False).run()"}[workload]},
            "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "clocks": sampler.summary(),
            "fft_plan": ctx.fft_plan_info() if workload == "fft" else None,
            "setup_s": setup_s,
        }
        emit(json.dumps(line))
    _lib.host_unregister(vel)
    if pos is not None:
        _lib.host_unregister(pos)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(workload, T):
    """Oracle (numpy restatement of the reference) on one host core, bounded sample."""
    per = cpu_sample_size(workload, T) * 16
    dt = _ref_worker((workload, T, per, 7))
    sample = f"{per} atoms x {T} frames, 1 process (numpy elementwise ops and pocketfft are single-threaded)"
    if workload in ("helfand", "helfand_direct"):
        sample += "; 5 sampled lags extrapolated by sum(T-lag)"
    return {"value": per * T / dt, "unit": "atom-frames/s", "cores": 1, "kind": "port", "sample": sample}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fft", choices=["fft", "windowed", "helfand", "helfand_direct"])
    ap.add_argument("--atoms", type=int, default=0, help="atoms per GPU (default: BASELINE config)")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="do not pin the rank to the CPUs of its GPU's NUMA node before allocating the host trajectory")
    ap.add_argument("--frames", type=int, default=0)
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # launched without torchrun: re-exec under it so that one process drives one GPU
            port = os.environ.get("MASTER_PORT", "29511")
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", port, os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
