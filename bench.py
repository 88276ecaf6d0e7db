#!/usr/bin/env python
"""Benchmark of the TVO hot path (BASELINE.json: E-step datapoints/s + EM-iteration time,
BSC H=256 D=144 S=64, EVO randparents + sparseflip P=16 C=2 G=2, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n N_PER_GPU] [--impl reference]

One "step" = one full EM iteration over the rank's shard: fused EVO E-step, M-step accumulation
(with the free energy), packed all-reduce + parameter update.  `value` is the whole-job E-step
throughput (datapoints/s, device-resident inputs, CUDA events, max over ranks); `ms_per_step` /
`em_iter_ms` is the EM-iteration time.  `e2e` is the E-step throughput through the public API with the
batch coming from pinned HOST memory and the substitution count read back every step.

`--impl reference` times the reference algorithm's CPU path (the oracle's numpy restatement with
the reference's own RNG usage, one process per host core like the reference's MPI mode) on a
bounded sample of the same workload; rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, D, S = 256, 144, 64
P, C, G = 16, 2, 2
PI_GEN = 5.0 / H
WORKLOAD = "BSC H256 D144 S64 EVO(randparents,sparseflip,P16,C2,G2) fp64 — BASELINE configs[1] at the north-star N"
ALG_BYTES_PER_DP = D * 8 + 2 * S * ((H + 63) // 64) * 8 + S * 8          # 5760 B (fp64, packed)
ALG_FLOP_PER_DP = (S + P * C * G) * (2 * H * D + 3 * D + 2 * H)          # 9.56 MFLOP (dense count)
HEADLINE = True


def set_workload(name):
    """`--workload c5` re-points the module constants at BASELINE configs[4] (BSC H1024 D256 S128, the data-sharded
    multi-GPU configuration); the default is the headline configs[1].  Called before anything forks."""
    global H, D, S, P, C, G, PI_GEN, WORKLOAD, ALG_BYTES_PER_DP, ALG_FLOP_PER_DP, HEADLINE
    if name == "c5":
        H, D, S, P, C, G = 1024, 256, 128, 16, 4, 2
        PI_GEN = 8.0 / H
        WORKLOAD = "BSC H1024 D256 S128 EVO(randparents,sparseflip,P16,C4,G2) fp64 — BASELINE configs[4]"
        HEADLINE = False
    ALG_BYTES_PER_DP = D * 8 + 2 * S * ((H + 63) // 64) * 8 + S * 8
    ALG_FLOP_PER_DP = (S + P * C * G) * (2 * H * D + 3 * D + 2 * H)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML through
    pynvml in a thread (a sample every 5 ms — the timed region of the default run is a few hundred ms, shorter
    than nvidia-smi's start-up); `nvidia-smi -lms` is the fallback when pynvml is missing."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, uuid=None):
        self.index, self.uuid, self.rows, self.proc = index, uuid, [], None
        self.nvml, self.handle, self.sm, self.mask, self.max_mhz = None, None, [], 0, None
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            except Exception:  # no UUID from torch, or NVML does not know it: LOCAL_RANK is the NVML index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            self.mask |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))

    def _loop(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            self._stop.wait(0.005)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1.0)
            reasons = [k for k, b in self.BITS.items() if self.mask & b]
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml, 5 ms period, timed regions only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------ CPU baseline
def _cpu_worker(args):
    """One process = one rank of the reference's data-parallel CPU mode: E-step (+ M-step) on its chunk."""
    import numpy as np
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import evo as oevo, models as om, states as ost
    seed, n, reps, batch = args
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((D, H))
    Y = (rng.random((n, H)) < PI_GEN) @ W.T + rng.standard_normal((n, D))
    m = om.BSC(H, D, W + 0.5 * rng.standard_normal((D, H)), [1.0], np.full(H, 1.0 / H))
    K = np.tile(ost.generate_unique_states(S, H, rng=rng)[None], (n, 1, 1))
    lpj = np.empty((n, S))
    te = tm = 0.0
    for rep in range(reps + 1):  # first repetition is warm-up
        t_e = t_m = 0.0
        for b0 in range(0, n, batch):
            sl = slice(b0, min(n, b0 + batch))
            t0 = time.perf_counter()
            lpj[sl] = m.log_pseudo_joint(Y[sl], K[sl])
            Kb, lb = K[sl], lpj[sl]
            ns, nl = oevo.evolve_states_np(rng, lb, Kb, lambda s: m.log_pseudo_joint(Y[sl], s), P, C, G,
                                           "randparents", "sparseflip", float(m.theta["pies"].mean()), 1.0 / H)
            ost.update_states_for_batch(ns, nl, np.arange(Kb.shape[0]), Kb, lb)
            K[sl], lpj[sl] = Kb, lb
            t1 = time.perf_counter()
            m.update_param_batch(Y[sl], K[sl], lpj[sl])
            m.free_energy(Y[sl], K[sl], lpj[sl])
            t2 = time.perf_counter()
            t_e += t1 - t0
            t_m += t2 - t1
        m.update_param_epoch()
        if rep > 0:
            te += t_e
            tm += t_m
    return te / reps, tm / reps


def cpu_baseline(n_per_core=4096, reps=1, batch=128, cores=None):
    """Reference algorithm on the host cores: `cores` processes x 1 thread (the reference's MPI layout,
    its faster CPU mode per BASELINE.md), each on its own chunk.  Returns dict for the JSON line."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(1000 + r, n_per_core, reps, batch) for r in range(cores)])
    wall = time.perf_counter() - t0
    te = max(r[0] for r in res)  # slowest rank gates the step, as under MPI
    tem = max(r[0] + r[1] for r in res)
    n_tot = n_per_core * cores
    return {
        "value": n_tot / te, "unit": "datapoints/s", "cores": cores, "kind": "port",
        "sample": f"{n_tot} datapoints ({cores} processes x {n_per_core}, batches of {batch}), {reps} timed E-step(s) "
                  f"after 1 warm-up; oracle numpy restatement with the reference's RNG usage; wall {wall:.1f}s",
        "em_iter_datapoints_per_s": n_tot / tem, "estep_s": te, "em_iter_s": tem,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, ems, last = [], [], None
    for _ in range(args.warmup_ref + args.steps_ref):
        last = cpu_baseline(n_per_core=args.cpu_n_per_core, reps=1)
        vals.append(last["value"])
        ems.append(last["em_iter_s"])
    vals, ems = vals[args.warmup_ref:], ems[args.warmup_ref:]
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": f"E-step datapoints/s (BSC H{H} D{D} S{S}); EM-iteration time in em_iter_ms",
        "value": v, "unit": "datapoints/s", "n_gpus": args.gpus, "steps": args.steps_ref, "warmup": args.warmup_ref,
        "ms_per_step": statistics.mean(ems) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": last["sample"]},
        "cpu_baseline": {**{k: last[k] for k in ("unit", "cores", "kind", "sample")}, "value": v},
        "e2e": {"value": v, "unit": "datapoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "em_iter_datapoints_per_s": last["em_iter_datapoints_per_s"], "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import numpy as np
    import torch as to
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert to.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU fallback)"
    import tvo_b200
    from tvo_b200 import _lib
    from tvo_b200.models import BSC
    from tvo_b200.utils import parallel
    from tvo_b200.variational import EVOVariationalStates

    if world > 1:
        parallel.init_processes()  # NCCL's own log (NCCL_DEBUG) is left alone; the JSON line is printed last
    else:
        to.cuda.set_device(local_rank)
        tvo_b200._set_device(to.device(f"cuda:{local_rank}"))
    dev = tvo_b200.get_device()
    N = args.n
    N_total = N * world

    # synthetic data of the BASELINE shape, generated on the device; identical theta on all ranks
    g0 = to.Generator(device=dev).manual_seed(1234)
    Wgen = to.randn(D, H, dtype=to.float64, device=dev, generator=g0)
    W0 = Wgen + 0.5 * to.randn(D, H, dtype=to.float64, device=dev, generator=g0)
    g = to.Generator(device=dev).manual_seed(99 + rank)
    Y = to.empty(N, D, dtype=to.float64, device=dev)
    for b0 in range(0, N, 262144):
        b1 = min(N, b0 + 262144)
        s = (to.rand(b1 - b0, H, device=dev, generator=g) < PI_GEN).double()
        Y[b0:b1] = s @ Wgen.t() + to.randn(b1 - b0, D, dtype=to.float64, device=dev, generator=g)
    del s
    np.random.seed(0)  # K init: the reference's scheme, S unique Bernoulli(1/H) states for every datapoint
    model = BSC(H, D, W0, to.tensor([1.0], dtype=to.float64), to.full((H,), 1.0 / H, dtype=to.float64),
                precision=to.float64)
    states = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", n_parents=P, n_generations=G,
                                  n_children=C, bitflip_frequency=1.0 / H, seed=2024)
    states.n_offset = rank * N
    idx = to.arange(N, device=dev)

    def em_step(events=None):
        if events:
            events[0].record()
        states.update_device(idx, Y, model)
        if events:
            events[1].record()
        model.update_param_batch(idx, Y, states)
        model.update_param_epoch()
        if events:
            events[2].record()

    def barrier():
        if world > 1:
            dist.barrier()
        to.cuda.synchronize()

    for _ in range(args.warmup):
        em_step()
    barrier()
    try:
        uuid = "GPU-" + str(to.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    if rank == 0:
        sampler.start()
    evs = [[to.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    _lib.reset_launch_count()
    _lib.kernel_timing_enable(True)  # cudaEvent pair around every launch of the fused E-step / M-step kernels
    barrier()
    t_wall = time.perf_counter()
    for k in range(args.steps):
        em_step(evs[k])
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = _lib.launch_count()
    k_ms, k_n = _lib.kernel_timing_read(_lib.TIMED_ESTEP_EVO_BSC)
    m_ms, m_n = _lib.kernel_timing_read(_lib.TIMED_BSC_MSTEP)
    _lib.kernel_timing_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    estep_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    total_ms = evs[0][0].elapsed_time(evs[-1][2]) / args.steps
    F_per_dp = float(model._last_F.item()) / N_total
    subs = states._take_nsubs()

    # ---- e2e: E-step through the public API with the batch in pinned host memory ----------------
    Yh = to.empty((N, D), dtype=to.float64, pin_memory=True)
    Yh.copy_(Y)
    chunk = min(N, args.e2e_chunk)
    bufs = [to.empty((chunk, D), dtype=to.float64, device=dev) for _ in range(2)]
    copy_stream = to.cuda.Stream(device=dev)
    n_e2e = max(1, min(args.steps, 5))
    barrier()
    t_e = []
    for k in range(0 if args.no_e2e else n_e2e + 1):
        to.cuda.synchronize()
        t0 = time.perf_counter()
        ready = [to.cuda.Event() for _ in range(2)]
        done = [to.cuda.Event() for _ in range(2)]
        nb = (N + chunk - 1) // chunk
        for c in range(nb):
            b0, b1 = c * chunk, min(N, (c + 1) * chunk)
            buf = bufs[c % 2][: b1 - b0]
            with to.cuda.stream(copy_stream):
                if c >= 2:
                    copy_stream.wait_event(done[c % 2])
                buf.copy_(Yh[b0:b1], non_blocking=True)
                ready[c % 2] = to.cuda.Event()
                ready[c % 2].record(copy_stream)
            to.cuda.current_stream().wait_event(ready[c % 2])
            states.update_device(idx[b0:b1], buf, model)
            done[c % 2] = to.cuda.Event()
            done[c % 2].record()
        nsubs_host = states._take_nsubs()  # D2H read of the step's result (one int64) -> sync
        t_e.append(time.perf_counter() - t0)
    e2e_s = max(statistics.mean(t_e[1:]), 1e-9) if t_e else float("nan")

    # ---- reduce timings over ranks (max) ----------------------------------------------------------
    tm = to.tensor([estep_ms, total_ms, e2e_s * 1e3], dtype=to.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    estep_ms, total_ms, e2e_ms = tm.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel = the fused EVO E-step kernel; its launches (one per chunk of <= 131072 datapoints)
        # are timed individually with CUDA events on the launching stream (tvo_kernel_timing_*)
        dp_per_launch = N * args.steps / max(k_n, 1)
        launch_ms = k_ms / max(k_n, 1)
        achieved = ALG_BYTES_PER_DP * dp_per_launch / (launch_ms * 1e-3) / 1e9
        traffic, issue = None, None
        try:
            if not HEADLINE:
                raise KeyError("profiles/traffic.json holds the headline shape only")
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj["estep_evo_bsc_bytes_per_dp"] * dp_per_launch
            # second reading of the same launch: warp instructions issued per second against the scheduler peak
            # (SMs x 4 schedulers x SM clock); the kernel is bound here, not by HBM
            sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
            ipd = tj["estep_evo_bsc_warp_instructions_per_dp"]
            ach = ipd * dp_per_launch / (launch_ms * 1e-3)
            peak = 148 * 4 * sm_clock * 1e6
            issue = {"achieved_warp_inst_per_s": ach, "peak_warp_inst_per_s": peak, "frac": ach / peak,
                     "warp_instructions_per_datapoint": ipd, "source": "profiles/traffic.json (ncu smsp__inst_executed.sum)"}
        except Exception:
            pass
        line = {
            "metric": f"E-step datapoints/s (BSC H{H} D{D} S{S}); EM-iteration time in em_iter_ms",
            "value": N_total / (estep_ms * 1e-3), "unit": "datapoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "N_per_gpu": N, "N_total": N_total, "batch": "whole shard per launch",
                       "l2": f"inputs (Y {N * D * 8 / 1e9:.2f} GB + K {N * S * ((H + 63) // 64) * 8 / 1e9:.2f} GB per GPU) "
                             "exceed the 126 MB L2",
                       "parallelism": f"data-parallel x{world}, one packed NCCL all-reduce per EM iteration"},
            "estep_ms": estep_ms, "em_iter_ms": total_ms, "em_iter_datapoints_per_s": N_total / (total_ms * 1e-3),
            "free_energy_per_datapoint": F_per_dp, "subs_last_step_per_datapoint": subs / N / max(args.steps, 1),
            "gpu_launches": launches, "wall_s_timed_region": t_wall,
            "clocks": clocks,
            "e2e": {"value": N_total / (e2e_ms * 1e-3), "unit": "datapoints/s", "h2d_bytes_per_step": N * D * 8,
                    "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms,
                    "what": "E-step via EVOVariationalStates.update_device on chunks copied from pinned host memory "
                            "(double-buffered copy stream), substitution count read back each step"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "kernel": "estep_evo_bsc_gram_kernel<double>",
                         "launch_ms": launch_ms, "launches_timed": k_n, "datapoints_per_launch": dp_per_launch,
                         "algorithmic_bytes_per_datapoint": ALG_BYTES_PER_DP,
                         "traffic_source": "profiles/traffic.json (ncu --set full dram bytes per datapoint x datapoints per launch)",
                         "share_of_estep": k_ms / args.steps / estep_ms,
                         "dense_equivalent_tflops": ALG_FLOP_PER_DP * N / (estep_ms * 1e-3) / 1e12,
                         "mstep_kernel_ms_per_step": m_ms / args.steps, "issue_roofline": issue,
                         "note": ("the kernel's DRAM traffic equals its algorithmic bytes (+ the a = W^T y row), but it is "
                                  "bound by integer-ALU issue (selection, mutation, dedup, sort: ~7.7k warp instructions "
                                  "per datapoint, see issue_roofline), not by HBM; see DESIGN.md section 5") if HEADLINE
                         else "large-H*S variant of the kernel (one CTA per SM, latency-bound): profiles/r01j_c5.md"},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(n_per_core=args.cpu_n_per_core).items()
                                    if k in ("value", "unit", "cores", "kind", "sample", "em_iter_datapoints_per_s")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=1_000_000, help="datapoints per GPU (weak scaling)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-chunk", type=int, default=131072)
    ap.add_argument("--cpu-n-per-core", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-fed leg (profiling runs)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2 = BASELINE configs[1] (the headline, default); c5 = configs[4] (H1024 D256 S128)")
    args = ap.parse_args()
    set_workload(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.steps_ref = max(1, min(args.steps, 3))
    args.warmup_ref = 1
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
