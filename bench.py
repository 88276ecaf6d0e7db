#!/usr/bin/env python
"""Benchmark of the TVO hot path (BASELINE.json: E-step datapoints/s + EM-iteration time,
BSC H=256 D=144 S=64, EVO randparents + sparseflip P=16 C=2 G=2, fp64, N = 1M).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n-total N] [--impl reference]

One "step" = one full EM iteration over the dataset: fused EVO E-step, M-step accumulation (with the free
energy), packed all-reduce + parameter update.  The dataset (N = 1M datapoints, the north star's size) is SHARDED
over the ranks (`"scaling": "strong"`); `value` is the whole-job E-step throughput (datapoints/s, device-resident
inputs, CUDA events, max over ranks); `ms_per_step` / `em_iter_ms` is the EM-iteration time.  For N > 1 the line
also carries `weak` (1M datapoints per GPU).  `e2e` is the E-step throughput through the public API with the batch
coming from pinned HOST memory and the substitution count read back every step.  At N = 1 the line also carries
`other_configs` (short runs of BASELINE configs[2..4]) and `cpu_baseline`.

`--impl reference` times the UNMODIFIED reference (baseline/_ref: `pip install --target` of /root/reference, done by
__graft_entry__.build()) on the box's host cores through its own public API (EVOVariationalStates.update timed
inside Trainer.em_step) on a bounded sample of the same workload; rank 0 only.  Without baseline/_ref it falls back
to the oracle's numpy restatement (`kind: "port"`).
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


# ------------------------------------------------------------------------------------ CPU arms
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_installed():
    return os.path.isdir(os.path.join(REF_DIR, "tvo", "variational"))


def _import_reference():
    """The unmodified reference from baseline/_ref.  h5py / munch are imported by it at module level
    (tvo/utils/__init__.py:2, tvo/utils/parallel.py:7, tvo/exp/_experiments.py:30) but are absent from the image and
    unused on this path: empty stand-in modules satisfy the imports."""
    import types
    for name in ("h5py", "munch"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                stub = types.ModuleType(name)
                if name == "munch":
                    class Munch(dict):
                        __getattr__ = dict.__getitem__
                        __setattr__ = dict.__setitem__
                    stub.Munch = Munch
                else:
                    stub.File = stub.Dataset = stub.Group = object
                sys.modules[name] = stub
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import tvo
    assert os.path.abspath(tvo.__file__).startswith(REF_DIR), "a different `tvo` package shadows baseline/_ref"
    return tvo


def _ref_worker(args):
    """One process of the reference's CPU path: Trainer.em_step over its own shard (run policy 'seq'), the E-step
    timed around the reference's own EVOVariationalStates.update calls."""
    seed, n, batch, threads, reps = args
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["MKL_NUM_THREADS"] = str(threads)
    import numpy as np
    import torch as to
    to.set_num_threads(threads)
    _import_reference()
    from tvo.models import BSC
    from tvo.trainer import Trainer
    from tvo.utils.data import TVODataLoader
    from tvo.variational import EVOVariationalStates
    to.manual_seed(seed)
    np.random.seed(seed)
    Wgen = to.randn(D, H, dtype=to.float64)
    Y = (to.rand(n, H) < PI_GEN).double() @ Wgen.t() + to.randn(n, D, dtype=to.float64)
    model = BSC(H, D, Wgen + 0.5 * to.randn(D, H, dtype=to.float64), to.tensor([1.0], dtype=to.float64),
                to.full((H,), 1.0 / H, dtype=to.float64), precision=to.float64)
    states = EVOVariationalStates(N=n, H=H, S=S, precision=to.float64, parent_selection="randparents",
                                  mutation="sparseflip", n_parents=P, n_generations=G, n_children=C,
                                  bitflip_frequency=1.0 / H)
    trainer = Trainer(model, TVODataLoader(Y, batch_size=min(batch, n)), states)
    t_update = [0.0]
    orig_update = states.update  # the reference's method, wrapped on the instance only to time it

    def timed_update(*a, **k):
        t0 = time.perf_counter()
        r = orig_update(*a, **k)
        t_update[0] += time.perf_counter() - t0
        return r

    states.update = timed_update
    te = tem = 0.0
    for rep in range(reps + 1):  # the first EM iteration is warm-up
        t_update[0] = 0.0
        t0 = time.perf_counter()
        trainer.em_step()
        dt = time.perf_counter() - t0
        if rep > 0:
            te += t_update[0]
            tem += dt
    return te / reps, tem / reps


def reference_baseline(n_total=8192, reps=1, batch=1024, cores=None):
    """The unmodified reference on the host cores, both layouts of BASELINE.md section 3: (a) one process with all
    threads, (b) one single-threaded process per core, each on its own shard (the reference's data-parallel layout
    without the all-reduce, i.e. an upper bound of its MPI mode).  The faster one is reported."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    res = {}
    with ctx.Pool(1) as pool:
        res["1 process x %d threads" % cores] = (pool.map(_ref_worker, [(1000, n_total, batch, cores, reps)]), n_total)
    n_each = max(64, n_total // cores)
    with ctx.Pool(cores) as pool:
        res["%d processes x 1 thread" % cores] = (pool.map(_ref_worker, [(1000 + r, n_each, batch, 1, reps)
                                                                         for r in range(cores)]), n_each * cores)
    wall = time.perf_counter() - t0
    best = None
    for layout, (rr, n) in res.items():
        te, tem = max(r[0] for r in rr), max(r[1] for r in rr)  # the slowest process gates the step
        cand = {"value": n / te, "em_iter_datapoints_per_s": n / tem, "estep_s": te, "em_iter_s": tem, "layout": layout,
                "n": n}
        if best is None or cand["value"] > best["value"]:
            best = cand
    others = {k: round(n / max(r[0] for r in rr), 1) for k, (rr, n) in res.items()}
    return {
        "value": best["value"], "unit": "datapoints/s", "cores": cores, "kind": "reference",
        "sample": f"unmodified tvlearn/tvo (baseline/_ref), {best['layout']}, {best['n']} datapoints in batches of "
                  f"{batch}, {reps} timed Trainer.em_step after 1 warm-up, E-step = time inside "
                  f"EVOVariationalStates.update; E-step datapoints/s per layout: {others}; wall {wall:.1f}s",
        "em_iter_datapoints_per_s": best["em_iter_datapoints_per_s"], "estep_s": best["estep_s"],
        "em_iter_s": best["em_iter_s"],
    }


def _port_worker(args):
    """One process = one rank of the reference's data-parallel CPU mode, oracle restatement (fallback arm)."""
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


def port_baseline(n_per_core=4096, reps=1, batch=128, cores=None):
    """Fallback when baseline/_ref is absent: the oracle's numpy restatement with the reference's RNG usage, `cores`
    processes x 1 thread (the reference's MPI layout)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_port_worker, [(1000 + r, n_per_core, reps, batch) for r in range(cores)])
    wall = time.perf_counter() - t0
    te = max(r[0] for r in res)
    tem = max(r[0] + r[1] for r in res)
    n_tot = n_per_core * cores
    return {
        "value": n_tot / te, "unit": "datapoints/s", "cores": cores, "kind": "port",
        "sample": f"{n_tot} datapoints ({cores} processes x {n_per_core}, batches of {batch}), {reps} timed E-step(s) "
                  f"after 1 warm-up; oracle numpy restatement with the reference's RNG usage (baseline/_ref absent); "
                  f"wall {wall:.1f}s",
        "em_iter_datapoints_per_s": n_tot / tem, "estep_s": te, "em_iter_s": tem,
    }


def cpu_baseline(args, reps=1):
    if reference_installed() and not args.port_baseline:
        return reference_baseline(n_total=args.cpu_n, reps=reps)
    return port_baseline(n_per_core=args.cpu_n_per_core, reps=reps)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    last = cpu_baseline(args, reps=args.steps_ref)  # steps_ref timed EM iterations after one warm-up iteration
    v = last["value"]
    line = {
        "impl": "reference", "metric": f"E-step datapoints/s (BSC H{H} D{D} S{S}); EM-iteration time in em_iter_ms",
        "value": v, "unit": "datapoints/s", "n_gpus": args.gpus, "steps": args.steps_ref, "warmup": 1,
        "ms_per_step": last["em_iter_s"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": last["sample"]},
        "cpu_baseline": {**{k: last[k] for k in ("unit", "cores", "kind", "sample")}, "value": v},
        "e2e": {"value": v, "unit": "datapoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "em_iter_datapoints_per_s": last["em_iter_datapoints_per_s"], "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def _time_em(to, dist, world, wl, steps, warmup, sampler=None):
    """`warmup` untimed + `steps` timed EM iterations of workload `wl` on this rank's shard.  Returns per-step E-step /
    EM-iteration ms (CUDA events on the launching stream) and the launch / kernel-timing counters."""
    from tvo_b200 import _lib
    model, states, Y = wl["model"], wl["states"], wl["Y"]
    idx = to.arange(Y.shape[0], device=Y.device)

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

    for _ in range(warmup):
        em_step()
    barrier()
    if sampler is not None:
        sampler.start()
    evs = [[to.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    _lib.reset_launch_count()
    _lib.kernel_timing_enable(True)  # cudaEvent pair around every launch of the fused E-step / M-step kernels
    barrier()
    t_wall = time.perf_counter()
    for k in range(steps):
        em_step(evs[k])
    barrier()
    t_wall = time.perf_counter() - t_wall
    out = {"launches": _lib.launch_count(), "wall_s": t_wall}
    out["k_ms"], out["k_n"] = _lib.kernel_timing_read(_lib.TIMED_ESTEP_EVO_BSC)
    out["m_ms"], out["m_n"] = _lib.kernel_timing_read(_lib.TIMED_BSC_MSTEP)
    _lib.kernel_timing_enable(False)
    out["clocks"] = sampler.stop() if sampler is not None else None
    out["estep_ms"] = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
    out["total_ms"] = evs[0][0].elapsed_time(evs[-1][2]) / steps
    tm = to.tensor([out["estep_ms"], out["total_ms"]], dtype=to.float64, device=Y.device)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    out["estep_ms"], out["total_ms"] = tm.tolist()
    return out


def _other_configs(to, dev, hbm_peak, peak_src, steps=3, warmup=2):
    """Short device-resident runs of BASELINE configs[2..4] on this GPU (driver-visible numbers for the other shapes):
    E-step datapoints/s, EM-iteration time and the HBM roofline of the E-step from SURVEY 8(d)'s bytes."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import workloads
    out = {}
    for key, make, N in (("configs[1] C2, converged model (workload realism: denser states)", workloads.make_c2_converged, 500_000),
                         ("configs[1] C2 in the stated fp32 mode (a = Y W on tcgen05 kind::tf32 + TMEM, 3xTF32 split)",
                          workloads.make_c2_fp32, 500_000),
                         ("configs[2] C3", workloads.make_c3, 200_000), ("configs[3] C4", workloads.make_c4, 500_000),
                         ("configs[4] C5 (one GPU's 1/8 share would be 125k)", workloads.make_c5, 100_000)):
        try:
            wl = make(dev, N)
            sampler = ClockSampler(dev.index or 0, None)
            r = _time_em(to, None, 1, wl, steps, 60 if "converged" in key else warmup, sampler)
            ach = wl["alg_bytes_per_dp"] * N / (r["estep_ms"] * 1e-3) / 1e9
            out[key] = {
                "workload": wl["name"], "N": N, "steps": steps, "warmup": warmup,
                "estep_datapoints_per_s": N / (r["estep_ms"] * 1e-3), "estep_ms": r["estep_ms"],
                "em_iter_ms": r["total_ms"], "em_iter_datapoints_per_s": N / (r["total_ms"] * 1e-3),
                "gpu_launches": r["launches"], "clocks": r["clocks"],
                "mean_active_units": workloads.mean_active_units(wl["states"]._Kp),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                             "traffic": None, "algorithmic_bytes_per_datapoint": wl["alg_bytes_per_dp"],
                             "peak_source": peak_src,
                             "what": "algorithmic E-step bytes (SURVEY 8d) / E-step time on the device (all kernels of "
                                     "the E-step, CUDA events)"},
            }
            del wl
            to.cuda.empty_cache()
        except Exception as exc:  # a failing side configuration must not take the headline line down
            out[key] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def run_gpu(args):
    import torch as to
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert to.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU fallback)"
    import tvo_b200
    from tvo_b200 import _lib
    from tvo_b200.utils import parallel
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import workloads

    if world > 1:
        parallel.init_processes()  # NCCL's own log (NCCL_DEBUG) is left alone; the JSON line is printed last
    else:
        to.cuda.set_device(local_rank)
        tvo_b200._set_device(to.device(f"cuda:{local_rank}"))
    dev = tvo_b200.get_device()
    N_total = args.n_total
    lo, hi = parallel.shard_bounds(N_total, world, rank)
    N = hi - lo
    make = workloads.make_c2 if HEADLINE else workloads.make_c5

    # ---- headline: N_total datapoints sharded over the ranks (strong scaling) -----------------------------------
    wl = make(dev, N, seed=1234, states_seed=2024, n_offset=lo, data_seed=99 + rank)
    try:
        uuid = "GPU-" + str(to.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid) if rank == 0 else None
    r = _time_em(to, dist, world, wl, args.steps, args.warmup, sampler)
    model, states, Y = wl["model"], wl["states"], wl["Y"]
    idx = to.arange(N, device=dev)
    F_per_dp = float(model._last_F.item()) / N_total
    subs = states._take_nsubs()
    mean_units = workloads.mean_active_units(states._Kp)

    # ---- e2e: E-step through the public API with the batch in pinned host memory --------------------------------
    e2e_ms = float("nan")
    if not args.no_e2e:
        Yh = to.empty((N, D), dtype=to.float64, pin_memory=True)
        Yh.copy_(Y)
        chunk = min(N, args.e2e_chunk)
        bufs = [to.empty((chunk, D), dtype=to.float64, device=dev) for _ in range(2)]
        copy_stream = to.cuda.Stream(device=dev)
        n_e2e = max(1, min(args.steps, 5))
        if world > 1:
            dist.barrier()
        t_e = []
        for k in range(n_e2e + 1):
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
            states._take_nsubs()  # D2H read of the step's result (one int64) -> sync
            t_e.append(time.perf_counter() - t0)
        tm = to.tensor([statistics.mean(t_e[1:]) * 1e3], dtype=to.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_ms = float(tm.item())
        del Yh, bufs

    # ---- weak scaling as an extra key (N > 1 only): 1M datapoints per GPU ----------------------------------------
    weak = None
    if world > 1 and not args.no_weak:
        del wl, model, states, Y, idx
        to.cuda.empty_cache()
        wlw = make(dev, args.n_weak, seed=1234, states_seed=2024, n_offset=rank * args.n_weak, data_seed=99 + rank)
        rw = _time_em(to, dist, world, wlw, max(2, min(args.steps, 5)), 3, None)
        weak = {"N_per_gpu": args.n_weak, "value": args.n_weak * world / (rw["estep_ms"] * 1e-3), "unit": "datapoints/s",
                "estep_ms": rw["estep_ms"], "em_iter_ms": rw["total_ms"],
                "em_iter_datapoints_per_s": args.n_weak * world / (rw["total_ms"] * 1e-3)}
        del wlw
        to.cuda.empty_cache()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s"
        clocks = r["clocks"]
        estep_ms, total_ms = r["estep_ms"], r["total_ms"]
        # dominant kernel = the fused EVO E-step kernel; its launches (one per chunk of <= 524288 datapoints) are timed
        # individually with CUDA events on the launching stream (tvo_kernel_timing_*)
        k_n = max(r["k_n"], 1)
        dp_per_launch = N * args.steps / k_n
        launch_ms = r["k_ms"] / k_n
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
                     "warp_instructions_per_datapoint": ipd, "source": tj.get("source")}
        except Exception:
            pass
        line = {
            "metric": f"E-step datapoints/s (BSC H{H} D{D} S{S}); EM-iteration time in em_iter_ms",
            "value": N_total / (estep_ms * 1e-3), "unit": "datapoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "N_total": N_total, "N_per_gpu": N, "batch": "whole shard per launch",
                       "l2": f"inputs (Y {N * D * 8 / 1e9:.2f} GB + K {N * S * ((H + 63) // 64) * 8 / 1e9:.2f} GB per GPU) "
                             "exceed the 126 MB L2",
                       "parallelism": f"datapoints sharded x{world} (N_total fixed), one packed NCCL all-reduce per EM "
                                      "iteration"},
            "estep_ms": estep_ms, "em_iter_ms": total_ms, "em_iter_datapoints_per_s": N_total / (total_ms * 1e-3),
            "free_energy_per_datapoint": F_per_dp, "subs_last_step_per_datapoint": subs / N / max(args.steps, 1),
            "mean_active_units": mean_units,
            "gpu_launches": r["launches"], "wall_s_timed_region": r["wall_s"],
            "clocks": clocks,
            "e2e": {"value": N_total / (e2e_ms * 1e-3), "unit": "datapoints/s", "h2d_bytes_per_step": N * D * 8,
                    "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms,
                    "h2d_gbs_per_gpu": N * D * 8 / (e2e_ms * 1e-3) / 1e9,
                    "what": "E-step via EVOVariationalStates.update_device on chunks copied from pinned host memory "
                            "(double-buffered copy stream), substitution count read back each step; bytes are per rank; "
                            "the leg is bound by the PCIe copy of Y (h2d_gbs_per_gpu), the E-step of all chunks but the "
                            "last is hidden behind it"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "estep_evo_bsc_v2_kernel<double,2,1,2>" if HEADLINE else "estep_evo_bsc_v2_kernel<double,4,4,4,unsigned short>",
                         "launch_ms": launch_ms, "launches_timed": r["k_n"], "datapoints_per_launch": dp_per_launch,
                         "algorithmic_bytes_per_datapoint": ALG_BYTES_PER_DP,
                         "traffic_source": "profiles/traffic.json (ncu --set full dram bytes per datapoint x datapoints per launch)",
                         "share_of_estep": r["k_ms"] / args.steps / estep_ms,
                         "mstep_kernel_ms_per_step": r["m_ms"] / args.steps, "issue_roofline": issue,
                         "note": ("the kernel's DRAM traffic equals its algorithmic bytes (+ the a = W^T y row), but it is "
                                  "bound by instruction issue (selection, mutation, dedup, sort on sparse bit states; see "
                                  "issue_roofline), not by HBM; DESIGN.md section 5") if HEADLINE
                         else "list-based kernel with two-byte positions (H > 256), d_h gathered from the a row; a = Y W by the library DGEMM at this H"},
        }
        if weak is not None:
            line["weak"] = weak
        if world == 1 and not args.no_other:
            del wl, model, states, Y, idx
            to.cuda.empty_cache()
            line["other_configs"] = _other_configs(to, dev, hbm_peak, peak_src)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(args).items()
                                    if k in ("value", "unit", "cores", "kind", "sample", "em_iter_datapoints_per_s")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n-total", "--n", dest="n_total", type=int, default=1_000_000,
                    help="datapoints of the whole job, sharded over the GPUs (strong scaling; the north star's N = 1M)")
    ap.add_argument("--n-weak", type=int, default=1_000_000, help="datapoints per GPU of the extra weak-scaling run")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-chunk", type=int, default=32768,
                    help="datapoints per host->device copy of the e2e leg (smaller: shorter un-overlapped tail)")
    ap.add_argument("--cpu-n", type=int, default=8192, help="datapoints of the reference arm's sample")
    ap.add_argument("--cpu-n-per-core", type=int, default=4096, help="datapoints per core of the port fallback")
    ap.add_argument("--port-baseline", action="store_true", help="time the oracle port even if baseline/_ref exists")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-fed leg (profiling runs)")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the short runs of the other BASELINE configurations")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2 = BASELINE configs[1] (the headline, default); c5 = configs[4] (H1024 D256 S128)")
    args = ap.parse_args()
    set_workload(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.steps_ref = max(1, min(args.steps, 3))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
