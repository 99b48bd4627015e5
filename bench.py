#!/usr/bin/env python
"""PVDER-v0 hot-path benchmark (driver contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is ONE env step of every environment of the batch: 2n = 30 half-cycle (1/120 s) ROS4-L
(Rosenbrock) sub-steps per env, fused with action, events/RNG, reward, observation and done in one kernel
launch (reference gym_PVDER/envs/PVDER_env.py:138-196).  Workload at N GPUs: 1,048,576
single-phase envs PER GPU (BASELINE.json metric: "env-steps/sec ... (1M envs)"), sharded by global
env index, no collective on the data path ("scaling": "weak").

  value        env-steps/s with state, actions and outputs resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the host-buffer C ABI call a Gym user makes
               (pvder_env_step_host: pinned numpy action H2D -> kernel -> obs/reward/done D2H, every step)
  roofline     FP64-compute bound: achieved = sub-steps/s x F, F = 2.2 kflop
               (1-ph) / 12.5 kflop (3-ph) per sub-step -- the survey's yardstick, which prices a DENSE LU
               and three right-hand sides per sub-step; the kernels exploit the sparsity, so frac can exceed 1
               and "executed" (ncu-counted FP64 flops of the committed kernel x sub-steps/s) is the hardware
               utilisation; peak = FP64 FMA peak measured live by the K0 micro-benchmark
               (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline the restated reference path (oracle O2: scipy LSODA with the reference's settings,
               Python RHS/Jacobian callbacks) timed on this box's host cores on a bounded sample
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on the first
# communicator), so file descriptor 1 points at stderr for the whole run and the line goes to the saved descriptor.
_REAL_STDOUT = None


def _stdout_to_stderr():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    out = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    os.write(out, (json.dumps(line) + "\n").encode())

# algorithmic flop per half-cycle sub-step: n = 11 -> 2.2 kflop, n = 23 -> 12.5 kflop.  The
# balanced three-phase reduction integrates 11 states, so it is measured with the n = 11 yardstick.
F_ALGO = {"model_1": 2.2e3, "model_2": 12.5e3, "model_2_balanced": 2.2e3}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: oracle O2 on the host cores
# ------------------------------------------------------------------------------------------------
REF_ENV_STEPS_PER_STEP = 64   # reference arm: one bench "step" = this many env steps on every host core


def _cpu_worker(args):
    model_type, n_sim, steps, warmup, seed = args
    import random
    import warnings

    warnings.filterwarnings("ignore")
    from oracle.env_oracle import OraclePVDEREnv

    env = OraclePVDEREnv(model_type=model_type, n_sim_time_steps_per_env_step=n_sim, solver="reference",
                         DISCRETE_REWARD=False, seed=seed)
    rng = random.Random(seed)
    env.reset()
    done_steps = 0
    failures = 0

    def one_step():
        # the reference asserts on a failed LSODA call (PVDER_env.py:177); with random actions the restated
        # path occasionally hits that deep in anti-windup operation -- count it and start a new episode
        nonlocal failures
        try:
            _, _, d, _ = env.step(rng.randrange(5))
        except AssertionError:
            failures += 1
            d = True
        if d:
            env.reset()

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
        done_steps += 1
    return time.perf_counter() - t0, done_steps, failures


def cpu_reference(model_type, n_sim, steps, warmup, envs_per_core=1):
    """All host cores, one oracle env per worker process; returns (env_steps_per_s, cores, sample)."""
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    jobs = [(model_type, n_sim, steps, warmup, 1000 + i) for i in range(cores * envs_per_core)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[1] for r in res)
    busy = max(r[0] for r in res) * envs_per_core
    sample = (f"{len(jobs)} oracle envs ({model_type}, n={n_sim}, LSODA rtol=atol=1e-4 hmax=1/120, random actions) x "
              f"{steps} env steps on {cores} processes; slowest worker {busy:.2f} s, wall incl. spawn {wall:.2f} s; "
              f"{sum(r[2] for r in res)} solver failures (episode restarted)")
    return total / busy, cores, sample


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
                pw.append(float(p[2]))
            except ValueError:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(local_rank):
    """Best effort: run this rank (and first-touch its pinned host buffers) on the CPUs of the NUMA node its GPU hangs
    off, so that N ranks do not push their device->host copies through one socket.  Returns a short description."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{dev}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return f"{dev}: no NUMA information"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"{dev}: node {node} has no allowed CPUs"
        os.sched_setaffinity(0, cpus)
        return f"{dev}: node {node}, {len(cpus)} cpus"
    except Exception as exc:      # sysfs layout, permissions, old torch ...
        return f"unavailable ({type(exc).__name__})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=160)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="model_1", choices=["model_1", "model_2"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--n-sim", type=int, default=15)
    ap.add_argument("--three-phase-mode", default="auto", choices=["auto", "balanced", "general", "split"],
                    help="model_2 only: auto/balanced reduction on phase a (default), general 23-state integration "
                         "with one thread per env, or split = general with three lanes per env")
    ap.add_argument("--grid-unbalance", type=float, nargs=2, default=(1.0, 1.0), metavar=("RB", "RC"),
                    help="model_2 general/split only: grid magnitude of phases b, c relative to phase a")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (0: min(steps, 40))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cfg", action="append", default=[], metavar="KEY=VALUE",
                    help="EnvConfig override for kernel studies (e.g. --cfg refine_input_level=0); recorded in config")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric, unit = "env-steps/sec", "env-steps/s"
    workload = (f"{args.envs_per_gpu} PVDER-v0 envs per GPU ({args.model}: "
                f"{'single-phase derId 10, 11 states' if args.model == 'model_1' else 'three-phase derId 50, 23 states, ' + args.three_phase_mode + ' integration'}), "
                f"n_sim_time_steps_per_env_step={args.n_sim} ({2 * args.n_sim} half-cycle sub-steps per env step), continuous reward, "
                f"default voltage events (Philox per-env streams), random actions, auto-reset at 40 s (160 steps/episode)")
    config = {"workload": workload, "envs_per_gpu": args.envs_per_gpu, "total_envs": args.envs_per_gpu * world,
              "sub_steps_per_env_step": 2 * args.n_sim, "sharding": f"env-index x{world}, no data-path collective",
              "l2_policy": "state+outputs per step (>=270 MB at 1M envs) exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        # a "step" of this arm is a bounded sample of the workload: REF_ENV_STEPS_PER_STEP env steps on every core
        v, cores, sample = cpu_reference(args.model, args.n_sim, args.steps * REF_ENV_STEPS_PER_STEP,
                                         args.warmup * REF_ENV_STEPS_PER_STEP)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * cores * REF_ENV_STEPS_PER_STEP / v, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                                 "note": "restated reference path (pvder unavailable): oracle O2"},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import gym_pvder_b200 as G
    from gym_pvder_b200 import _cabi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    config["host_affinity"] = bind_to_gpu_numa_node(local_rank) if world > 1 else "not set (single rank)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    n = args.envs_per_gpu
    K, Wm = args.steps, args.warmup
    overrides = {}
    for kv in args.cfg:
        key, _, val = kv.partition("=")
        overrides[key] = json.loads(val)
    if overrides:
        config["config_overrides"] = overrides
    cfg = G.EnvConfig(model_type=args.model, n_sim_time_steps_per_env_step=args.n_sim, max_sim_time=40.0,
                      DISCRETE_REWARD=False, goals_list=["voltage_regulation"], event_mode="philox", seed=2026,
                      auto_reset=True, balanced_three_phase=args.three_phase_mode, grid_unbalance_ratio=tuple(args.grid_unbalance),
                      **overrides)
    fkey = "model_2_balanced" if (args.model == "model_2" and args.three_phase_mode in ("auto", "balanced")) else args.model
    env = G.PVDERVecEnv(n, device=dev, env_offset=rank * n, config=cfg)
    env.reset()
    acts = torch.empty((K + Wm, n), dtype=torch.int32, device=dev)
    for s in range(K + Wm):
        _cabi.check(lib.pvder_sample_actions(cfg.c.seed, s, C.c_void_p(acts[s].data_ptr()), n, rank * n,
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for s in range(Wm):
        env.step(acts[s])
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for s in range(K):
        env.step(acts[Wm + s])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    env.check_status()
    stats = env.stats().clone()
    if world > 1:
        dist.all_reduce(stats)                       # optional episode-statistics reduction (NCCL), off the timed path
    value = world * n * K / (ms_max * 1e-3)
    kernel_ms = ms / K

    # ---- e2e leg: host buffers through the handle C ABI ----------------------------------------------
    Ke = args.e2e_steps or min(K, 40)
    h = C.c_void_p()
    _cabi.check(lib.pvder_env_create(C.byref(cfg.c), n, rank * n, C.byref(h)))
    n_act = min(8, K + Wm)
    h_act = torch.empty((n_act, n), dtype=torch.int32).pin_memory()
    h_act.copy_(acts[:n_act].cpu())
    h_obs = torch.empty((n, 11), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float64).pin_memory()
    h_done = torch.empty(n, dtype=torch.uint8).pin_memory()
    _cabi.check(lib.pvder_env_reset_host(h, C.c_void_p(h_obs.data_ptr()), None))

    def host_step(s):
        _cabi.check(lib.pvder_env_step_host(h, C.c_void_p(h_act[s % n_act].data_ptr()), C.c_void_p(h_obs.data_ptr()), None,
                                            C.c_void_p(h_rew.data_ptr()), C.c_void_p(h_done.data_ptr())))

    We = max(3, Wm)      # warm-up: also lets the handle's measured copy/kernel ratio (chunk schedule) settle
    for s in range(We):
        host_step(s)
    _cabi.check(lib.pvder_env_kernel_ms(h, None, None))      # kernel-span counters restart with the timed region
    barrier()
    t0 = time.perf_counter()
    for s in range(Ke):
        host_step(We + s)
    barrier()
    wall = time.perf_counter() - t0
    tw = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    e2e_value = world * n * Ke / float(tw.item())
    kms, kcnt = C.c_double(), C.c_int64()
    _cabi.check(lib.pvder_env_kernel_ms(h, C.byref(kms), C.byref(kcnt)))
    e2e_checksum = float(h_rew.sum())
    pchunks, pratio = C.c_int32(), C.c_double()
    _cabi.check(lib.pvder_env_pipeline_info(h, C.byref(pchunks), C.byref(pratio)))
    _cabi.check(lib.pvder_env_destroy(h))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: FP64 FMA peak measured live (K0) ----------------------------------------------------
    tf, pms = C.c_double(), C.c_double()
    _cabi.check(lib.pvder_fp64_peak(4000, C.byref(tf), C.byref(pms)))
    sub_per_launch = n * 2 * args.n_sim
    achieved = sub_per_launch * F_ALGO[fkey] / (kernel_ms * 1e-3) * 1e-12
    roofline = {"bound": "fp64", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved / tf.value,
                "traffic": None, "kernel": ("pvder::step_kernel_split3" if (args.model == "model_2" and args.three_phase_mode == "split") else
                           "pvder::step_kernel<%s>" % {"model_1": "Model1ph", "model_2": "Model3ph", "model_2_balanced": "Model3phBal" if args.three_phase_mode == "balanced" else "Model3ph,auto"}[fkey]),
                "kernel_ms": kernel_ms, "flop_per_sub_step": F_ALGO[fkey],
                "peak_source": "FP64 FMA micro-benchmark pvder_fp64_peak run in this process (MEASURED_PEAKS.json has no FP64 entry)",
                "hbm_algorithmic_bytes_per_launch": n * (2 * 8 * _cabi.sd_fields(cfg.n_state) + 4 * 8 + 4 + 44 + 8 + 1),
                "note": "SURVEY.md 8d: the path is FP64-compute bound, not HBM/tensor; the contract's enum has no fp64 "
                        "member so the bound is named explicitly"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
        hb = roofline["hbm_algorithmic_bytes_per_launch"] / (kernel_ms * 1e-3) * 1e-9
        roofline["hbm_gbs_achieved"] = hb
        roofline["hbm_frac_of_measured"] = hb / peaks["hbm_gbs"]
    except Exception:
        pass

    try:   # DRAM bytes of one launch from the committed ncu --set full capture (not measured in this run)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tr = json.load(fh)
        tkey = args.model if args.model == "model_1" else "model_2_" + args.three_phase_mode   # absent key -> traffic stays null
        roofline["traffic"] = tr[tkey]["bytes"] * (n / tr["envs"])
        roofline["traffic_source"] = tr[tkey]["source"] + " (ncu dram__bytes_read+write per launch, scaled by envs)"
    except Exception:
        pass

    try:   # executed FP64 flops per sub-step of the committed kernel (ncu instruction counters, profiles/)
        with open(os.path.join(ROOT, "profiles", "ncu_flops.json")) as fh:
            fl = json.load(fh)
        tkey = args.model if args.model == "model_1" else "model_2_" + args.three_phase_mode
        ex = fl[tkey]["flop_per_sub_step"] * sub_per_launch / (kernel_ms * 1e-3) * 1e-12
        roofline["executed"] = {"flop_per_sub_step": fl[tkey]["flop_per_sub_step"], "tflops": ex, "frac": ex / tf.value,
                                "source": fl[tkey]["source"]}
        roofline["note"] += ("; 'achieved'/'frac' use the survey's yardstick F (dense LU + 3 right-hand sides per sub-step), "
                             "'executed' counts the FP64 flops the kernel really issues (2 per DFMA, 1 per DADD/DMUL)")
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_reference(args.model, args.n_sim, 160 * 30, 2, envs_per_core=1)   # ~10-20 s of CPU work
        cpu = {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
               "note": "restated reference path (pvder unavailable): oracle O2"}

    st = stats.cpu().numpy()
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "sub_steps_per_sec": value * 2 * args.n_sim,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": (44 + 8 + 1) * n,
                    "steps": Ke, "kernel_ms_in_e2e": kms.value / max(1, kcnt.value), "reward_checksum": e2e_checksum,
                    "chunks": pchunks.value, "copy_to_kernel_time_ratio": pratio.value},
            "gpu_launches": K, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "episode_stats": {"envs": st[10], "windup_sub_steps": st[9], "exact_sub_steps": st[11], "failed": st[3],
                              "note": "counters of the episode in progress at the end of the run (auto-reset clears them)"}}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    _stdout_to_stderr()
    main()
