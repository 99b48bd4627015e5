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
  roofline     FP64-compute bound.  achieved/frac = the FP64 flops the kernel really EXECUTES (ncu
               instruction counters of the committed library, profiles/ncu_flops.json, per sub-step) x the sub-steps/s
               measured live, over the FP64 FMA peak measured live by the K0 micro-benchmark (MEASURED_PEAKS.json has
               no FP64 entry) -- reported only when the loaded library is the profiled one (sha256 check), else null.
               frac_yardstick keeps SURVEY 8d's algorithmic yardstick (2.2 / 12.5 kflop per sub-step: a DENSE LU +
               three right-hand sides, which the sparse kernels undercut -- it can exceed 1 and is not a hardware
               fraction)
  configs      the other BASELINE.json configurations, each a short device-timed leg: config 4 as written ("strong":
               1 Mi envs sharded over the N ranks, full 160-step episode), model_2 (the reference's reset() default) in
               auto and three-lane mode, config 3 (65,536 envs, sags + insolation, discrete reward), config 5 (262,144
               envs, fused Q-net policy in a CUDA graph), config 1 (one env of the restated reference path on ONE core)
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


def library_hashes(lib_path):
    """sha256 of the loaded shared library and of the SASS of its step kernels (cuobjdump, when available)."""
    import hashlib

    out = {"so_sha256": None, "sass_sha256": None}
    try:
        with open(lib_path, "rb") as fh:
            out["so_sha256"] = hashlib.sha256(fh.read()).hexdigest()
    except OSError:
        return out
    try:
        sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True, timeout=120).stdout
        keep, on = [], False
        for ln in sass.splitlines():
            if "Function :" in ln:
                on = "step_kernel" in ln
            if on:
                keep.append(ln.strip())
        if keep:
            out["sass_sha256"] = hashlib.sha256("\n".join(keep).encode()).hexdigest()
    except Exception:
        pass
    return out


def profile_matches(entry, hashes):
    """True when a profiles/ncu_*.json entry was captured from the library that is loaded now."""
    if entry.get("so_sha256") and entry["so_sha256"] == hashes.get("so_sha256"):
        return True
    return bool(entry.get("sass_sha256")) and entry["sass_sha256"] == hashes.get("sass_sha256")


def timed_steps(torch, env, acts, warm, steps):
    """Device time (CUDA events on the launching stream) of `steps` env steps after `warm` untimed ones, in ms."""
    for s in range(warm):
        env.step(acts[s % len(acts)])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for s in range(steps):
        env.step(acts[(warm + s) % len(acts)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def cpu_single_env_one_core(model_type, n_sim, steps=160):
    """BASELINE config 1: one env of the restated reference path (oracle O2), random agent seeded 0, one full episode on ONE
    core of this box (examples/gym_PVDER_environment_import_test.py:15-31)."""
    import random
    import warnings

    warnings.filterwarnings("ignore")
    from oracle.env_oracle import OraclePVDEREnv

    env = OraclePVDEREnv(model_type=model_type, n_sim_time_steps_per_env_step=n_sim, solver="reference", seed=0)
    rng = random.Random(0)
    env.reset()
    t0 = time.perf_counter()
    done_steps = 0
    failed = False
    for _ in range(steps):
        try:
            _, _, d, _ = env.step(rng.randrange(5))
        except AssertionError:
            failed = True
            break
        done_steps += 1
        if d:
            break
    dt = time.perf_counter() - t0
    return {"env_steps": done_steps, "ms_per_env_step": 1e3 * dt / max(1, done_steps), "env_steps_per_s": done_steps / dt,
            "sub_steps_per_s": done_steps * 2 * n_sim / dt, "cores": 1, "kind": "port", "solver_failed": failed}


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
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the legs of the other BASELINE configurations")
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
              "l2_policy": "state+outputs per step (>=270 MB at 1M envs) exceed the 126 MB L2; no flush needed",
              "host_affinity": "each rank bound to its GPU's NUMA node when the box exposes one (multi-rank runs), else unbound",
              "reference_arm_sample": f"the CPU arm times a bounded sample of this workload: {REF_ENV_STEPS_PER_STEP} env steps of "
                                      "one env per host core per bench step (1 Mi envs on the CPU path would take hours)"}

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
    host_affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else "not set (single rank)"
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
    # the same leg with the opt-in compact result formats (half obs, f32 reward, done bits: 26.1 instead of 53 B/env)
    h_obs16 = torch.empty((n, 11), dtype=torch.float16).pin_memory()
    h_rew32 = torch.empty(n, dtype=torch.float32).pin_memory()
    h_bits = torch.empty((n + 31) // 32, dtype=torch.int32).pin_memory()

    def host_step_compact(s):
        _cabi.check(lib.pvder_env_step_host_compact(h, C.c_void_p(h_act[s % n_act].data_ptr()), C.c_void_p(h_obs16.data_ptr()),
                                                    C.c_void_p(h_rew32.data_ptr()), C.c_void_p(h_bits.data_ptr())))

    for s in range(We):
        host_step_compact(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(Ke):
        host_step_compact(We + s)
    barrier()
    tw2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw2, op=dist.ReduceOp.MAX)
    e2e_compact = world * n * Ke / float(tw2.item())
    _cabi.check(lib.pvder_env_destroy(h))

    # ---- host-copy ceiling: all ranks copy one step's results (53 B/env) device -> pinned host at the same time -------
    # What the box's host side can take when N ranks pull their outputs concurrently; the e2e leg cannot beat
    # min(device rate, this ceiling).
    d2h_bytes = (44 + 8 + 1) * n
    d_src = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    h_dst = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    for _ in range(3):
        h_dst.copy_(d_src, non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    c0.record()
    for _ in range(reps):
        h_dst.copy_(d_src, non_blocking=True)
    c1.record()
    barrier()
    tc = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    copy_gbs = world * d2h_bytes * reps / (float(tc.item()) * 1e-3) * 1e-9        # aggregate over ranks
    copy_ceiling = copy_gbs * 1e9 / (44 + 8 + 1)                                   # env-steps/s the copies alone allow
    del d_src, h_dst

    # ---- config 4 as written: 1 Mi envs SHARDED over the ranks, full 160-step episode (strong scaling) --------------
    extra = {}
    total_strong = 1 << 20
    if not args.no_extra_configs and total_strong % world == 0:
        ns_ = total_strong // world
        cfg_s = G.EnvConfig(model_type="model_1", n_sim_time_steps_per_env_step=15, max_sim_time=40.0, DISCRETE_REWARD=False,
                            event_mode="philox", seed=2026, auto_reset=False)
        env_s = G.PVDERVecEnv(ns_, device=dev, env_offset=rank * ns_, config=cfg_s)
        env_s.reset()
        a_s = torch.empty((8, ns_), dtype=torch.int32, device=dev)
        for j in range(8):
            _cabi.check(lib.pvder_sample_actions(cfg_s.c.seed, j, C.c_void_p(a_s[j].data_ptr()), ns_, rank * ns_,
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for j in range(160):
            env_s.step(a_s[j % 8])
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        all_done = bool(env_s.done.all())
        env_s.check_status()
        st_s = env_s.stats().clone()
        if world > 1:
            dist.all_reduce(st_s)
        extra["strong_config4"] = {
            "workload": "BASELINE config 4: 1,048,576 single-phase envs sharded over the ranks, continuous reward, one full "
                        "160-step episode, auto-reset off", "total_envs": total_strong, "envs_per_gpu": ns_, "steps": 160,
            "value": total_strong * 160 / (float(ts.item()) * 1e-3), "unit": unit, "ms_per_step": float(ts.item()) / 160,
            "scaling": "strong", "all_done_at_step_160": all_done, "windup_sub_steps": float(st_s[9].item()),
            "waves_per_gpu": ns_ / (2 * 148 * 128)}
        del env_s, a_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the other configurations (rank 0, short legs; device-resident like `value`) ----------------------------------
    if not args.no_extra_configs:
        def leg(nenv, steps, warm, **kw):
            c_ = G.EnvConfig(n_sim_time_steps_per_env_step=15, max_sim_time=40.0, event_mode="philox", seed=2026,
                             auto_reset=True, **kw)
            e_ = G.PVDERVecEnv(nenv, device=dev, config=c_)
            e_.reset()
            a_ = torch.empty((8, nenv), dtype=torch.int32, device=dev)
            for j in range(8):
                _cabi.check(lib.pvder_sample_actions(c_.c.seed, j, C.c_void_p(a_[j].data_ptr()), nenv, 0,
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            ms_ = timed_steps(torch, e_, a_, warm, steps)
            e_.check_status()
            return {"envs": nenv, "steps": steps, "ms_per_step": ms_ / steps, "value": nenv * steps / (ms_ * 1e-3), "unit": unit}

        big = 1 << 20
        extra["model_2_auto"] = dict(leg(big, 20, 5, model_type="model_2", DISCRETE_REWARD=False, balanced_three_phase="auto"),
                                     workload="three-phase derId 50 (the reference's reset() default), balanced reduction + redo list")
        extra["model_2_split"] = dict(leg(big, 10, 3, model_type="model_2", DISCRETE_REWARD=False, balanced_three_phase="split"),
                                      workload="three-phase derId 50, general 23-state model, three lanes per env")
        extra["model_2_split_unbalanced_grid"] = dict(
            leg(big, 10, 3, model_type="model_2", DISCRETE_REWARD=False, balanced_three_phase="split",
                grid_unbalance_ratio=(0.95, 1.03)), workload="as model_2_split on a grid with phase magnitudes 1 : 0.95 : 1.03")
        sag = {"voltage": {"min": 0.90, "max": 1.02, "ENABLE": True}, "insolation": {"ENABLE": True}}
        extra["config3"] = dict(leg(65536, 40, 5, model_type="model_1", DISCRETE_REWARD=True, events_spec=sag),
                                workload="BASELINE config 3: 65,536 envs, voltage sags to 0.90 pu + insolation events, DISCRETE_REWARD")
        try:
            from gym_pvder_b200.rollout import DQNRollout

            c5 = G.EnvConfig(model_type="model_2", n_sim_time_steps_per_env_step=15, max_sim_time=40.0, DISCRETE_REWARD=True,
                             event_mode="philox", seed=1, auto_reset=True)
            v5 = G.PVDERVecEnv(262144, device=dev, config=c5)
            v5.reset()
            r5 = DQNRollout(v5, epsilon=0.1, use_cuda_graph=True, policy="fused").collect(40)
            extra["config5"] = {"envs": 262144, "steps": 40, "ms_per_step": r5["ms_per_iteration"], "value": r5["env_steps_per_s"],
                                "unit": unit, "workload": "BASELINE config 5: 262,144 envs (model_2), fused Q-net 11-100-5 + "
                                "epsilon-greedy + replay-ring writer, step -> collect captured in one CUDA graph"}
            del v5
        except Exception as exc:     # never lose the headline line to an auxiliary leg
            extra["config5"] = {"error": f"{type(exc).__name__}: {exc}"}
        if not args.no_cpu_baseline:
            extra["config1_cpu_1core"] = {
                "workload": "BASELINE config 1: single env, random agent, one 160-step episode of the restated reference path "
                            "(oracle O2, pvder unavailable) on ONE host core",
                "model_1": cpu_single_env_one_core("model_1", 15), "model_2": cpu_single_env_one_core("model_2", 15)}

    # ---- roofline: FP64 FMA peak measured live (K0) ----------------------------------------------------
    tf, pms = C.c_double(), C.c_double()
    _cabi.check(lib.pvder_fp64_peak(4000, C.byref(tf), C.byref(pms)))
    sub_per_launch = n * 2 * args.n_sim
    yard = sub_per_launch * F_ALGO[fkey] / (kernel_ms * 1e-3) * 1e-12
    kname = {"model_1": "pvder::step_kernel<Model1ph>", "model_2": "pvder::step_kernel<Model3ph>",
             "model_2_balanced": "pvder::step_kernel<Model3phBal%s>" % (", auto" if args.three_phase_mode == "auto" else "")}[fkey]
    if args.model == "model_2" and args.three_phase_mode == "split":
        kname = "pvder::step_kernel_split3"
    roofline = {"bound": "fp64", "achieved": None, "peak": tf.value, "unit": "TFLOP/s", "frac": None, "traffic": None,
                "kernel": kname, "kernel_ms": kernel_ms,
                "frac_yardstick": yard / tf.value, "achieved_yardstick": yard, "flop_per_sub_step_yardstick": F_ALGO[fkey],
                "peak_source": "FP64 FMA micro-benchmark pvder_fp64_peak run in this process (MEASURED_PEAKS.json has no FP64 entry; "
                               "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2)",
                "hbm_algorithmic_bytes_per_launch": n * (2 * 8 * _cabi.sd_fields(cfg.n_state) + 4 * 8 + 4 + 44 + 8 + 1),
                "note": "SURVEY.md 8d: the path is FP64-compute bound, not HBM/tensor; the contract's enum has no fp64 member so "
                        "the bound is named explicitly.  achieved/frac = FP64 flops the kernel EXECUTES (ncu instruction "
                        "counters of this very library: 2 per DFMA, 1 per DADD/DMUL) per second over the measured FMA peak; "
                        "frac_yardstick prices SURVEY 8d's dense-LU yardstick and is not a hardware fraction"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
        hb = roofline["hbm_algorithmic_bytes_per_launch"] / (kernel_ms * 1e-3) * 1e-9
        roofline["hbm_gbs_achieved"] = hb
        roofline["hbm_frac_of_measured"] = hb / peaks["hbm_gbs"]
    except Exception:
        pass

    hashes = library_hashes(_cabi.LIB_PATH)
    roofline["library"] = hashes
    tkey = args.model if args.model == "model_1" else "model_2_" + args.three_phase_mode
    try:   # executed FP64 flops per sub-step of THIS library (ncu instruction counters, profiles/): the roofline fraction
        with open(os.path.join(ROOT, "profiles", "ncu_flops.json")) as fh:
            fl = json.load(fh)
        if overrides or args.n_sim != fl.get("n_sim", 15):
            roofline["profile_check"] = "config overrides in force: the committed counters do not apply"
        elif not profile_matches(fl, hashes):
            roofline["profile_check"] = ("loaded library differs from the profiled one (profiles/ncu_flops.json: so %s...): "
                                         "executed flops / traffic not reported" % str(fl.get("so_sha256"))[:12])
        else:
            roofline["profile_check"] = "loaded library == profiled library (sha256)"
            ex = fl[tkey]["flop_per_sub_step"] * sub_per_launch / (kernel_ms * 1e-3) * 1e-12
            roofline["achieved"], roofline["frac"] = ex, ex / tf.value
            roofline["flop_per_sub_step"] = fl[tkey]["flop_per_sub_step"]
            roofline["fp64_inst_per_sub_step"] = fl[tkey].get("fp64_inst_per_sub_step")
            roofline["executed_source"] = fl[tkey]["source"]
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                tr = json.load(fh)
            if profile_matches(tr, hashes):
                roofline["traffic"] = tr[tkey]["bytes"] * (n / tr["envs"])
                roofline["traffic_source"] = tr[tkey]["source"] + " (ncu dram__bytes_read+write per launch, scaled by envs)"
    except Exception as exc:
        roofline.setdefault("profile_check", f"no committed ncu summary for this kernel ({type(exc).__name__})")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_reference(args.model, args.n_sim, 160 * 30, 2, envs_per_core=1)   # ~10-20 s of CPU work
        cpu = {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
               "note": "restated reference path (pvder unavailable): oracle O2"}

    st = stats.cpu().numpy()
    launches_per_step = 2 if (args.model == "model_2" and args.three_phase_mode == "auto") else 1
    device_rate = value
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "host_affinity_rank0": host_affinity,
            "sub_steps_per_sec": value * 2 * args.n_sim,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": (44 + 8 + 1) * n,
                    "steps": Ke, "kernel_ms_in_e2e": kms.value / max(1, kcnt.value), "reward_checksum": e2e_checksum,
                    "chunks": pchunks.value, "copy_to_kernel_time_ratio": pratio.value,
                    "host_copy_ceiling_gbs": copy_gbs, "host_copy_ceiling_env_steps_per_s": copy_ceiling,
                    "frac_of_min_device_rate_and_copy_ceiling": e2e_value / min(device_rate, copy_ceiling),
                    "compact": {"value": e2e_compact, "unit": unit, "d2h_bytes_per_step": 22 * n + 4 * n + 4 * ((n + 31) // 32),
                                "formats": "obs IEEE half, reward f32, done bit-packed (pvder_env_step_host_compact, opt-in)"},
                    "note": "host_copy_ceiling: all ranks copying one step's outputs (53 B/env) device -> pinned host at the "
                            "same time, nothing else running"},
            "gpu_launches": K * launches_per_step, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "configs": extra,
            "episode_stats": {"envs": st[10], "windup_sub_steps": st[9], "exact_sub_steps": st[11], "failed": st[3],
                              "note": "counters of the episode in progress at the end of the run (auto-reset clears them)"}}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    _stdout_to_stderr()
    main()
