"""Generates tests/golden/golden_continuous_clamp_model_1.npz: single-phase episodes in which the DER reaches its current
limit, integrated by the CPU oracle with the anti-windup clamp decided (nearly) continuously, as pvder does inside its
right-hand side: oracle/env_oracle.py solver="tight_continuous", clamp sampled 64 times per half-cycle
(the 16-sample solution differs from it by 4e-3 of the normal tolerance: converged; a clamp evaluated inside the RHS makes
a tight LSODA stall on the switching surface -- see that file's header).

The kernel, like the oracle's "tight" tier, samples the clamp once per half-cycle; these fixtures are what the
MEASURED gap between the two semantics is computed from and what the windup tolerance of the GPU/CPU tests is derived
from (DESIGN.md "Tolerances"), instead of the kernel being compared with an oracle that shares its sampling.

  0  events disabled, the +Q cycle 1,1,2,0,3,4 of BASELINE config 2 (limit reached at env step 103)
  1  sags + insolation events (seed 101), a +Q-biased random policy (limit reached at ~ env step 55)

Stores per env step: state (delta form), obs, reward, oracle windup count, and the same from the "tight" (half-cycle
sampled clamp) tier for the gap table.  Restated reference path (pvder unavailable): parity unpinned.
~15 minutes per trajectory, run in parallel.

    python tests/golden/make_golden_continuous.py
"""
import multiprocessing as mp
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_STEPS = 160
CYCLE = [1, 1, 2, 0, 3, 4]


def schedule(kind):
    if kind == 0:
        return [CYCLE[k % len(CYCLE)] for k in range(N_STEPS)]
    rng = random.Random(101)
    return [rng.choice([1, 1, 1, 0, 2, 3, 4]) for _ in range(N_STEPS)]


def run(job):
    kind, solver = job
    import gym_pvder_b200 as G
    import helpers as H
    from oracle.env_oracle import EventTable, OraclePVDEREnv

    cfg = G.EnvConfig(model_type="model_1", events_spec=H.SAG_SPEC, event_mode="table")
    ev = H.random_events(101) if kind == 1 else EventTable()
    v, s = H.oracle_tables(ev, cfg.c)
    env = OraclePVDEREnv(model_type="model_1", solver=solver, events=ev, DISCRETE_REWARD=False)
    env.reset()
    acts = schedule(kind)
    obs = np.zeros((N_STEPS, 11))
    rew = np.zeros(N_STEPS)
    state = np.zeros((N_STEPS, cfg.n_state))
    windup = np.zeros(N_STEPS, dtype=np.int64)
    for k, a in enumerate(acts):
        o, r, d, _ = env.step(a)
        obs[k], rew[k] = o, r
        state[k] = H.oracle_delta_state(env)
        windup[k] = env.windup_substeps
    return kind, solver, np.array(acts, dtype=np.int32), v[:, 0], s[:, 0], obs, rew, state, windup


if __name__ == "__main__":
    jobs = [(k, sv) for k in (0, 1) for sv in ("tight_continuous", "tight")]
    with mp.get_context("spawn").Pool(len(jobs)) as pool:
        res = {(k, sv): r for k, sv, *r in pool.map(run, jobs)}
    out = {}
    for key in ("actions", "vgrid_tab", "sinsol_tab"):
        i = {"actions": 0, "vgrid_tab": 1, "sinsol_tab": 2}[key]
        arr = np.stack([res[(k, "tight_continuous")][i] for k in (0, 1)])
        out[key] = arr if key == "actions" else arr.T          # tables: [ev_count, traj]
    for sv, tag in (("tight_continuous", ""), ("tight", "_sampled")):
        out["obs" + tag] = np.stack([res[(k, sv)][3] for k in (0, 1)])
        out["reward" + tag] = np.stack([res[(k, sv)][4] for k in (0, 1)])
        out["state" + tag] = np.stack([res[(k, sv)][5] for k in (0, 1)])
        out["windup" + tag] = np.stack([res[(k, sv)][6] for k in (0, 1)])
    path = os.path.join(ROOT, "tests", "golden", "golden_continuous_clamp_model_1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
