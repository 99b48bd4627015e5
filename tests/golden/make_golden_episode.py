"""Generates tests/golden/golden_episode_{model_1,model_2}.npz: FULL 160-step episodes (40 s of simulated time, 4800
half-cycle sub-steps) from the CPU oracle's tight-tolerance LSODA path (oracle/env_oracle.py, solver="tight").

Trajectories (BASELINE.json config 2 and the north_star's "over a full episode"):
  0  events disabled, action 0 throughout                      (config 2: fixed actions, no events)
  1  events disabled, the cycle 1,1,2,0,3,4 repeated           (config 2, second schedule of SURVEY.md 8d)
  2  sags to 0.90 pu + insolation steps at t = 1..38 s, random actions, restricted to the actions that keep the DER
     below its current limit is NOT done: whatever regime the policy drives it into is in the fixture; `windup[k]` is
     the oracle's count of anti-windup sub-steps up to env step k (tests apply the windup tolerance from there on)

Like make_golden.py these come from the RESTATED reference path (pvder is unavailable): parity stays
"unpinned" in the sense of DESIGN.md.  Takes ~2-4 minutes per trajectory; the trajectories run in parallel.

    python tests/golden/make_golden_episode.py
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
        return [0] * N_STEPS
    if kind == 1:
        return [CYCLE[k % len(CYCLE)] for k in range(N_STEPS)]
    rng = random.Random(77)
    return [rng.randrange(5) for _ in range(N_STEPS)]


def run(job):
    model_type, kind = job
    import gym_pvder_b200 as G
    import helpers as H
    from oracle.env_oracle import EventTable, OraclePVDEREnv

    # one event grid for all three (38 instants); "events disabled" = a table that holds the defaults (1.0 pu, 100)
    cfg = G.EnvConfig(model_type=model_type, events_spec=H.SAG_SPEC, event_mode="table")
    ev = H.random_events(2000) if kind == 2 else EventTable()
    v, s = H.oracle_tables(ev, cfg.c)
    env = OraclePVDEREnv(model_type=model_type, solver="tight", events=ev, DISCRETE_REWARD=False)
    env.reset()
    acts = schedule(kind)
    obs = np.zeros((N_STEPS, 11))
    rew = np.zeros(N_STEPS)
    state = np.zeros((N_STEPS, cfg.n_state))
    windup = np.zeros(N_STEPS, dtype=np.int64)
    done = np.zeros(N_STEPS, dtype=bool)
    for k, a in enumerate(acts):
        o, r, d, _ = env.step(a)
        obs[k], rew[k], done[k] = o, r, d
        state[k] = H.oracle_delta_state(env)
        windup[k] = env.windup_substeps
    print(model_type, "trajectory", kind, "done; windup sub-steps:", env.windup_substeps, flush=True)
    return model_type, kind, np.array(acts, dtype=np.int32), obs, rew, state, windup, done, v, s


def main():
    jobs = [(m, k) for m in ("model_1", "model_2") for k in (0, 1, 2)]
    with mp.get_context("spawn").Pool(len(jobs)) as pool:
        res = pool.map(run, jobs)
    for model_type in ("model_1", "model_2"):
        rs = sorted([r for r in res if r[0] == model_type], key=lambda r: r[1])
        out = os.path.join(ROOT, "tests", "golden", f"golden_episode_{model_type}.npz")
        np.savez_compressed(out, actions=np.stack([r[2] for r in rs]), obs=np.stack([r[3] for r in rs]),
                            reward=np.stack([r[4] for r in rs]), state=np.stack([r[5] for r in rs]),
                            windup=np.stack([r[6] for r in rs]), done=np.stack([r[7] for r in rs]),
                            vgrid_tab=np.concatenate([r[8] for r in rs], axis=1),
                            sinsol_tab=np.concatenate([r[9] for r in rs], axis=1))
        print("wrote", out)


if __name__ == "__main__":
    main()
