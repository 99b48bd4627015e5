"""Generates tests/golden/golden_{model_1,model_2}.npz from the CPU oracle's tight-tolerance LSODA
path (oracle/env_oracle.py, solver="tight").

The reference itself cannot be imported in this image (its simulator dependency `pvder` is not
vendored/installed), so these vectors are produced by the RESTATED reference path,
not by the reference: parity stays "unpinned" in the sense of DESIGN.md.  They freeze the oracle's
answers so that GPU tests do not depend on scipy's LSODA build and run in milliseconds.

    python tests/golden/make_golden.py
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gym_pvder_b200 as G  # noqa: E402  (host-side config only: event grid indices)
import helpers as H  # noqa: E402
from oracle.env_oracle import OraclePVDEREnv  # noqa: E402

N_ENVS, N_STEPS = 6, 16


def main():
    for model_type in ("model_1", "model_2"):
        cfg = G.EnvConfig(model_type=model_type, events_spec=H.SAG_SPEC, event_mode="table")
        ns = cfg.n_state
        acts = np.zeros((N_ENVS, N_STEPS), dtype=np.int32)
        obs = np.zeros((N_ENVS, N_STEPS, 11))
        rew = np.zeros((N_ENVS, N_STEPS))
        state = np.zeros((N_ENVS, N_STEPS, ns))
        vt, st = [], []
        for i in range(N_ENVS):
            ev = H.random_events(1000 + i)
            v, s = H.oracle_tables(ev, cfg.c)
            vt.append(v)
            st.append(s)
            rng = random.Random(50 + i)
            env = OraclePVDEREnv(model_type=model_type, solver="tight", events=ev, DISCRETE_REWARD=True)
            env.reset()
            for k in range(N_STEPS):
                a = 0 if i == 0 else rng.randrange(5)
                acts[i, k] = a
                o, r, d, _ = env.step(a)
                obs[i, k] = o
                rew[i, k] = r
                state[i, k] = H.oracle_delta_state(env)
            print(model_type, "env", i, "done; windup sub-steps:", env.windup_substeps, flush=True)
        out = os.path.join(ROOT, "tests", "golden", f"golden_{model_type}.npz")
        np.savez_compressed(out, actions=acts, obs=obs, reward=rew, state=state,
                            vgrid_tab=np.concatenate(vt, axis=1), sinsol_tab=np.concatenate(st, axis=1))
        print("wrote", out)


if __name__ == "__main__":
    main()
