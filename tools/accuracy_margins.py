"""Design study (not product, not test): worst error / tolerance of the integrator scheme on the golden
fixtures (tight-oracle trajectories with sags to 0.90 pu, insolation steps, random actions, windup), per state.
Runs the kernel source compiled as plain C++ (tests/host_emul) with -DPVDER_SCHEME=<n>.

    python tools/accuracy_margins.py [4|6] [model_1|model_2]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import emul_harness as E  # noqa: E402
import helpers as H  # noqa: E402

scheme = sys.argv[1] if len(sys.argv) > 1 else "4"
model_type = sys.argv[2] if len(sys.argv) > 2 else "model_1"
lib = f"/tmp/libpvder_emul_s{scheme}.so"
subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                f"-DPVDER_SCHEME={scheme}", "-o", lib, E._SRC], check=True)
E._lib = C.CDLL(lib)

gold = np.load(f"tests/golden/golden_{model_type}.npz")
acts = gold["actions"]
n, nsteps = acts.shape
em = E.EmulVecEnv(n, model_type=model_type, events_spec=H.SAG_SPEC, event_mode="table", DISCRETE_REWARD=True)
em.set_event_tables(gold["vgrid_tab"], gold["sinsol_tab"])
em.reset()
ns = em.ns
worst = np.zeros(ns)
worst_abs = np.zeros(ns)
worst_obs = np.zeros(11)
rew_mismatch = 0
B = 6 * em.cfg.phases
for s in range(nsteps):
    obs, rew, done, _ = em.step(acts[:, s])
    ref = gold["state"][:, s].T            # [ns, n]
    err = np.abs(em.sd[:ns, :n] - ref)
    tol = H.RTOL * np.abs(ref) + H.ATOL
    tol[B + 3] = 2e-4
    tol[B + 4] = 5e-6
    worst = np.maximum(worst, (err / tol).max(axis=1))
    worst_abs = np.maximum(worst_abs, err.max(axis=1))
    eo = np.abs(obs - gold["obs"][:, s])
    worst_obs = np.maximum(worst_obs, (eo / (H.RTOL * np.abs(gold["obs"][:, s]) + H.ATOL)).max(axis=0))
    rew_mismatch += int((rew != gold["reward"][:, s]).sum())
names = ["iR", "iI", "xR", "xI", "uR", "uI"] * em.cfg.phases + ["Vdc", "xDC", "xQ", "xPLL", "delta"]
print(f"scheme {scheme} {model_type}: {n} envs x {nsteps} env steps; integer reward mismatches {rew_mismatch}")
print(" state err/tol:", " ".join(f"{a}:{w:.2f}" for a, w in zip(names[-11:], worst[-11:])))
print(" state abs err:", " ".join(f"{a}:{w:.1e}" for a, w in zip(names[-11:], worst_abs[-11:])))
print(" obs   err/tol:", " ".join(f"{w:.2f}" for w in worst_obs))
