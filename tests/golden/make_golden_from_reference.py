"""PINNING SCRIPT -- turns "parity unpinned" into pinned the moment the reference can run.

Runs the UNMODIFIED reference environment (gym_PVDER.envs.PVDER_env.PVDER: reset 316-334, step 138-196, on top of the
third-party `pvder` simulator it imports at :26-35) and stores what it returns -- observations, rewards, done flags --
plus the event values it drew, as tests/golden/reference_pinned.npz.  tests/test_reference_pinning.py then compares the
oracle tiers (O2 reference-configured LSODA, O1 tight) and the CUDA kernels (C++ build on the CPU tier, the device on
the GPU tier) against those vectors.

`pvder` is NOT available in the image this repo was built in (not installed, not in the wheelhouse, no network; SURVEY.md
8c), so the committed tree carries no such file and the test skips with a loud PARITY UNPINNED.  Where it IS available:

    pip install pvder            # or: PYTHONPATH=/path/to/SolarPV-DER-simulation-utility
    python tests/golden/make_golden_from_reference.py [--reference /path/to/gym-SolarPVDER-environment]

Search order for the reference env: --reference, $PVDER_REFERENCE, baseline/_ref, oracle/_ref, /root/reference.
`gym` is only needed for its base classes; when it is missing a stand-in with Env / spaces / seeding is installed so the
reference module imports unmodified.

Cases (BASELINE.json configs 1-2):
  config1_model_2   default kwargs (n = 15, 40 s, DISCRETE_REWARD, voltage_regulation), the reset() default three-phase
                    model, global `random` seeded with 0 (the reference's unseeded event generator, PVDER_env.py:400-411),
                    a random agent drawing from numpy RandomState(0); one full 160-step episode
  config2_zero      single-phase model_1 (setup_PVDER_simulation('model_1'), the reference's own method), events disabled,
                    action 0 throughout, continuous reward
  config2_cycle     same with the action cycle 1,1,2,0,3,4 (reaches the current limit at env step ~103)
For every case: actions, obs[160, 11] float64, reward, done, and the event tables probed from the reference's own
SimulationEvents object at the instants 1..38 s (value in force from each instant on).
"""
import argparse
import os
import random
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N_STEPS = 160
CYCLE = [1, 1, 2, 0, 3, 4]
OUT = os.path.join(ROOT, "tests", "golden", "reference_pinned.npz")


def _install_gym_stand_in():
    """Just enough of old gym for `class PVDER(gym.Env, ...)` and its class-level spaces to import."""
    gym = types.ModuleType("gym")

    class Env:
        metadata, spec = {}, None

        @property
        def unwrapped(self):
            return self

    class Discrete:
        def __init__(self, n):
            self.n = n

        def contains(self, x):
            return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

        __contains__ = contains

        def sample(self):
            return random.randrange(self.n)

    class Box:
        def __init__(self, low, high, shape, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool((x >= self.low).all() and (x <= self.high).all())

        __contains__ = contains

    gym.Env = Env
    gym.error = types.ModuleType("gym.error")
    gym.spaces = types.ModuleType("gym.spaces")
    gym.spaces.Discrete, gym.spaces.Box = Discrete, Box
    gym.utils = types.ModuleType("gym.utils")
    gym.utils.seeding = types.ModuleType("gym.utils.seeding")
    sys.modules.update({"gym": gym, "gym.error": gym.error, "gym.spaces": gym.spaces, "gym.utils": gym.utils,
                        "gym.utils.seeding": gym.utils.seeding})
    return gym


def load_reference(reference=None):
    """Returns (PVDER class, description) or (None, reason)."""
    candidates = [reference, os.environ.get("PVDER_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"),
                  os.path.join(ROOT, "oracle", "_ref"), "/root/reference"]
    for c in candidates:
        if c and os.path.isdir(c) and c not in sys.path:
            sys.path.insert(0, c)
    try:
        import pvder  # noqa: F401
    except ImportError as exc:
        return None, f"the reference's simulator `pvder` is not importable ({exc})"
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot
    try:
        import gym  # noqa: F401
    except ImportError:
        _install_gym_stand_in()
    try:
        from gym_PVDER.envs.PVDER_env import PVDER
    except Exception as exc:   # the 2020-era module on a modern stack
        return None, f"gym_PVDER.envs.PVDER_env does not import ({type(exc).__name__}: {exc})"
    import pvder

    return PVDER, f"pvder {getattr(pvder, '__version__', 'unknown version')}"


def probe_events(env, instants):
    """Event values in force from each instant on, read from the reference's own SimulationEvents object."""
    ev = env.sim.simulation_events
    v, s = [], []
    for t in instants:
        g = ev.grid_events(t + 1e-9)
        so = ev.solar_events(t + 1e-9)
        v.append(float(np.abs(np.atleast_1d(g)[0])) if not np.isscalar(g) else float(g))
        s.append(float(np.atleast_1d(so)[0]))
    return np.array(v), np.array(s)


def run_case(PVDER, name):
    kw = dict(goals_list=["voltage_regulation"], n_sim_time_steps_per_env_step=15, max_sim_time=40.0)
    model_type = "model_2" if name == "config1_model_2" else "model_1"
    env = PVDER(DISCRETE_REWARD=(name == "config1_model_2"), **kw)
    if env.spec is None:
        env.spec = types.SimpleNamespace(id="PVDER-v0", max_episode_steps=500)     # what gym.make attaches
    if name != "config1_model_2":
        env.update_env_events([{"voltage": {"ENABLE": False}}])
    random.seed(0)
    if model_type == "model_2":
        obs0 = env.reset()
    else:   # reset() with the single-phase model: the reference's own steps of reset(), model_type passed through
        if hasattr(env, "sim"):
            env.cleanup_PVDER_simulation()
        env.initialize_environment_variables()
        env.setup_PVDER_simulation(model_type="model_1")
        obs0 = np.array(env.state)
    rs = np.random.RandomState(0)
    acts = {"config1_model_2": [int(rs.randint(0, 5)) for _ in range(N_STEPS)], "config2_zero": [0] * N_STEPS,
            "config2_cycle": [CYCLE[k % 6] for k in range(N_STEPS)]}[name]
    vt, st = probe_events(env, np.arange(1.0, 39.0, 1.0))
    obs = np.zeros((N_STEPS, 11))
    rew = np.zeros(N_STEPS)
    done = np.zeros(N_STEPS, dtype=bool)
    for k, a in enumerate(acts):
        o, r, d, _ = env.step(a)
        obs[k], rew[k], done[k] = np.asarray(o, dtype=np.float64), r, d
        if d:
            break
    return {"model_type": model_type, "discrete": name == "config1_model_2", "actions": np.array(acts, dtype=np.int32),
            "obs0": np.asarray(obs0, dtype=np.float64), "obs": obs, "reward": rew, "done": done, "steps_run": k + 1,
            "vgrid_tab": vt, "sinsol_tab": st}


CASES = ("config1_model_2", "config2_zero", "config2_cycle")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=None)
    args = ap.parse_args()
    PVDER, what = load_reference(args.reference)
    if PVDER is None:
        print("PARITY UNPINNED:", what)
        print("nothing written; install pvder (github.com/sibyjackgrove/SolarPV-DER-simulation-utility) and re-run")
        return 2
    out = {"_source": np.array(f"unmodified reference env on {what}")}
    for name in CASES:
        res = run_case(PVDER, name)
        for k, v in res.items():
            out[f"{name}/{k}"] = np.asarray(v)
        print(name, "steps", res["steps_run"], "final obs", res["obs"][res["steps_run"] - 1])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
