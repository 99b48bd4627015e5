"""Dev tool: soak run -- 1 Mi envs x 3 full episodes (480 env steps, auto-reset) per kernel with a random policy, voltage sags
to 0.90 pu and insolation steps: ~5e8 env steps per kernel through every path the events and the policy reach (fine steps,
library-transcendental redo under partial masks in the three-lane kernel, anti-windup operation, auto-reset).  Reports
failures (must be 0), the slow-path counters and the time."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_pvder_b200 as G

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 480
out = {}
for name, kw in (("model_1", dict(model_type="model_1")), ("model_2_auto", dict(model_type="model_2", balanced_three_phase="auto")),
                 ("model_2_split", dict(model_type="model_2", balanced_three_phase="split")),
                 ("model_2_split_unbalanced_grid", dict(model_type="model_2", balanced_three_phase="split", grid_unbalance_ratio=(0.95, 1.03)))):
    env = G.PVDERVecEnv(n, DISCRETE_REWARD=True, seed=77, auto_reset=True,
                        events_spec={"voltage": {"min": 0.90, "max": 1.02, "ENABLE": True}, "insolation": {"ENABLE": True}}, **kw)
    env.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    episodes_done = 0
    bad = 0
    for s in range(steps):
        obs, rew, done, _ = env.step(env.sample_actions())
        if s % 160 == 159:
            episodes_done += int(done.sum())
            bad += int((~torch.isfinite(obs)).any(dim=1).sum())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = env.stats().cpu().tolist()
    out[name] = {"envs": n, "env_steps": n * steps, "wall_s": round(dt, 3), "env_steps_per_s": n * steps / dt,
                 "episodes_completed": episodes_done, "non_finite_obs_rows": bad, "failed_envs_now": st[3],
                 "status_nonzero": int((env.status != 0).sum())}
    assert bad == 0 and st[3] == 0, out[name]
print(json.dumps(out, indent=1))
