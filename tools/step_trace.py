"""Dev tool: per-env-step kernel time over a full episode plus the slow-path counters
(anti-windup sub-steps, sub-steps redone with library transcendentals)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_pvder_b200 as G

model = sys.argv[1] if len(sys.argv) > 1 else "model_1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 160
env = G.PVDERVecEnv(n, model_type=model, DISCRETE_REWARD=False, seed=2026, auto_reset=False)
env.reset()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
acts = [env.sample_actions().clone() for _ in range(3)]
for a in acts:
    env.step(a)
env.reset()
env._step_index = 0
rows = []
for s in range(steps):
    a = env.sample_actions()
    ev[s].record()
    env.step(a)
    ev[s + 1].record()
    if s % 10 == 9 or s < 6:
        st = env.stats().cpu().numpy().copy()
        rows.append((s, st[9], st[11]))
torch.cuda.synchronize()
ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(steps)]
prev = (0, 0)
for s, w, x in rows:
    print(f"step {s:3d}: kernel {ms[s]:.3f} ms   cumulative windup sub-steps/env {w / n:.3f}  exact sub-steps/env {x / n:.3f}")
print("mean ms", sum(ms) / len(ms), "min", min(ms), "max", max(ms))
