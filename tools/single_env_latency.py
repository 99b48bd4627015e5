"""Dev tool: config 1 on the GPU path -- latency of ONE PVDER-v0 env stepped through the Gym facade
(host-buffer C ABI: action H2D, one kernel launch, obs/reward/done D2H, wait), random agent, one episode."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_pvder_b200 as gym_PVDER

out = {}
for model in ("model_2", "model_1"):
    env = gym_PVDER.make("PVDER-v0", model_type=model, seed=0)
    env.reset()
    for _ in range(5):
        env.step(env.action_space.sample())
    env.reset()
    n, done = 0, False
    t0 = time.perf_counter()
    while not done:
        _, _, done, _ = env.step(env.action_space.sample())
        n += 1
    dt = time.perf_counter() - t0
    out[model] = {"env_steps": n, "ms_per_env_step": 1e3 * dt / n, "env_steps_per_s": n / dt}
    env.close()
print(json.dumps({"config": "single PVDER-v0 env, random agent, one 160-step episode through the Gym facade", **out}))
