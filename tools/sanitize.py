"""Dev tool: small reset/step/stats run for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_pvder_b200 as G

for model, bal in (("model_1", True), ("model_2", True), ("model_2", False), ("model_2", "split")):
    env = G.PVDERVecEnv(333, model_type=model, balanced_three_phase=bal, auto_reset=True, n_sim_time_steps_per_env_step=2,
                        max_sim_time=1.0, events_spec={"voltage": {"min": 0.9}, "insolation": {"ENABLE": True}}, obs_f64=True)
    env.reset()
    for s in range(32):
        env.step(env.sample_actions())
    env.record_trajectory(5, 3)
    env.step(env.sample_actions())
    env.generate_events()
    print(model, bal, env.stats().cpu().tolist()[:4])
torch.cuda.synchronize()
print("sanitize run done")
