"""BASELINE.json config 5: DQN-style collect loop over 262,144 batched envs, observation/action
round trip to the policy every step, all on one GPU.  Prints one JSON line per variant."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gym_pvder_b200 as G
from gym_pvder_b200.rollout import DQNRollout

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 160
model = sys.argv[3] if len(sys.argv) > 3 else "model_2"
out = {}
for name, graph, policy in (("fused_policy_cuda_graph", True, "fused"), ("policy_cuda_graph", True, "torch"),
                            ("policy_eager", False, "torch")):
    venv = G.PVDERVecEnv(n, model_type=model, DISCRETE_REWARD=True, auto_reset=True, seed=1)
    venv.reset()
    out[name] = DQNRollout(venv, epsilon=0.1, use_cuda_graph=graph, policy=policy).collect(steps)
venv = G.PVDERVecEnv(n, model_type=model, DISCRETE_REWARD=True, auto_reset=True, seed=1)
venv.reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    venv.step(venv.sample_actions())
e0.record()
for _ in range(steps):
    venv.step(venv.sample_actions())
e1.record()
torch.cuda.synchronize()
out["env_only_random_actions"] = {"env_steps_per_s": n * steps / (e0.elapsed_time(e1) * 1e-3),
                                  "ms_per_iteration": e0.elapsed_time(e1) / steps}
out["policy_overhead_ms_per_step"] = out["policy_cuda_graph"]["ms_per_iteration"] - out["env_only_random_actions"]["ms_per_iteration"]
out["fused_policy_overhead_ms_per_step"] = out["fused_policy_cuda_graph"]["ms_per_iteration"] - out["env_only_random_actions"]["ms_per_iteration"]
print(json.dumps({"config": f"DQN rollout, {n} envs, {model}, Q-net 11-100-5 fp32, eps=0.1", **out}))
