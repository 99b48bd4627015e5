"""Random agent on PVDER-v0 -- the reference's examples/gym_PVDER_environment_import_test.py:15-31
with only the import changed (BASELINE.json config 1 runs the same loop on the CPU oracle)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_pvder_b200 as gym_PVDER

env = gym_PVDER.make('PVDER-v0')
for i in range(1):
    observation = env.reset()
    done = False
    while not done:
        action = env.action_space.sample()
        print('Action:', action)
        observation, reward, done, _ = env.step(action)
        env.render()
print(env.unwrapped.sim.PV_model.Q_ref, env.unwrapped.steps)
