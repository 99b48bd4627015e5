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

# The three-lane kernel's divergent regions: the shared-memory exchange of the phase sums and every vote run under a
# partial member mask when only some env groups of a warp need the out-of-line path.  (a) split mode with a few envs whose
# PLL angle is kicked far from lock (their stages leave the incremental range: library-transcendental redo for those groups
# only) and a few that blow up (quarantine votes); (b) auto mode with unbalanced envs handed over through the redo list
# (step_kernel_split3<LIST>).
for mode in ("split", "auto"):
    env = G.PVDERVecEnv(333, model_type="model_2", balanced_three_phase=mode, auto_reset=False, n_sim_time_steps_per_env_step=3,
                        max_sim_time=2.0, events_spec={"voltage": {"min": 0.9}, "insolation": {"ENABLE": True}})
    env.reset()
    for s in range(3):
        env.step(env.sample_actions())
    for i in (1, 14, 15, 77, 200, 331):
        env.sd[22, i] += 1.3                      # PLL angle 75 degrees off: pull-in with stages out of range
    for i in (5, 31, 250):
        for ph in range(3):
            env.sd[6 * ph + 2, i] *= 11.0         # duty-cycle integrators scaled: the DC link collapses (quarantine)
            env.sd[6 * ph + 3, i] *= 11.0
    for i in (40, 41, 99):
        env.sd[6, i] *= 1.05                      # phase b current off: an unbalanced stored state (auto: redo list)
    for s in range(12):
        env.step(env.sample_actions())
    print("divergent regions", mode, env.stats().cpu().tolist()[:4], "exact sub-steps", int(env.si[11].sum()))
torch.cuda.synchronize()
print("sanitize run 2 done")
