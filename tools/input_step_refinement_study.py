"""Design study (not product, not test): where the one above-floor point of the full-episode fixture comes from.
Replays the random/sag trajectory of tests/golden/golden_episode_model_1.npz with a numpy ROS4-L (the kernel's scheme)
and refines ONLY the first half-cycle sub-step after an action / event into 1, 2 or 4 equal steps.  Quoted in DESIGN.md.

    python tools/input_step_refinement_study.py
"""
import sys, math
sys.path.insert(0, "."); sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import integrator_study as S
import numpy as np
import helpers as H
import gym_pvder_b200 as G
from oracle.env_oracle import OraclePVDEREnv, EventTable
from oracle.pvder_model import Inputs
gold = np.load('tests/golden/golden_episode_model_1.npz')
t = 2
acts = gold['actions'][t]; vt = gold['vgrid_tab'][:, t]; st = gold['sinsol_tab'][:, t]
cfg = G.EnvConfig(model_type='model_1', events_spec=H.SAG_SPEC, event_mode='table')
c = cfg.c
env = OraclePVDEREnv(model_type='model_1', solver='tight', events=EventTable(), DISCRETE_REWARD=False); env.reset()
n = env.model.n
y0 = env.y.copy()   # t = 0: delta = wte
def run(refine):
    y = y0.copy(); Q = env.Q_ref; V = env.Vdc_ref; vg = 1.0; si = 100.0; k = 0
    out = []
    for s in range(8):
        a = acts[s]
        Q += env.delQ_pu if a == 1 else -env.delQ_pu if a == 2 else 0.0
        V += env.delVdc_pu if a == 3 else -env.delVdc_pu if a == 4 else 0.0
        changed = True      # an action (possibly 0) starts every env step
        for sub in range(30):
            if k >= c.ev_start_k and (k - c.ev_start_k) % c.ev_step_k == 0:
                j = (k - c.ev_start_k) // c.ev_step_k
                if j < c.ev_count:
                    if vt[j] != vg or st[j] != si: changed = True
                    vg, si = vt[j], st[j]
            inp = Inputs(Vgrid=vg, Sinsol=si, Q_ref=Q, Vdc_ref=V, freeze=(False,)*6)
            f, J = S.make_aut(env.model, inp)
            m = refine if changed else 1
            for _ in range(m): y = S.rosenbrock_step(f, J, y, S.H / m, S.ROS4L)
            changed = False
            k += 1
        out.append(y.copy())
    return np.array(out)
for refine in (1, 2, 4):
    Y = run(refine)
    yr = gold['state'][t, :8]
    err = np.abs(Y - yr); tol = 1e-5*np.abs(yr) + 1e-7; tol[:, 9] = 2e-4; tol[:, 10] = 5e-6
    print("refine first sub-step x%d:" % refine, "step 4 err/tol", np.round((err/tol)[4], 2), " worst steps 1..7", round((err/tol)[1:].max(), 2))
