"""Design study (not product, not test): worst error / tolerance at n_sim_time_steps_per_env_step = 1 over the first 60 env steps (the PLL
pull-in after reset) for cheaper start-up fine-step settings than the default 12 sub-steps at level 3, kernel source compiled as plain
C++ against the tight oracle.  Result: 12@3 (default) 0.12; 12@2 / 8@2 / 16@2 1.59; 8@3 1.28; 6@3 5.6; 4@3 24 -- the default is needed."""
import sys, random
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, emul_harness as E, helpers as H
from oracle.env_oracle import OraclePVDEREnv
def run(model_type, **kw):
    ev = H.random_events(11)
    em = E.EmulVecEnv(1, model_type=model_type, events_spec=H.SAG_SPEC, event_mode="table", DISCRETE_REWARD=True,
                      n_sim_time_steps_per_env_step=1, max_sim_time=4.0, **kw)
    em.set_event_tables(*H.oracle_tables(ev, em.cfg.c))
    orc = OraclePVDEREnv(model_type=model_type, solver="tight", events=ev, DISCRETE_REWARD=True,
                         n_sim_time_steps_per_env_step=1, max_sim_time=4.0)
    em.reset(); orc.reset()
    rng = random.Random(5)
    worst = 0.0; at=-1
    P = em.cfg.phases; B = 6*P
    for s in range(60):
        a = rng.randrange(5)
        oo, orw, od, _ = orc.step(a)
        eo, erw, ed, _ = em.step([a])
        y, yr = em.sd[:orc.model.n, 0], H.oracle_delta_state(orc)
        tol = 1e-5*np.abs(yr) + 1e-7; tol[B+3] = 2e-4; tol[B+4] = 5e-6
        r = float((np.abs(y-yr)/tol).max())
        ro = float((np.abs(eo[0]-oo)/(1e-5*np.abs(oo)+1e-7)).max())
        r = max(r, ro)
        if r > worst: worst, at = r, s
    return worst, at
for kw in (dict(), dict(startup_substeps=12, startup_level=2), dict(startup_substeps=8, startup_level=3), dict(startup_substeps=6, startup_level=3),
           dict(startup_substeps=8, startup_level=2), dict(startup_substeps=16, startup_level=2), dict(startup_substeps=4, startup_level=3)):
    print(kw, 'model_1 worst err/tol %.2f at step %d' % run('model_1', **kw), '| model_2 %.2f at %d' % run('model_2', **kw))
