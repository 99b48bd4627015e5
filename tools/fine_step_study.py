"""Design study (not product, not test): worst error / tolerance over the FULL-episode fixtures
(tests/golden/golden_episode_*.npz) for the fine-step settings of EnvConfig (refine_input_level, refine_on_action,
startup_substeps, startup_level), with the kernel source compiled as plain C++ (tests/host_emul).

    python tools/fine_step_study.py [model_1|model_2]
"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, emul_harness as E, helpers as H
model = sys.argv[1] if len(sys.argv) > 1 else "model_1"
gold = np.load(f'tests/golden/golden_episode_{model}.npz')
acts = gold['actions']; n, _ = acts.shape
for name, kw in [("off", dict(refine_input_level=0, startup_level=0)),
                 ("events only L1", dict(refine_input_level=1, refine_on_action=False, startup_level=0)),
                 ("events+actions L1", dict(refine_input_level=1, refine_on_action=True, startup_level=0)),
                 ("events only L2", dict(refine_input_level=2, refine_on_action=False, startup_level=0)),
                 ("events+actions L1 + startup 12@3", dict(refine_input_level=1, refine_on_action=True, startup_substeps=12, startup_level=3)),
                 ("events L1 + startup 12@3", dict(refine_input_level=1, refine_on_action=False, startup_substeps=12, startup_level=3))]:
    em = E.EmulVecEnv(n, model_type=model, events_spec=H.SAG_SPEC, event_mode='table', DISCRETE_REWARD=False, **kw)
    em.set_event_tables(gold['vgrid_tab'], gold['sinsol_tab']); em.reset()
    ns = em.ns; B = 6 * em.cfg.phases
    worst = np.zeros((n, ns)); at = np.zeros((n, ns), int)
    for s in range(160):
        em.step(acts[:, s])
        for i in range(n):
            if gold['windup'][i, s] > 0:
                continue
            y = em.sd[:ns, i]; yr = gold['state'][i, s]
            err = np.abs(y - yr); tol = 1e-5 * np.abs(yr) + 1e-7; tol[B + 3] = 2e-4; tol[B + 4] = 5e-6
            r = err / tol
            upd = r > worst[i]; worst[i][upd] = r[upd]; at[i][upd] = s
    print(name, "| windup counts", em.si[10, :n], "gold", gold['windup'][:, -1])
    for i in range(n):
        j = int(np.argmax(worst[i]))
        print(f"   traj{i}: worst err/tol {worst[i, j]:.2f} state {j} at step {at[i, j]}; tail", np.round(worst[i][-11:], 2))
