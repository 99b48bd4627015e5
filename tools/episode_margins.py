"""Design study (not product, not test): worst error / tolerance of a scheme over the FULL-episode fixtures
(tests/golden/golden_episode_*.npz), per trajectory and state, with the kernel source compiled as plain C++.

    python tools/episode_margins.py [4|6] [model_1|model_2]
"""
import sys, ctypes as C, subprocess
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, emul_harness as E, helpers as H
scheme=sys.argv[1]; model=sys.argv[2]
lib=f"/tmp/libpvder_emul_s{scheme}.so"
subprocess.run(["g++","-O2","-std=c++17","-ffp-contract=off","-fPIC","-shared","-Wno-unknown-pragmas",f"-DPVDER_SCHEME={scheme}","-o",lib,E._SRC],check=True)
E._lib=C.CDLL(lib)
gold=np.load(f'tests/golden/golden_episode_{model}.npz')
acts=gold['actions']; n,_=acts.shape
em=E.EmulVecEnv(n, model_type=model, events_spec=H.SAG_SPEC, event_mode='table', DISCRETE_REWARD=False)
em.set_event_tables(gold['vgrid_tab'], gold['sinsol_tab']); em.reset()
ns=em.ns; B=6*em.cfg.phases
worst=np.zeros((n,ns)); at=np.zeros((n,ns),int)
for s in range(160):
    em.step(acts[:,s])
    for i in range(n):
        y=em.sd[:ns,i]; yr=gold['state'][i,s]
        err=np.abs(y-yr); tol=1e-5*np.abs(yr)+1e-7; tol[B+3]=2e-4; tol[B+4]=5e-6
        r=err/tol
        upd=r>worst[i]; worst[i][upd]=r[upd]; at[i][upd]=s
for i in range(n):
    j=int(np.argmax(worst[i])); print(f"scheme {scheme} {model} traj{i}: worst err/tol {worst[i,j]:.2f} state {j} at step {at[i,j]}; per-state", np.round(worst[i][-11:],2))
print("vgrid table traj2 first events:", gold['vgrid_tab'][:4,2], gold['sinsol_tab'][:4,2])
