"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Bit-exact numpy twins of the integer / event / reward parts of the CUDA step kernel
(gym-solarpvder-environment_b200/csrc/pvder_env_step.cuh):

  * philox4x32_10, event tables and random actions  <- draw_event / actions_kernel, which restate
    pvder's create_random_events as driven by reference gym_PVDER/envs/PVDER_env.py:400-411 and
    ``env.action_space.sample()`` (examples/gym_PVDER_environment_import_test.py:23) with a
    counter-based per-env stream instead of Python's global unseeded ``random``;
  * outputs_twin: observation (PVDER_env.py:531-542) and reward (PVDER_env.py:231-301) from a
    given fp64 state, every operation individually rounded in the kernel's order, so that
    "same state => same integer reward" can be asserted bit-for-bit at any batch size.

PARITY UNPINNED for the RNG stream itself (the reference's stream is unseeded and therefore
has no reproducible values); what is pinned is the distribution contract (one event per
instant, type chosen among the enabled ones, value uniform in [min, max]).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
STREAM_EVENTS, STREAM_ACTIONS = 0, 1
SQRT2 = 1.4142135623730951
ROT3 = [(1.0, 0.0), (-0.5, -0.86602540378443864676), (-0.5, 0.86602540378443864676)]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Counters are array-likes of uint32, key scalars."""
    c = [np.asarray(x, dtype=np.uint64) & MASK32 for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        n0 = ((p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0)) & MASK32
        n1 = p1 & MASK32
        n2 = ((p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1)) & MASK32
        n3 = p0 & MASK32
        c = [n0, n1, n2, n3]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c


def u53(hi, lo):
    a = hi >> np.uint64(5)
    b = lo >> np.uint64(6)
    return (a * np.uint64(67108864) + b).astype(np.float64) * (1.0 / 9007199254740992.0)


def event_tables_twin(seed, n_envs, env_offset, episode, ev_count, voltage_enable, insol_enable, v_min, v_max,
                      s_min, s_max):
    """[ev_count, n_envs] tables of the Vgrid / Sinsol values in force from event instant j on."""
    env = np.arange(env_offset, env_offset + n_envs, dtype=np.uint64)
    ep = np.broadcast_to(np.asarray(episode, dtype=np.uint64), (n_envs,))
    vt = np.ones((max(1, ev_count), n_envs))
    st = np.full((max(1, ev_count), n_envs), 100.0)
    V = np.ones(n_envs)
    S = np.full(n_envs, 100.0)
    for j in range(ev_count):
        if voltage_enable or insol_enable:
            r = philox4x32_10(env, ep, np.uint64(j), np.uint64(STREAM_EVENTS), seed & 0xFFFFFFFF, seed >> 32)
            if voltage_enable and insol_enable:
                is_v = (r[0] >> np.uint64(31)) != 0
            else:
                is_v = np.full(n_envs, bool(voltage_enable))
            u = u53(r[1], r[2])
            V = np.where(is_v, v_min + (v_max + (-v_min)) * u, V)
            S = np.where(~is_v, s_min + (s_max + (-s_min)) * u, S)
        vt[j] = V
        st[j] = S
    return vt, st


def sample_actions_twin(seed, step_index, n_envs, env_offset):
    env = np.arange(env_offset, env_offset + n_envs, dtype=np.uint64)
    r = philox4x32_10(env, np.uint64(step_index & 0xFFFFFFFF), np.uint64(step_index >> 32), np.uint64(STREAM_ACTIONS),
                      seed & 0xFFFFFFFF, seed >> 32)
    return ((r[0] * np.uint64(5)) >> np.uint64(32)).astype(np.int32)


def outputs_twin(par, phases, y, Qref, Vdcref, Vgrid, Sinsol, k, max_sim_time, goal, discrete, reward_terms=None):
    """par: object with Rt, Xt, vgs, Vrms_ref, p_target, q_target, np_iph100, np_irs, kappa, pv_scale.
    y: [ns, N].  Returns (obs[N,11] float64, reward[N] float64, Vrms[N])."""
    y = np.asarray(y, dtype=np.float64)
    B = 6 * phases
    vg = Vgrid * par.vgs
    P = np.zeros(y.shape[1])
    Q = np.zeros(y.shape[1])
    v2 = np.zeros(y.shape[1])
    vaR = vaI = None
    for ph in range(phases):
        rr, ri = ROT3[ph] if phases == 3 else (1.0, 0.0)
        jR, jI = y[6 * ph], y[6 * ph + 1]
        vkR = vg * rr + (par.Rt * jR + (-(par.Xt * jI)))
        vkI = vg * ri + (par.Xt * jR + par.Rt * jI)
        P = P + 0.5 * (vkR * jR + vkI * jI)
        Q = Q + 0.5 * (vkI * jR + (-(vkR * jI)))
        v2 = v2 + (vkR * vkR + vkI * vkI)
        if ph == 0:
            vaR, vaI = vkR, vkI
    Vrms = np.sqrt(v2) / SQRT2 if phases == 1 else np.sqrt(v2 / 3.0) / SQRT2
    np_iph = par.np_iph100 * (Sinsol / 100.0)
    e = np.exp(par.kappa * y[B])
    Ipv = np_iph - par.np_irs * (e - 1.0)
    Ppv = np.maximum(Ipv * y[B] * par.pv_scale, 0.0)
    obs = np.stack([y[0], y[1], vaR, vaI, P, Q, y[B], Ppv, np.broadcast_to(Vdcref, P.shape),
                    np.broadcast_to(Qref, P.shape), (np.asarray(k, dtype=np.float64) / 120.0) / max_sim_time], axis=1)
    terms = reward_terms if reward_terms is not None else (goal,)      # default: the goal's required term (ids coincide)
    rsum = np.zeros(P.shape)
    for tid in terms:
        lo, hi = 0.01, 0.05
        if tid == 0:
            x, target = Vrms, par.Vrms_ref
        elif tid == 1:
            x = Q
            target = np.asarray(Qref, dtype=np.float64) if goal == 0 else par.q_target
            if discrete:
                target = np.where(target == 0.0, 1e-6, target)
        elif tid == 2:
            x, target, hi = P, par.p_target, 0.03
        else:
            x, target, lo = y[B], np.asarray(Vdcref, dtype=np.float64), 0.02
        if discrete:
            err = np.abs(x + (-target)) / np.abs(target)
            rsum = rsum + np.where(err <= lo, 1.0, np.where(err >= hi, -5.0, -1.0))
        else:
            d = x + (-target)
            rsum = rsum + (-(d * d))
    return obs, rsum, Vrms
