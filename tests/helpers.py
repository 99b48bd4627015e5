"""Shared helpers of the test-suite (oracle-side plumbing)."""
import math
import random

import numpy as np

from oracle.env_oracle import DEFAULT_EVENTS_SPEC, EventTable, OraclePVDEREnv, create_random_events

W = 2.0 * math.pi * 60.0
SAG_SPEC = {"voltage": {"min": 0.90, "max": 1.02, "ENABLE": True}, "insolation": {"ENABLE": True}}


def full_spec(partial):
    spec = {k: dict(v) for k, v in DEFAULT_EVENTS_SPEC.items()}
    for k, v in (partial or {}).items():
        spec[k].update(v)
    return spec


def oracle_tables(events: EventTable, c):
    """Forward-filled [ev_count, 1] tables (value in force from instant j on) of an oracle EventTable."""
    K = max(1, c.ev_count)
    v = np.ones((K, 1))
    s = np.full((K, 1), 100.0)
    for j in range(c.ev_count):
        T = (c.ev_start_k + j * c.ev_step_k) / 120.0
        v[j, 0] = events.vgrid(T)
        s[j, 0] = events.sinsol(T)
    return v, s


def table_to_events(vcol, scol, c):
    """Inverse: an oracle EventTable reproducing forward-filled columns."""
    ev = EventTable()
    for j in range(c.ev_count):
        T = (c.ev_start_k + j * c.ev_step_k) / 120.0
        ev.add_grid_event(T, float(vcol[j]))
        ev.add_solar_event(T, float(scol[j]))
    return ev


def oracle_delta_state(orc: OraclePVDEREnv):
    y = orc.y.copy()
    y[-1] -= W * orc.t()
    return y


# Tolerances (DESIGN.md "Tolerances"): per-unit electrical + controller states and all 11
# observations |err| <= 1e-5*|ref| + 1e-7 pu; the two PLL states have their own absolute bounds:
# xPLL is a frequency deviation in rad/s that is ~0 at lock (2e-4 rad/s = 5e-7 of w_e = 377 rad/s),
# delta is the PLL angle in rad.
RTOL, ATOL = 1e-5, 1e-7


def assert_state_close(y_gpu, y_ref, phases, rtol=RTOL, atol=ATOL, xpll_atol=2e-4, delta_atol=5e-6, what=""):
    B = 6 * phases
    main = list(range(B + 3))
    np.testing.assert_allclose(y_gpu[main], y_ref[main], rtol=rtol, atol=atol, err_msg=f"{what} states")
    assert abs(y_gpu[B + 3] - y_ref[B + 3]) <= xpll_atol, f"{what} xPLL {y_gpu[B + 3]} vs {y_ref[B + 3]}"
    assert abs(y_gpu[B + 4] - y_ref[B + 4]) <= delta_atol, f"{what} delta {y_gpu[B + 4]} vs {y_ref[B + 4]}"


def random_events(seed, partial=SAG_SPEC):
    return create_random_events(full_spec(partial), random.Random(seed))


def assert_episode_step_close(y, obs, y_ref, obs_ref, phases, in_windup, what="", atol=ATOL):
    """One env step of a full-episode fixture against the oracle tier that samples the anti-windup clamp like the kernel
    (once per half-cycle): the normal tolerances, or -- from the first env step on at which the oracle reports anti-windup
    sub-steps -- 2e-4 relative / 1e-6 absolute on the electrical and controller states and observations, 2e-3 rad/s /
    2e-5 rad on the PLL states.  Why looser there: near the limit the clamp decision |i_ref| > iref_limit is knife-edge,
    so a 1e-7 difference between two integrators can flip it for one half-cycle, which moves the frozen integrator by up to
    Ki * error * h.  MEASURED (kernel source vs this tier, 8 full random-policy episodes with a +Q bias, 673 clamped env
    steps, identical windup counts in all of them): worst error 13.9x the normal tolerance (1.4e-4 relative), median
    episode 0.07x; on the committed fixtures 0.064x.  The bound is that measurement with a 1.4x margin; the gap to the
    CONTINUOUS clamp semantics is a separate, larger number (CONTINUOUS_CLAMP_ATOL below)."""
    if not in_windup:
        assert_state_close(y, y_ref, phases, atol=atol, what=what)
        np.testing.assert_allclose(obs, obs_ref, rtol=RTOL, atol=atol, err_msg=f"{what} obs")
    else:
        assert_state_close(y, y_ref, phases, rtol=2e-4, atol=1e-6, xpll_atol=2e-3, delta_atol=2e-5, what=what + " (windup)")
        np.testing.assert_allclose(obs, obs_ref, rtol=2e-4, atol=1e-6, err_msg=f"{what} obs (windup)")


def voltage_error_margin(cfg, sd_cols):
    """Distance of the relative voltage error (PVDER_env.py:280-285) of the given envs (columns of the sd matrix) from
    the two class thresholds 0.01 / 0.05: a discrete reward may legitimately differ between two implementations only
    for a state that sits on a threshold to within the trajectory tolerance."""
    from oracle import twin

    ns = cfg.n_state
    _, _, vrms = twin.outputs_twin(cfg.par, cfg.phases, sd_cols[:ns], sd_cols[ns], sd_cols[ns + 1], sd_cols[ns + 2],
                                   sd_cols[ns + 3], np.zeros(sd_cols.shape[1]), cfg.max_sim_time, 0, True)
    err = np.abs(vrms - cfg.par.Vrms_ref) / abs(cfg.par.Vrms_ref)
    return np.minimum(np.abs(err - 0.01), np.abs(err - 0.05))


def assert_reward_mismatches_sit_on_a_threshold(cfg, sd_cols, mismatch, budget=1, tol=1e-5):
    """At most `budget` envs may carry a different discrete reward, and each of them only because its voltage error is
    within `tol` of a class threshold."""
    idx = np.flatnonzero(mismatch)
    assert len(idx) <= budget, f"{len(idx)} discrete rewards differ"
    if len(idx):
        margin = voltage_error_margin(cfg, sd_cols[:, idx])
        assert (margin < tol).all(), f"discrete reward differs {margin} away from a threshold"


# Anti-windup regime against the CONTINUOUS-clamp oracle (oracle/env_oracle.py solver="tight_continuous", fixtures
# tests/golden/golden_continuous_clamp_model_1.npz).  The kernel -- like the "tight" oracle tier -- decides the clamp once
# per half-cycle; pvder decides it inside every right-hand-side call.  MEASURED gap between the two
# semantics (tight-oracle tier against tight-oracle tier, tests/golden/make_golden_continuous.py), worst |state error|:
#   +Q cycle of config 2 (limit reached at env step 103, 57 clamped env steps)            9.5e-4 pu  (xQ, at the onset)
#   +Q-biased random policy with sags (clamp releases for a few half-cycles at step 128)  7.6e-3 pu  (xQ: the integrator
#       runs at Ki_Q dQ ~ 10 pu/s while released, so the instant of re-engagement -- resolved to 1/120 s -- matters)
# and 0.016 x the normal tolerance before the limit is reached.  The kernel reproduces the sampled tier (0.06x .. 14x the
# normal tolerance, windup counts identical), so the bounds below are the gap itself with a 1.3x..2x margin, not a
# statement about the integrator.  The reference's own LSODA call (O2) FAILS at the release of the second trajectory.
CONTINUOUS_CLAMP_ATOL = {0: 2e-3, 1: 1e-2}


def assert_vs_continuous_clamp(y, obs, gold, traj, s, what=""):
    """One env step against the continuous-clamp fixture: normal tolerances until the oracle first reports an active clamp
    (in either clamp semantics), the measured-gap bound afterwards."""
    clamped = gold["windup"][traj, s] > 0 or gold["windup_sampled"][traj, s] > 0
    yr, orf = gold["state"][traj, s], gold["obs"][traj, s]
    if not clamped:
        assert_state_close(y, yr, 1, what=what)
        np.testing.assert_allclose(obs, orf, rtol=RTOL, atol=ATOL, err_msg=f"{what} obs")
    else:
        atol = CONTINUOUS_CLAMP_ATOL[traj]
        np.testing.assert_allclose(y[:9], yr[:9], rtol=0, atol=atol, err_msg=f"{what} states (continuous clamp)")
        assert abs(y[10] - yr[10]) <= atol and abs(y[9] - yr[9]) <= 100 * atol, f"{what} PLL states (continuous clamp)"
        np.testing.assert_allclose(obs, orf, rtol=0, atol=atol, err_msg=f"{what} obs (continuous clamp)")
