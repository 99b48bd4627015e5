"""Pins the CPU oracle (CPU tier).  The reference holds NO golden vectors or numeric known-answer
tests for this path (gym_PVDER/tests/test_gym_PVDER.py checks types/shapes/counters only) and its
simulator dependency is absent, so the oracle is pinned against (a) the constants the reference
ships (config_der.json, PVDER_env.py:56-75) through the known answers recorded in SURVEY.md
Appendix B, (b) its own internal consistency (analytic Jacobian vs finite differences, tight vs
reference-configured LSODA), and (c) the committed golden fixtures it generated."""
import math
import random

import numpy as np
import pytest

import helpers as H
from oracle.env_oracle import (DEFAULT_EVENTS_SPEC, EventTable, OraclePVDEREnv, create_random_events, discrete_class)
from oracle.pvder_model import Inputs, PVDERModel, load_der_params, ppv_and_slope


def test_per_unit_known_answers():
    """SURVEY.md Appendix B (computed from config_der.json + A.0)."""
    p10, p50 = load_der_params("10"), load_der_params("50")
    for p in (p10, p50):
        assert p.extras["Zbase"] == 5.0 and p.Ibase == 100.0
        assert p.extras["Lbase"] == pytest.approx(0.01326291192, rel=1e-9)
        assert p.extras["Cbase"] == pytest.approx(5.30516477e-4, rel=1e-8)
        assert p.extras["Z2"] == pytest.approx(complex(0.322, 1.108))
        assert p.extras["vag"] == pytest.approx(40.83)
        assert p.extras["a"] == pytest.approx(81.556977, rel=1e-7)
        assert p.Rf == pytest.approx(4e-4) and p.Lf == pytest.approx(1.884955592e-3, rel=1e-9)
        assert p.extras["Z1"] == pytest.approx(complex(3.8e-4, 1.122e-2))
        assert p.Vrms_ref == pytest.approx(0.354)
    assert p10.C == pytest.approx(0.05654866776, rel=1e-9) and p10.Vdc_ref0 == 1.5
    assert p50.C == pytest.approx(0.5654866776, rel=1e-9) and p50.Vdc_ref0 == 1.1
    assert ppv_and_slope(p10, 1.5, 100.0)[0] == pytest.approx(0.2263919636, rel=1e-9)
    assert ppv_and_slope(p50, 1.1, 100.0)[0] == pytest.approx(0.9150979508, rel=1e-9)
    # gains of derId 10 are the doubled derId 50 gains (config_der.json:12-14 vs :76-78)
    for k in ("Kp_GCC", "Ki_GCC", "Kp_DC", "Ki_DC", "Kp_Q", "Ki_Q"):
        assert getattr(p10, k) == 2 * getattr(p50, k)
    assert p10.wp == p50.wp == 20e4


def test_steady_state_known_answers():
    """SURVEY.md A.6 probed values; P at the PCC ~ P_ref = 45.4 kW (PVDER_env.py:69)."""
    m10, m50 = PVDERModel(load_der_params("10")), PVDERModel(load_der_params("50"))
    y0, ma0, ia0 = m10.steady_state()
    assert ma0 == pytest.approx(complex(0.66818, 0.01600), abs=1e-5)
    assert ia0 == pytest.approx(complex(0.90308, 0.01856), abs=1e-5)
    y0, ma0, ia0 = m50.steady_state()
    assert ma0 == pytest.approx(complex(0.91126, 0.02940), abs=1e-5)
    assert ia0 == pytest.approx(complex(1.21614, 0.03366), abs=1e-5)
    out = m50.outputs(y0, Inputs(Vdc_ref=1.1))
    assert 45.3e3 < out["P_PCC"] * 50e3 < 45.8e3
    assert out["Vrms"] == pytest.approx(0.3542, abs=1e-4) and abs(out["Q_PCC"]) < 1e-12
    # everything but the PLL is at rest at y0 once the PLL is locked
    f = np.array(m50.rhs(list(y0), 0.0, Inputs(Vdc_ref=1.1)))
    assert np.abs(f[[2, 3, 4, 5, 18, 19, 20]]).max() < 1e-9


@pytest.mark.parametrize("der", ["10", "50"])
def test_jacobian_matches_finite_differences(der):
    p = load_der_params(der)
    m = PVDERModel(p)
    rng = np.random.default_rng(0)
    y = np.array(m.steady_state()[0])
    y = y * (1 + 0.05 * rng.standard_normal(m.n))
    y[-2], y[-1] = 0.4, 7.1
    for k in range(p.phases):
        y[6 * k + 4:6 * k + 6] = 1e-3 * rng.standard_normal(2)
    nofrz = (False,) * (4 * p.phases + 2)
    inp = Inputs(Vgrid=0.97, Sinsol=90, Q_ref=0.05, Vdc_ref=p.Vdc_ref0 * 1.01, freeze=nofrz)
    J = m.jac(list(y), 0.0123, inp)
    assert int((J != 0).sum()) == (41 if der == "10" else 165)
    Jfd = np.zeros_like(J)
    for j in range(m.n):
        h = 1e-6 * max(1.0, abs(y[j]))
        yp, ym = y.copy(), y.copy()
        yp[j] += h
        ym[j] -= h
        Jfd[:, j] = (np.array(m.rhs(list(yp), 0.0123, inp)) - np.array(m.rhs(list(ym), 0.0123, inp))) / (2 * h)
    assert np.max(np.abs(J - Jfd) / (1 + np.abs(Jfd))) < 1e-6


def test_stiffness_spectrum():
    """fast pair |lambda| ~ 1e6, slowest -1/3."""
    p = load_der_params("10")
    m = PVDERModel(p)
    y0 = m.steady_state()[0]
    out = m.outputs(y0, Inputs(Vdc_ref=1.5))
    y0[-1] = math.atan2(out["vaI"], out["vaR"]) + math.pi / 2      # PLL at lock: vd = 0 (A.4)
    ev = np.linalg.eigvals(m.jac(list(y0), 0.0, Inputs(Vdc_ref=1.5, freeze=(False,) * 6)))
    assert 0.9e6 < np.abs(ev).max() < 1.1e6 and ev.real.min() == pytest.approx(-1.0e5, rel=1e-2)
    assert ev.real.max() == pytest.approx(-1.0 / 3.0, rel=1e-4)


def test_discrete_reward_thresholds():
    """PVDER_env.py:280-285 / :292-297."""
    assert discrete_class(0.0, 0.05) == 1 and discrete_class(0.01, 0.05) == 1
    assert discrete_class(0.010001, 0.05) == -1 and discrete_class(0.0499, 0.05) == -1
    assert discrete_class(0.05, 0.05) == -5 and discrete_class(0.03, 0.03) == -5


def test_event_generator_contract():
    """PVDER_env.py:400-411 + :60-61: 38 instants 1..38 s, default voltage only in [0.98, 1.02]."""
    tab = create_random_events(DEFAULT_EVENTS_SPEC, random.Random(1))
    assert [T for T, _ in tab.grid] == [float(t) for t in range(1, 39)] and tab.solar == []
    assert all(0.98 <= v <= 1.02 for _, v in tab.grid)
    both = create_random_events(H.full_spec(H.SAG_SPEC), random.Random(1))
    assert len(both.grid) + len(both.solar) == 38 and both.grid and both.solar
    assert all(85.0 <= s <= 100.0 for _, s in both.solar)
    ev = EventTable()
    ev.add_grid_event(2.0, 0.95)
    assert ev.vgrid(1.999) == 1.0 and ev.vgrid(2.0) == 0.95 and ev.sinsol(5.0) == 100.0   # left-closed, defaults


@pytest.mark.parametrize("model_type", ["model_1", "model_2"])
def test_reference_config_lsoda_close_to_tight(model_type):
    """the reference's solver settings are themselves only ~1e-4..1e-3 accurate."""
    import warnings

    warnings.filterwarnings("ignore")
    ev = H.random_events(3)
    a = OraclePVDEREnv(model_type=model_type, solver="reference", events=ev, DISCRETE_REWARD=False)
    b = OraclePVDEREnv(model_type=model_type, solver="tight", events=ev, DISCRETE_REWARD=False)
    a.reset()
    b.reset()
    for act in [1, 3, 0, 2, 4, 0]:
        oa, ra, da, _ = a.step(act)
        ob, rb, db, _ = b.step(act)
        np.testing.assert_allclose(oa, ob, rtol=0, atol=2e-3)
    assert a.rhs_evals > 300


def test_env_semantics_time_steps():
    """reference test_time_steps (tests:87-111) on the oracle: 150 steps, last obs element 1.0."""
    import warnings

    warnings.filterwarnings("ignore")
    env = OraclePVDEREnv(n_sim_time_steps_per_env_step=10, max_sim_time=25.0, model_type="model_1", seed=0)
    ob = env.reset()
    assert ob[-1] == 0.0 and ob.shape == (11,)
    done, steps = False, 0
    rng = random.Random(0)
    while not done and steps < 12:
        ob, r, done, _ = env.step(rng.randrange(5))
        steps += 1
        assert isinstance(r, int) and r in (1, -1, -5)
    assert env.done_substep == 3000 and env.k == 20 * steps
    assert env.delQ_pu == pytest.approx(25 * 10 / 50e3) and env.delVdc_pu == pytest.approx(0.02 * 10 / 500)
    with pytest.raises(AssertionError):
        env.step(5)


@pytest.mark.parametrize("model_type", ["model_1", "model_2"])
def test_golden_fixture_is_reproducible(model_type):
    """The committed fixtures are what the oracle's tight path produces (first 2 steps of env 1)."""
    gold = np.load(f"tests/golden/golden_{model_type}.npz")
    import gym_pvder_b200 as G

    cfg = G.EnvConfig(model_type=model_type, events_spec=H.SAG_SPEC, event_mode="table")
    ev = H.table_to_events(gold["vgrid_tab"][:, 1], gold["sinsol_tab"][:, 1], cfg.c)
    env = OraclePVDEREnv(model_type=model_type, solver="tight", events=ev, DISCRETE_REWARD=True)
    env.reset()
    for s in range(2):
        o, r, d, _ = env.step(int(gold["actions"][1, s]))
        np.testing.assert_allclose(o, gold["obs"][1, s], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(H.oracle_delta_state(env), gold["state"][1, s], rtol=1e-8, atol=1e-9)
        assert r == gold["reward"][1, s]


def test_product_parameters_equal_oracle_parameters():
    """Two independent per-unit conversions (product host code vs oracle) agree bit for bit."""
    import gym_pvder_b200 as G

    for der in ("10", "50"):
        par, ex = G.load_der_parameters(der)
        o = load_der_params(der)
        assert (par.Rf, par.Rt, par.Xt, par.vgs, par.Vrms_ref, par.iref_limit) == (o.Rf, o.Rt, o.Xt, o.vgs, o.Vrms_ref, o.iref_limit)
        assert par.inv_Lf == 1.0 / o.Lf and par.inv_C == 1.0 / o.C and par.kappa == o.kappa
        assert par.np_irs == o.Np * 1.2e-7 and ex["Vdc_ref0"] == o.Vdc_ref0
