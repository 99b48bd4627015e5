"""The reference's own 7 API tests (gym_PVDER/tests/test_gym_PVDER.py) re-expressed against the
drop-in package: same calls, same assertions, old-gym API."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gym(cuda):
    import gym_pvder_b200 as G

    return G


def test_make(gym):                                   # reference tests:11-14
    env = gym.make("PVDER-v0")
    assert env.spec.id == "PVDER-v0"
    assert isinstance(env.unwrapped, gym.PVDER)


def test_env(gym, capsys):                            # reference tests:17-40
    env = gym.spec("PVDER-v0").make()
    ob_space, act_space = env.observation_space, env.action_space
    ob = env.reset()
    assert ob_space.contains(ob), "Reset observation: {!r} not in space".format(ob)
    a = act_space.sample()
    observation, reward, done, _info = env.step(a)
    assert ob_space.contains(observation)
    assert np.isscalar(reward)
    assert isinstance(done, bool)
    for mode in env.metadata.get("render.modes", []):
        env.render(mode=mode)
    assert "Reward:" in capsys.readouterr().out
    env.close()


def test_random_rollout(gym):                         # reference tests:43-54
    env = gym.make("PVDER-v0")
    ob = env.reset()
    for _ in range(10):
        assert env.observation_space.contains(ob)
        a = env.action_space.sample()
        assert env.action_space.contains(a)
        ob, _reward, done, _info = env.step(a)
        if done:
            break
    env.close()


def test_discrete_reward(gym):                        # reference tests:57-69
    env = gym.make("PVDER-v0", DISCRETE_REWARD=True, goals_list=["voltage_regulation"])
    ob = env.reset()
    for _ in range(10):
        ob, _reward, done, _info = env.step(env.action_space.sample())
        assert isinstance(_reward, int), "Reward should be discrete if DISCRETE_REWARD is True"
    env.close()


def test_continuous_reward(gym):                      # reference tests:72-84
    env = gym.make("PVDER-v0", DISCRETE_REWARD=False, goals_list=["voltage_regulation"])
    ob = env.reset()
    for _ in range(10):
        ob, _reward, done, _info = env.step(env.action_space.sample())
        assert isinstance(_reward, float), "Reward should be float if DISCRETE_REWARD is False"
    env.close()


def test_time_steps(gym):                             # reference tests:87-111
    _max_sim_time, _n = 25.0, 10
    env = gym.make("PVDER-v0", n_sim_time_steps_per_env_step=_n, max_sim_time=_max_sim_time)
    ob = env.reset()
    assert env.n_sim_time_steps_per_env_step == _n
    assert env.max_sim_time == _max_sim_time
    done, steps = False, 0
    while not done:
        assert env.observation_space.contains(ob)
        ob, _reward, done, _info = env.step(env.action_space.sample())
        steps += 1
    assert round(env.steps * env.n_sim_time_steps_per_env_step * env.sim.tInc, 6) == round(env.max_sim_time, 6)
    assert round(ob[-1] * env.max_sim_time, 6) == round(env.max_sim_time, 6)
    # step after done returns the cached tuple (PVDER_env.py:145-152, 196)
    ob2, r2, d2, _ = env.step(0)
    assert d2 and r2 == _reward and np.array_equal(ob2, ob) and env.steps == steps
    env.close()


def test_update_env_events(gym):                      # reference tests:114-133
    new_spec = {"voltage": {"min": 0.95}}
    env = gym.make("PVDER-v0")
    env.update_env_events(event_spec_list=[new_spec])
    ob = env.reset()
    assert env.env_events_spec["voltage"]["min"] == 0.95
    assert env.sim.simulation_events._events_spec["voltage"]["min"] == 0.95
    assert env.config.c.ev_v_min == 0.95
    for _ in range(10):
        assert env.observation_space.contains(ob)
        ob, _reward, done, _info = env.step(env.action_space.sample())
    with pytest.raises(ValueError):
        env.update_env_events([{"voltage": {"bogus": 1}}])
    with pytest.raises(ValueError):
        env.update_env_events([{"frequency": {"min": 1}}])
    env.close()


def test_single_env_matches_vector_env(gym, cuda):
    """The facade is an N = 1 view of the same kernels."""
    import torch

    env = gym.PVDER(goals_list=["voltage_regulation"], n_sim_time_steps_per_env_step=15, max_sim_time=40.0,
                    DISCRETE_REWARD=True, model_type="model_2", seed=11)
    ob = env.reset()
    v = gym.PVDERVecEnv(1, device=cuda, config=env.config)
    ov = v.reset()
    assert np.array_equal(ov.cpu().numpy()[0], ob)
    for a in [1, 3, 0, 2, 4, 1]:
        ob, r, d, _ = env.step(a)
        o2, r2, d2, _ = v.step(torch.tensor([a], dtype=torch.int32, device=cuda))
        assert np.array_equal(o2.cpu().numpy()[0], ob) and int(r2[0]) == r and bool(d2[0]) == d
    with pytest.raises(AssertionError):
        env.step(9)
    env.close()


def test_calc_returns(gym, capsys):                   # reference PVDER_env.py:458-497
    env = gym.PVDER(goals_list=["voltage_regulation"], n_sim_time_steps_per_env_step=60, max_sim_time=3.0,
                    DISCRETE_REWARD=True, model_type="model_2", seed=2)
    res = env.calc_returns(n_episodes=1)
    assert set(res) == {"voltage_regulation", "power_regulation", "Q_regulation"}
    assert set(res["Q_regulation"]) == {"random", "inc", "dec", "no_change"}
    assert res["Q_regulation"]["no_change"]["return"] == -15.0      # 3 steps x (-5): Q is far from 5.5 kVAR
    env.close()
