"""Tiny environment registry with the old-gym surface the reference relies on:
``register`` / ``make`` / ``spec`` and the ``TimeLimit`` wrapper that ``gym.make`` puts around
the env (reference gym_PVDER/__init__.py:3-10; tests use env.spec.id, env.unwrapped and
attribute pass-through such as env.steps / env.sim, gym_PVDER/tests/test_gym_PVDER.py:11-14,
:108).  When a real ``gym`` is importable the id is registered there as well."""
from __future__ import annotations

import copy


class EnvSpec:
    def __init__(self, id, entry_point, kwargs=None, max_episode_steps=None):
        self.id = id
        self.entry_point = entry_point
        self.kwargs = dict(kwargs or {})
        self.max_episode_steps = max_episode_steps

    def make(self, **kwargs):
        kw = copy.deepcopy(self.kwargs)
        kw.update(kwargs)
        env = self.entry_point(**kw)           # like gym.make: the entry point only sees its kwargs ...
        env.spec = self                        # ... and gym attaches the spec afterwards
        if hasattr(env, "max_sim_time_user"):  # its clamp depends on spec.max_episode_steps (PVDER_env.py:565-573)
            env.max_sim_time = env.max_sim_time_user
        if self.max_episode_steps is not None:
            env = TimeLimit(env, self.max_episode_steps)
        return env

    def __repr__(self):
        return f"EnvSpec({self.id})"


class TimeLimit:
    """gym.wrappers.TimeLimit (old API: 4-tuple step)."""

    def __init__(self, env, max_episode_steps):
        self.env = env
        self._max_episode_steps = int(max_episode_steps)
        self._elapsed_steps = None

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def spec(self):
        return self.env.spec

    def step(self, action):
        assert self._elapsed_steps is not None, "Cannot call env.step() before calling reset()"
        observation, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return observation, reward, done, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)

    def render(self, mode="vector", **kwargs):
        return self.env.render(mode=mode, **kwargs)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)


registry: dict[str, EnvSpec] = {}


def register(id, entry_point, kwargs=None, max_episode_steps=None):
    registry[id] = EnvSpec(id, entry_point, kwargs, max_episode_steps)
    # also expose the id through a real gym / gymnasium, when one is importable (neither is in this image)
    for modname in ("gym", "gymnasium"):
        try:
            mod = __import__(modname)
            known = getattr(mod.envs.registry, "env_specs", mod.envs.registry)
            if id not in known:
                mod.register(id=id, entry_point=entry_point, kwargs=dict(kwargs or {}), max_episode_steps=max_episode_steps)
        except ImportError:
            continue
        except Exception as exc:      # a registry API this shim does not know: say so instead of hiding it
            import warnings

            warnings.warn(f"could not register {id} with {modname}: {exc}")
    return registry[id]


def spec(id):
    if id not in registry:
        raise KeyError(f"No registered env with id: {id}")
    return registry[id]


def make(id, **kwargs):
    return spec(id).make(**kwargs)
