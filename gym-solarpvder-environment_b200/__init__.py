"""B200-native batched PVDER-v0 environment (drop-in for the hot path of
sibyjackgrove/gym-SolarPVDER-environment: reference gym_PVDER/__init__.py:3-10 registration,
gym_PVDER/envs/PVDER_env.py reset/step).

    import gym_pvder_b200 as gym_PVDER
    env = gym_PVDER.make('PVDER-v0')                 # single env, old-gym API, numpy in/out
    venv = gym_PVDER.PVDERVecEnv(num_envs=1 << 20)   # N envs, torch CUDA tensors, one launch/step

The compute path is the hand-written sm_100a library csrc/libpvder_b200.so behind the C ABI of
include/pvder_b200.h; there is no CPU fallback (import works without a GPU, stepping does not).
"""
from .registration import make, register, spec, registry, EnvSpec, TimeLimit  # noqa: F401
from .spaces import Box, Discrete  # noqa: F401
from .config import (EnvConfig, load_der_parameters, DEFAULT_EVENTS_SPEC, MODEL_SPEC)  # noqa: F401
from .envs.pvder_env import PVDER  # noqa: F401
from .envs.vec_env import PVDERVecEnv  # noqa: F401
from . import _cabi  # noqa: F401

register(
    id="PVDER-v0",
    entry_point=PVDER,
    kwargs={"n_sim_time_steps_per_env_step": 15,      # reference gym_PVDER/__init__.py:6-9
            "max_sim_time": 40.0,
            "DISCRETE_REWARD": True,
            "goals_list": ["voltage_regulation"]},
    max_episode_steps=500,
)

__version__ = "0.1.0"
