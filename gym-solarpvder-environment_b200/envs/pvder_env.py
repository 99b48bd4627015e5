"""Single-environment facade with the reference's Gym contract (old API).

``PVDER`` mirrors reference gym_PVDER/envs/PVDER_env.py: class-level specs 44-102, validated
kwargs 561-620, ``reset`` 316-334, ``step`` 138-196, ``render`` 336-364, ``update_env_events``
413-438, ``state`` 531-542.  It is an N = 1 view of the CUDA path: every call goes through the
host-buffer C ABI (``pvder_env_step_host``: action host->device, one kernel launch, obs/reward/
done device->host).  Python types follow the reference's tests: reward is ``int`` in discrete
mode and ``float`` otherwise, ``done`` is ``bool`` (gym_PVDER/tests/test_gym_PVDER.py:34-35,66,81).
"""
from __future__ import annotations

import copy
import ctypes as C
import logging
import os
import time
import types

import numpy as np

from .. import _cabi
from .. import config as cfgmod
from ..config import EnvConfig
from ..spaces import Box, Discrete
from .env_utilities import Utilities


class PVDER(Utilities):
    count = 0
    metadata = {"render.modes": ["vector", "human"]}
    observed_quantities = ["iaR", "iaI", "vaR", "vaI", "P_PCC", "Q_PCC", "Vdc", "Ppv", "Vdc_ref", "Q_ref", "tStart"]
    action_space = Discrete(5)
    observation_space = Box(low=-10, high=10, shape=(len(observed_quantities),), dtype=np.float32)

    env_model_spec = cfgmod.MODEL_SPEC
    env_sim_spec = {"sim_time_step": cfgmod.SIM_TIME_STEP,
                    "n_sim_time_steps_per_env_step": {"default": cfgmod.N_SIM_DEFAULT, "min": cfgmod.N_SIM_MIN},
                    "min_sim_time": cfgmod.MIN_SIM_TIME}
    env_reward_spec = {"reward_list": {"default": ["voltage_error"],
                                       "valid": ["voltage_error", "power_error", "Q_error", "Vdc_error"]},
                       **cfgmod.REWARD_SPEC}
    env_action_spec = {"action_list": {"default": ["Q_control"], "valid": ["Q_control", "Vdc_control"]},
                       **cfgmod.ACTION_SPEC}
    env_goal_spec = cfgmod.GOAL_SPEC   # class-level default; every instance works on its own copy
    default_goal = cfgmod.DEFAULT_GOAL

    def __init__(self, goals_list=None, n_sim_time_steps_per_env_step=None, max_sim_time=None, DISCRETE_REWARD=None,
                 verbosity="INFO", model_type="model_2", seed=None, spec=None):
        PVDER.count += 1
        self.name = "GymDER_" + str(PVDER.count)
        self.logger = logging.getLogger(self.name)
        self.logger.setLevel(getattr(logging, verbosity, logging.INFO))
        self.spec = spec if spec is not None else types.SimpleNamespace(id="PVDER-v0", max_episode_steps=500)
        self.verbosity = verbosity
        self.model_type = model_type
        self.env_events_spec = copy.deepcopy(cfgmod.DEFAULT_EVENTS_SPEC)   # per instance
        self.n_sim_time_steps_per_env_step = n_sim_time_steps_per_env_step
        self.max_sim_time_user = max_sim_time
        self.max_sim_time = max_sim_time
        self.goals_list = goals_list if goals_list is not None else list(self.default_goal)
        self.DISCRETE_REWARD = DISCRETE_REWARD
        self._seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)  # reference RNG is unseeded
        self._handle = None
        self._handle_key = None
        # per instance (the reference mutates a class-level dict); my_spec = required (PVDER_env.py:445-452)
        self.env_goal_spec = copy.deepcopy(cfgmod.GOAL_SPEC)
        self.update_env_goal(None, None)
        self._obs = np.zeros((1, _cabi.OBS_DIM), dtype=np.float32)
        self._obs64 = np.zeros((1, _cabi.OBS_DIM), dtype=np.float64)
        self._rew = np.zeros(1, dtype=np.float64)
        self._done_buf = np.zeros(1, dtype=np.uint8)
        self._act = np.zeros(1, dtype=np.int32)
        self.initialize_environment_variables()

    # ---- validated properties (PVDER_env.py:544-620) ----------------------------------------
    @property
    def n_sim_time_steps_per_env_step(self):
        return self.__n

    @n_sim_time_steps_per_env_step.setter
    def n_sim_time_steps_per_env_step(self, n):
        self.__n = cfgmod.validate_n_sim(n)
        self._sim_time_per_env_step = cfgmod.SIM_TIME_STEP * self.__n
        self._delQref = cfgmod.ACTION_SPEC["delQref"] * self.__n
        self._delVdcref = cfgmod.ACTION_SPEC["delVdcref"] * self.__n

    @property
    def max_sim_time(self):
        return self.__max_sim_time

    @max_sim_time.setter
    def max_sim_time(self, t):
        self.__max_sim_time = cfgmod.validate_max_sim_time(t, self.__n, self.spec.max_episode_steps)

    @property
    def goals_list(self):
        return self.__goals_list

    @goals_list.setter
    def goals_list(self, goals):
        self.__goals_list = cfgmod.validate_goals(goals)

    @property
    def DISCRETE_REWARD(self):
        return self.__discrete

    @DISCRETE_REWARD.setter
    def DISCRETE_REWARD(self, flag):
        self.__discrete = cfgmod.validate_discrete(flag)

    @property
    def unwrapped(self):
        return self

    # ---- lifecycle -----------------------------------------------------------------------
    def initialize_environment_variables(self):
        """PVDER_env.py:303-314."""
        self._steps = 0
        self._reward = 0
        self._step_time = 0.0
        self.done = False
        self.CONVERGENCE_FAILURE = False
        self.RUNTIME_ERROR = False
        self.initialize_stats()

    def _destroy(self):
        if self._handle is not None:
            _cabi.load().pvder_env_destroy(self._handle)
            self._handle = None

    def setup_PVDER_simulation(self, model_type=None):
        """PVDER_env.py:366-398: a fresh simulator at t = 0 with new random events.  The device handle (streams,
        buffers) is created once per model and re-configured afterwards -- new seed, goal, reward list, event ranges --
        instead of being destroyed and re-created at every reset."""
        lib = _cabi.load()
        self.max_sim_time = self.max_sim_time_user
        self._episode_seed = (self._seed + 0x9E3779B97F4A7C15 * (self._episodes + 1)) & 0xFFFFFFFFFFFFFFFF
        goal = self.goals_list[0]
        self.config = EnvConfig(model_type=model_type or self.model_type,
                                n_sim_time_steps_per_env_step=self.n_sim_time_steps_per_env_step,
                                max_sim_time=self.max_sim_time, DISCRETE_REWARD=self.DISCRETE_REWARD,
                                goals_list=self.goals_list, reward_list=self.env_goal_spec[goal]["reward"]["my_spec"],
                                events_spec=self.env_events_spec, event_mode="philox",
                                seed=self._episode_seed, max_episode_steps=self.spec.max_episode_steps)
        key = (self.config.phases, self.config.c.balanced3)
        if self._handle is not None and self._handle_key == key:
            _cabi.check(lib.pvder_env_reconfigure(self._handle, C.byref(self.config.c)))
        else:
            self._destroy()
            h = C.c_void_p()
            _cabi.check(lib.pvder_env_create(C.byref(self.config.c), 1, 0, C.byref(h)))
            self._handle, self._handle_key = h, key
        ex = self.config.extras
        self.sim = types.SimpleNamespace(
            name="DER_sim_b200", tInc=cfgmod.SIM_TIME_STEP, tStart=0.0, tStop=0.0, Sbase=ex["Sbase"], Vbase=ex["Vbase"],
            Ibase=ex["Ibase"], SOLVER_CONVERGENCE=True,
            simulation_events=types.SimpleNamespace(_events_spec=self.env_events_spec),
            PV_model=types.SimpleNamespace(Sbase=ex["Sbase"], Vdcbase=ex["Vdcbase"], Q_ref=0.0, Vdc_ref=ex["Vdc_ref0"],
                                           Vrms_ref=self.config.par.Vrms_ref))

    _episodes = 0

    def reset(self):
        """PVDER_env.py:316-334."""
        self.initialize_environment_variables()
        self.setup_PVDER_simulation()
        self._episodes += 1
        _cabi.check(_cabi.load().pvder_env_reset_host(self._handle, self._obs.ctypes.data, self._obs64.ctypes.data))
        self._sync_sim()
        return self._obs[0].copy()

    def step(self, action):
        """PVDER_env.py:138-196."""
        time_start = time.time()
        if self._handle is None:
            raise RuntimeError("Cannot call env.step() before calling reset()")
        if self.done:
            self.logger.warning("%s:Simulation completed - Reset environment to start new simulation!", self.name)
            return self._obs[0].copy(), self._reward, self.done, {}
        assert action in self.action_space, \
            'The action "{}" is not available in the environment action space!'.format(action)   # :201
        self.update_action_stats(int(action))
        self._steps += 1
        self._act[0] = int(action)
        _cabi.check(_cabi.load().pvder_env_step_host(self._handle, self._act.ctypes.data, self._obs.ctypes.data,
                                                     self._obs64.ctypes.data, self._rew.ctypes.data,
                                                     self._done_buf.ctypes.data))
        self.sim.tStop = self.sim.tStart + self._sim_time_per_env_step
        # a failed integration is reported by the kernel as reward -100 + done (PVDER_env.py:170-172); only then
        # is the status word fetched (one synchronous copy less per step)
        status = _cabi.STATUS_OK
        if self._done_buf[0] and self._rew[0] == -100.0:
            status = int(self._state_i32()[_cabi.SI_STATUS])
        # NONFINITE: the integrator failed.  UNBALANCED cannot occur here (the facade runs the three-phase model in
        # 'auto' mode, which hands such an env to the general model), but is surfaced the same way if it ever does.
        self.CONVERGENCE_FAILURE = status in (_cabi.STATUS_NONFINITE, _cabi.STATUS_UNBALANCED)
        assert not self.CONVERGENCE_FAILURE, "Convergence flag should be true to calculate reward!"   # :177
        self._reward = int(self._rew[0]) if self.DISCRETE_REWARD else float(self._rew[0])
        self.sim.tStart = self.sim.tStop
        self.update_reward_stats()
        self.done = bool(self._done_buf[0])
        self._sync_sim()
        self._step_time = time.time() - time_start
        self.update_time_stats()
        return self._obs[0].copy(), self._reward, self.done, {}

    def _state_i32(self):
        si = np.zeros(_cabi.SI_FIELDS, dtype=np.int32)
        _cabi.check(_cabi.load().pvder_env_state_host(self._handle, None, si.ctypes.data))
        return si

    def _sync_sim(self):
        o = self._obs64[0]
        pv = self.sim.PV_model
        pv.ia, pv.va, pv.S_PCC = complex(o[0], o[1]), complex(o[2], o[3]), complex(o[4], o[5])
        pv.Vdc, pv.Ppv, pv.Vdc_ref, pv.Q_ref = float(o[6]), float(o[7]), float(o[8]), float(o[9])
        self.sim.tStart = float(o[10]) * self.max_sim_time
        self.sim.tStop = self.sim.tStart

    @property
    def state(self):
        """PVDER_env.py:531-542 (float64 tuple of the 11 observed quantities)."""
        return tuple(float(v) for v in self._obs64[0])

    def render(self, mode="vector"):
        """PVDER_env.py:336-364: print de-normalised observations; 'human' plots are out of scope."""
        pv, sim = self.sim.PV_model, self.sim
        items = {"ia": pv.ia * sim.Ibase, "Vdc": pv.Vdc * pv.Vdcbase, "va": pv.va * sim.Vbase,
                 "Ppv": pv.Ppv * sim.Sbase, "S_PCC": pv.S_PCC * sim.Sbase, "Q_ref": pv.Q_ref * sim.Sbase,
                 "Vdc_ref": pv.Vdc_ref * pv.Vdcbase, "tStart": sim.tStart}
        for k, v in items.items():
            print("{}:{:.2f},".format(k, v), end=" ")
        print("\nReward:{:.5f}".format(self._reward))
        self.show_step_time()

    def update_env_events(self, event_spec_list):
        """PVDER_env.py:413-438 (per-instance spec; takes effect at the next reset)."""
        assert isinstance(event_spec_list, list), "event_spec_list should be a list!"
        for event_spec in event_spec_list:
            assert isinstance(event_spec, dict), "Event spec should be a dictionary!"
            assert len(event_spec.keys()) == 1, "Only one event type should be specified at at time!"
        merged = {}
        for event_spec in event_spec_list:
            for kind, params in event_spec.items():
                merged.setdefault(kind, {}).update(params)
        current = copy.deepcopy(self.env_events_spec)
        for kind, params in merged.items():
            if kind not in current:
                raise ValueError("{} is not a valid event!".format(kind))
            for k, v in params.items():
                if k not in current[kind]:
                    raise ValueError("{} is not a valid paramter for {} event!".format(k, kind))
                current[kind][k] = v
        self.env_events_spec.clear()
        self.env_events_spec.update(current)

    def update_env_goal(self, goal_type=None, goal_spec=None):
        """PVDER_env.py:445-456.  (None, None): every goal's `my_spec` = its required reward/action terms, as in the
        reference.  Otherwise ``goal_spec = {"reward": [...]}`` sets the reward list of ``goal_type`` to its required
        term plus the chosen optional ones (env_goal_spec, :78-93) -- the branch the reference left "under
        construction"; it takes effect at the next reset(), when the simulator is rebuilt."""
        if goal_type is None and goal_spec is None:
            for g in self.env_goal_spec:
                self.env_goal_spec[g]["reward"]["my_spec"] = list(self.env_goal_spec[g]["reward"]["required"])
                self.env_goal_spec[g]["action"]["my_spec"] = list(self.env_goal_spec[g]["action"]["required"])
            return
        if goal_type not in self.env_goal_spec:
            raise ValueError("{} is not a valid goal!".format(goal_type))
        extra = set(goal_spec or {}) - {"reward"}
        if extra:
            raise ValueError("{} is not a valid goal spec entry!".format(sorted(extra)))
        self.env_goal_spec[goal_type]["reward"]["my_spec"] = cfgmod.validate_reward_list(goal_type,
                                                                                          (goal_spec or {}).get("reward"))

    def calc_returns(self, n_episodes=2, action_specs=("random", "inc", "dec", "no_change"), batched=True):
        """Average return of fixed policies for every goal (reference PVDER_env.py:458-497, same
        action mapping: 'inc' -> action 0, 'dec' -> 1, 'no_change' -> 2).  batched (default): the whole sweep runs as
        vector rollouts on the device (PVDERVecEnv.calc_returns: one launch per env step for all policies and episodes
        of a goal); batched=False: the reference's serial loop through this env's own reset()/step()."""
        if batched:
            from .vec_env import PVDERVecEnv

            venv = PVDERVecEnv(1, seed=self._seed, model_type=self.model_type,
                               n_sim_time_steps_per_env_step=self.n_sim_time_steps_per_env_step,
                               max_sim_time=self.max_sim_time_user, DISCRETE_REWARD=self.DISCRETE_REWARD,
                               goals_list=self.goals_list, events_spec=self.env_events_spec,
                               max_episode_steps=self.spec.max_episode_steps)
            venv._goal_rewards = {g: list(self.env_goal_spec[g]["reward"]["my_spec"]) for g in self.env_goal_spec}
            self.env_average_return = venv.calc_returns(n_episodes=n_episodes, action_specs=action_specs)
            self.pp.pprint(self.env_average_return)
            return self.env_average_return
        self.env_average_return = {}
        saved = self.goals_list
        for goal in self.env_goal_spec:
            self.env_average_return[goal] = {}
            self.goals_list = [goal]
            for spec in action_specs:
                total = 0.0
                for _ in range(n_episodes):
                    self.reset()
                    done, ret = False, 0.0
                    while not done:
                        action = {"random": None, "inc": 0, "dec": 1, "no_change": 2}[spec]
                        if action is None:
                            action = self.action_space.sample()
                        _, reward, done, _ = self.step(action)
                        ret += reward
                    total += ret
                pv = self.sim.PV_model
                self.env_average_return[goal][spec] = {"return": total / n_episodes,
                                                       "ref": [pv.Vdc_ref * self.sim.Vbase, pv.Q_ref * self.sim.Sbase]}
        self.goals_list = saved
        self.pp.pprint(self.env_average_return)
        return self.env_average_return

    def seed(self, seed=None):
        self._seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        return [self._seed]

    def close(self):
        self._destroy()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass
