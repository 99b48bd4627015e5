"""Vectorised PVDER-v0: N independent environments stepped by ONE CUDA launch.

The per-step contract is the reference's (gym_PVDER/envs/PVDER_env.py:138-196) applied
element-wise: ``step(actions[N]) -> (obs[N,11] f32, reward[N], done[N] bool, info)``; the env
state lives in two SoA device tensors and never leaves HBM.  Tensors are torch CUDA tensors whose
raw pointers are handed to libpvder_b200.so (C ABI, include/pvder_b200.h) on torch's current
stream -- torch is plumbing (memory + streams) only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _cabi
from ..config import EnvConfig
from ..spaces import Box, Discrete


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class PVDERVecEnv:
    metadata = {"render.modes": ["vector"]}
    observed_quantities = ["iaR", "iaI", "vaR", "vaI", "P_PCC", "Q_PCC", "Vdc", "Ppv", "Vdc_ref", "Q_ref", "tStart"]

    def __init__(self, num_envs, device="cuda", seed=0, env_offset=0, model_type="model_2",
                 n_sim_time_steps_per_env_step=15, max_sim_time=40.0, DISCRETE_REWARD=True,
                 goals_list=("voltage_regulation",), events_spec=None, event_mode="philox", auto_reset=False,
                 obs_f64=False, micro=1, balanced_three_phase="auto", grid_unbalance_ratio=(1.0, 1.0),
                 config=None, validate_actions=False, **config_kwargs):
        """config_kwargs: further EnvConfig fields (reward_list, der_id, config_file, refine_input_level,
        refine_on_action, startup_substeps, startup_level, max_episode_steps)."""
        import torch

        self.torch = torch
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("PVDERVecEnv needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PVDERVecEnv needs a CUDA device (sm_100a); there is no CPU fallback")
        if self.device.index is None:      # 'cuda' != 'cuda:0' for torch: compare like with like in step()
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._obs_f64 = bool(obs_f64)
        # True: step() checks every action against Discrete(5) and raises like the reference's assertion (PVDER_env.py:201)
        # -- one device reduction + sync per step, for debugging; False: a bad action sets PVDER_STATUS_BAD_ACTION for that
        # env (step skipped, cached reward returned) until its next valid step; poll with check_status()
        self.validate_actions = bool(validate_actions)
        self.cfg = config or EnvConfig(model_type=model_type,
                                       n_sim_time_steps_per_env_step=n_sim_time_steps_per_env_step,
                                       max_sim_time=max_sim_time, DISCRETE_REWARD=DISCRETE_REWARD,
                                       goals_list=list(goals_list), events_spec=events_spec,
                                       event_mode=event_mode, seed=seed, auto_reset=auto_reset, micro=micro,
                                       balanced_three_phase=balanced_three_phase,
                                       grid_unbalance_ratio=tuple(grid_unbalance_ratio), **config_kwargs)
        self.num_envs = int(num_envs)
        self.env_offset = int(env_offset)
        self.ns = self.cfg.n_state
        self.ld = (self.num_envs + 31) // 32 * 32
        self.action_space = Discrete(_cabi.N_ACTIONS)                             # PVDER_env.py:50
        self.observation_space = Box(-10, 10, (_cabi.OBS_DIM,), np.float32)       # PVDER_env.py:51
        with torch.cuda.device(self.device):
            dev = self.device
            self.sd = torch.zeros((_cabi.sd_fields(self.ns), self.ld), dtype=torch.float64, device=dev)
            self.si = torch.zeros((_cabi.SI_FIELDS, self.ld), dtype=torch.int32, device=dev)
            self.obs = torch.zeros((self.num_envs, _cabi.OBS_DIM), dtype=torch.float32, device=dev)
            self.obs64 = torch.zeros((self.num_envs, _cabi.OBS_DIM), dtype=torch.float64, device=dev) if obs_f64 else None
            self.reward = torch.zeros(self.num_envs, dtype=torch.float64, device=dev)
            self.reward_i = torch.zeros(self.num_envs, dtype=torch.int32, device=dev) if self.cfg.DISCRETE_REWARD else None
            self.done = torch.zeros(self.num_envs, dtype=torch.uint8, device=dev)
            self._actions = torch.zeros(self.num_envs, dtype=torch.int32, device=dev)
            self._stats = torch.zeros(16, dtype=torch.float64, device=dev)
            self.vgrid_tab = self.sinsol_tab = None
            if self.cfg.c.event_mode == _cabi.EVENT_MODES["table"]:
                k = max(1, self.cfg.c.ev_count)
                self.vgrid_tab = torch.ones((k, self.ld), dtype=torch.float64, device=dev)
                self.sinsol_tab = torch.full((k, self.ld), 100.0, dtype=torch.float64, device=dev)
        self._initialised = False
        self._step_index = 0
        self.launches = 0
        self.traj = None
        self._traj_stride = 1

    # ---- plumbing -------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _cfgp(self):
        return C.byref(self.cfg.c)

    # ---- API ------------------------------------------------------------------------------
    def reset(self, mask=None):
        """PVDER.reset (PVDER_env.py:316-334) for all envs, or those with mask != 0."""
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=self.torch.uint8).contiguous()
        with self.torch.cuda.device(self.device):
            _cabi.check(self.lib.pvder_reset(self._cfgp(), _ptr(self.sd), _ptr(self.si), self.ld, _ptr(m),
                                             0 if self._initialised else 1, _ptr(self.obs), _ptr(self.obs64),
                                             self.num_envs, self.env_offset, self._stream()))
        self._initialised = True
        self.launches += 1
        return self.obs

    def step(self, actions):
        """PVDER.step (PVDER_env.py:138-196) for every env; one kernel launch."""
        if not self._initialised:
            raise RuntimeError("Cannot call step() before reset()")
        t = self.torch
        if not isinstance(actions, t.Tensor):
            actions = t.as_tensor(np.asarray(actions), device=self.device)
        if actions.numel() != self.num_envs or actions.dim() > 2:     # checked BEFORE any copy: nothing is broadcast
            raise ValueError(f"actions must have num_envs = {self.num_envs} elements, got shape {tuple(actions.shape)}")
        if actions.dtype != t.int32 or actions.device != self.device or not actions.is_contiguous():
            self._actions.copy_(actions.to(self.device).reshape(self.num_envs))
            actions = self._actions
        if self.validate_actions and bool(((actions < 0) | (actions >= _cabi.N_ACTIONS)).any()):
            raise AssertionError("an action is not available in the environment action space!")      # PVDER_env.py:201
        with t.cuda.device(self.device):
            if self.traj is None:
                _cabi.check(self.lib.pvder_step(self._cfgp(), _ptr(self.sd), _ptr(self.si), self.ld, _ptr(actions),
                                                _ptr(self.vgrid_tab), _ptr(self.sinsol_tab), _ptr(self.obs),
                                                _ptr(self.obs64), _ptr(self.reward), _ptr(self.reward_i), _ptr(self.done),
                                                self.num_envs, self.env_offset, self._stream()))
            else:
                _cabi.check(self.lib.pvder_step_record(self._cfgp(), _ptr(self.sd), _ptr(self.si), self.ld, _ptr(actions),
                                                       _ptr(self.vgrid_tab), _ptr(self.sinsol_tab), _ptr(self.obs),
                                                       _ptr(self.obs64), _ptr(self.reward), _ptr(self.reward_i),
                                                       _ptr(self.done), self.num_envs, self.env_offset, _ptr(self.traj),
                                                       self.traj.shape[2], self._traj_stride, self._stream()))
        self.launches += 1
        self._step_index += 1
        reward = self.reward_i if self.cfg.DISCRETE_REWARD else self.reward
        return self.obs, reward, self.done.view(t.bool), {}

    def record_trajectory(self, n_envs=1, stride=1):
        """Record, from the next step() on, the state after every half-cycle sub-step of the envs
        0, stride, 2*stride, ... (n_envs of them) -- the per-sub-step time series the reference plots through
        pvder's SimulationResults (PVDER_env.py:358-364).  After each step ``self.traj`` holds
        [n_sub_per_step, n_state + 2, n_envs] (rows: state in stored order with the PLL angle as delta, then
        Vgrid and Sinsol in force).  ``record_trajectory(0)`` switches recording off."""
        t = self.torch
        if n_envs <= 0:
            self.traj = None
            return None
        if (n_envs - 1) * stride >= self.num_envs:
            raise ValueError("recorded envs exceed num_envs")
        self._traj_stride = int(stride)
        self.traj = t.zeros((self.cfg.c.n_sub_per_step, self.ns + 2, int(n_envs)), dtype=t.float64, device=self.device)
        return self.traj

    def sample_actions(self, out=None):
        """action_space.sample() for every env, on device (Philox stream 1)."""
        out = self._actions if out is None else out
        with self.torch.cuda.device(self.device):
            _cabi.check(self.lib.pvder_sample_actions(self.cfg.c.seed, self._step_index, _ptr(out), self.num_envs,
                                                      self.env_offset, self._stream()))
        self.launches += 1
        return out

    def set_event_tables(self, vgrid, sinsol):
        """event_mode='table': value in force from event instant j on, shape [ev_count, num_envs]."""
        t = self.torch
        self.vgrid_tab[:, :self.num_envs].copy_(t.as_tensor(vgrid, dtype=t.float64))
        self.sinsol_tab[:, :self.num_envs].copy_(t.as_tensor(sinsol, dtype=t.float64))

    def generate_events(self):
        """Materialise the Philox event tables (what the step kernel draws on the fly)."""
        t = self.torch
        k = max(1, self.cfg.c.ev_count)
        v = t.ones((k, self.ld), dtype=t.float64, device=self.device)
        s = t.full((k, self.ld), 100.0, dtype=t.float64, device=self.device)
        ep = self.si[_cabi.SI_EPISODE].contiguous()
        with t.cuda.device(self.device):
            _cabi.check(self.lib.pvder_generate_events(self._cfgp(), _ptr(ep), _ptr(v), _ptr(s), self.ld, self.num_envs,
                                                       self.env_offset, self._stream()))
        return v[:, :self.num_envs], s[:, :self.num_envs]

    def stats(self):
        """Episode statistics summed over this shard (env_utilities.py:12-46) as a device tensor[16]."""
        with self.torch.cuda.device(self.device):
            _cabi.check(self.lib.pvder_stats_reduce(_ptr(self.sd), _ptr(self.si), self.ld, self.cfg.phases, self.num_envs,
                                                    _ptr(self._stats), self._stream()))
        return self._stats

    # ---- checkpoint / resume ----------------------------------------------------------------
    def state_dict(self):
        """Everything needed to resume bit-identically: the two SoA state tensors (ODE states, references, event
        values, counters, RNG episode keys), the action-stream position and the last outputs."""
        d = {"sd": self.sd.clone(), "si": self.si.clone(), "step_index": int(self._step_index),
             "obs": self.obs.clone(), "done": self.done.clone(), "num_envs": self.num_envs, "env_offset": self.env_offset,
             "model_type": self.cfg.model_type, "seed": int(self.cfg.seed)}
        if self.vgrid_tab is not None:
            d["vgrid_tab"], d["sinsol_tab"] = self.vgrid_tab.clone(), self.sinsol_tab.clone()
        return d

    def load_state_dict(self, d):
        if (d["num_envs"], d["env_offset"], d["model_type"]) != (self.num_envs, self.env_offset, self.cfg.model_type):
            raise ValueError("checkpoint was taken from a different shard / model")
        if int(d["seed"]) != int(self.cfg.seed):
            raise ValueError("checkpoint was taken with another seed (event and action streams would differ)")
        self.sd.copy_(d["sd"])
        self.si.copy_(d["si"])
        self.obs.copy_(d["obs"])
        self.done.copy_(d["done"])
        if self.vgrid_tab is not None and "vgrid_tab" in d:
            self.vgrid_tab.copy_(d["vgrid_tab"])
            self.sinsol_tab.copy_(d["sinsol_tab"])
        self._step_index = int(d["step_index"])
        self._initialised = True

    # ---- state views ------------------------------------------------------------------------
    @property
    def y(self):
        return self.sd[:self.ns, :self.num_envs]

    def field(self, name):
        return self.sd[_cabi.sd_index(self.ns)[name], :self.num_envs]

    @property
    def k(self):
        return self.si[_cabi.SI_K, :self.num_envs]

    @property
    def steps(self):
        return self.si[_cabi.SI_STEPS, :self.num_envs]

    @property
    def status(self):
        return self.si[_cabi.SI_STATUS, :self.num_envs]

    def check_status(self):
        """Raise like the reference does (AssertionError, PVDER_env.py:177/201) if any env failed."""
        st = self.status
        if bool((st == _cabi.STATUS_BAD_ACTION).any()):
            raise AssertionError("an action is not available in the environment action space!")
        if bool((st >= _cabi.STATUS_NONFINITE).any()):
            raise AssertionError("Convergence flag should be true to calculate reward!")

    def close(self):
        pass

    # ---- goals (PVDER_env.py:445-456) and the return sweep (:458-497) -----------------------
    def update_env_goal(self, goal_type=None, goal_spec=None):
        """PVDER.update_env_goal.  (None, None) installs every goal's required reward/action terms, like the reference
        (:447-452).  Otherwise ``goal_spec = {"reward": [...]}`` selects the reward list (`my_spec`, :249) of
        ``goal_type`` -- its required term plus optional ones (env_goal_spec :78-93) -- the part the reference left
        "under construction" (:454-456); it takes effect at once if goal_type is the current goal."""
        from ..config import GOAL_SPEC, validate_reward_list

        if goal_type is None and goal_spec is None:
            self._goal_rewards = {}
            reward_list = None
        else:
            if goal_type not in GOAL_SPEC:
                raise ValueError("{} is not a valid goal!".format(goal_type))
            extra = set(goal_spec or {}) - {"reward"}
            if extra:
                raise ValueError("{} is not a valid goal spec entry!".format(sorted(extra)))
            reward_list = validate_reward_list(goal_type, (goal_spec or {}).get("reward"))
            if not hasattr(self, "_goal_rewards"):
                self._goal_rewards = {}
            self._goal_rewards[goal_type] = reward_list
            if goal_type != self.cfg.goals_list[0]:
                return
        self._set_goal(self.cfg.goals_list[0], reward_list)

    def _set_goal(self, goal, reward_list=None):
        import dataclasses

        if reward_list is None:
            reward_list = getattr(self, "_goal_rewards", {}).get(goal)
        fields = {f.name: getattr(self.cfg, f.name) for f in dataclasses.fields(self.cfg)}
        fields.update(goals_list=[goal], reward_list=reward_list)
        self.cfg = EnvConfig(**fields)

    def calc_returns(self, n_episodes=2, action_specs=("random", "inc", "dec", "no_change"), goals=None):
        """PVDER.calc_returns (PVDER_env.py:458-497) as a BATCHED evaluation: for every goal, the fixed policies
        'random' / 'inc' / 'dec' / 'no_change' x n_episodes run as ONE vector env (env j = spec j // n_episodes, episode
        j % n_episodes), one launch per env step, so the whole sweep costs 3 x episode_steps launches instead of
        3 x 4 x n_episodes serial episodes.  Same action mapping as the reference ('inc' -> 0, 'dec' -> 1, 'no_change'
        -> 2, :480-486).  Env j draws the events and random actions of GLOBAL env index env_offset + j, so a serial
        N = 1 run with env_offset = j reproduces its number.  Returns {goal: {spec: {'return', 'ref'}}} and stores it in
        ``env_average_return``."""
        from ..config import GOAL_SPEC
        import dataclasses

        t = self.torch
        specs = list(action_specs)
        fixed = {"inc": 0, "dec": 1, "no_change": 2}
        for sp in specs:
            if sp != "random" and sp not in fixed:
                raise ValueError(f"unknown action spec {sp!r}")
        n = len(specs) * int(n_episodes)
        ex = self.cfg.extras
        out = {}
        for goal in (goals or GOAL_SPEC):
            fields = {f.name: getattr(self.cfg, f.name) for f in dataclasses.fields(self.cfg)}
            fields.update(goals_list=[goal], reward_list=getattr(self, "_goal_rewards", {}).get(goal), auto_reset=False)
            env = PVDERVecEnv(n, device=self.device, env_offset=self.env_offset, config=EnvConfig(**fields))
            env.reset()
            acts = t.zeros(n, dtype=t.int32, device=self.device)
            rand = t.zeros(n, dtype=t.int32, device=self.device)
            is_random = t.tensor([specs[j // n_episodes] == "random" for j in range(n)], device=self.device)
            const = t.tensor([fixed.get(specs[j // n_episodes], 0) for j in range(n)], dtype=t.int32, device=self.device)
            ret = t.zeros(n, dtype=t.float64, device=self.device)
            for _ in range(env.cfg.episode_steps):
                env.sample_actions(rand)
                t.where(is_random, rand, const, out=acts)
                _, rew, done, _ = env.step(acts)
                ret += rew.to(t.float64)
            assert bool(done.all())
            env.check_status()
            ret = ret.view(len(specs), n_episodes).mean(dim=1).cpu().numpy()
            vdc = (env.field("Vdc_ref").view(len(specs), n_episodes)[:, -1] * ex["Vbase"]).cpu().numpy()
            q = (env.field("Q_ref").view(len(specs), n_episodes)[:, -1] * ex["Sbase"]).cpu().numpy()
            out[goal] = {sp: {"return": float(ret[i]), "ref": [float(vdc[i]), float(q[i])]} for i, sp in enumerate(specs)}
        self.env_average_return = out
        return out
