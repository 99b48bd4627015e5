"""Host-side configuration of the PVDER-v0 environment.

Mirrors the class-level specs of reference gym_PVDER/envs/PVDER_env.py:44-102 and the kwargs
validation of :561-620, converts one derId of the DER parameter file to per-unit
 and packs everything the kernels need into the C struct ``pvder_env_config``.
"""
from __future__ import annotations

import copy
import ctypes as C
import json
import math
import os
from dataclasses import dataclass, field

from . import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
DER_CONFIG_FILE = os.path.join(_HERE, "der_config.json")

SUBSTEPS_PER_SEC = 120                       # half-cycle grid (README.md:11)
SIM_TIME_STEP = 1.0 / 60.0                   # env_sim_spec['sim_time_step'], PVDER_env.py:63
N_SIM_DEFAULT, N_SIM_MIN = 60, 1             # PVDER_env.py:64
MIN_SIM_TIME = 1.0                           # PVDER_env.py:65
MODEL_SPEC = {"model_1": {"DERModelType": "SinglePhase", "derId": "10"},            # PVDER_env.py:56-58
              "model_2": {"DERModelType": "ThreePhaseUnbalanced", "derId": "50"}}
DEFAULT_EVENTS_SPEC = {                      # PVDER_env.py:60-61
    "insolation": {"t_events_start": 1.0, "t_events_stop": 39.0, "t_events_step": 1.0,
                   "min": 85.0, "max": 100.0, "ENABLE": False},
    "voltage": {"t_events_start": 1.0, "t_events_stop": 39.0, "t_events_step": 1.0,
                "min": 0.98, "max": 1.02, "ENABLE": True}}
REWARD_SPEC = {"reference_values": {"P_ref": 45.4e3, "Q_ref": 5.5e3}, "DISCRETE_REWARD": True}   # :67-71
ACTION_SPEC = {"delQref": 25, "delVdcref": 0.02}                                                # :73-75
GOAL_SPEC = {                                                                                   # :78-93
    "voltage_regulation": {"reward": {"required": ["voltage_error"], "optional": ["Q_error", "Vdc_error"]},
                           "action": {"required": ["Q_control"], "optional": ["Vdc_control"]}},
    "power_regulation": {"reward": {"required": ["power_error"], "optional": ["Vdc_error"]},
                         "action": {"required": ["Vdc_control"], "optional": []}},
    "Q_regulation": {"reward": {"required": ["Q_error"], "optional": ["Vdc_error"]},
                     "action": {"required": ["Q_control"], "optional": []}}}
DEFAULT_GOAL = ["voltage_regulation"]

# PV module constants
_ISCR, _KV, _T0, _IRS, _QE, _KB, _A = 8.03, 0.0017, 298.15, 1.2e-7, 1.602e-19, 1.38e-23, 1.92
VBASE, SBASE = 500.0, 50e3
VGRID_RATED = 20415.0
Z2_ACTUAL = complex(1.61, 5.54)


_REFERENCE_SECTIONS = ("basic_specs", "basic_options", "module_parameters", "inverter_ratings", "circuit_parameters",
                       "controller_gains", "steadystate_values", "initial_states")


def read_der_config(der_id, config_file=None):
    """One DER parameter set as a flat dict.  Accepts this project's flat ``der_config.json`` and the REFERENCE's own
    file format (config_der.json: nested sections per derId, ``parent_config`` inheritance -- a child entry such as
    "50_type1" (config_der.json:23-26) names its parent and overrides single keys of single sections), which is what
    the reference hands to ``DERModel(configFile=...)`` (PVDER_env.py:374-378)."""
    with open(config_file or DER_CONFIG_FILE) as fh:
        table = json.load(fh)
    der_id = str(der_id)
    if der_id not in table:
        raise ValueError(f"DER id {der_id!r} is not in {config_file or DER_CONFIG_FILE}")
    entry = table[der_id]
    if "phases" in entry:                       # flat layout
        return dict(entry)
    # reference layout: resolve the parent chain (root first), merge section by section
    chain, seen = [], set()
    cur = der_id
    while cur:
        if cur in seen:
            raise ValueError(f"parent_config cycle at DER id {cur!r}")
        if cur not in table:
            raise ValueError(f"parent_config {cur!r} of DER id {der_id!r} is not in the file")
        seen.add(cur)
        chain.append(table[cur])
        cur = table[cur].get("parent_config", "")
    merged = {}
    for ent in reversed(chain):
        for sec in _REFERENCE_SECTIONS:
            merged.setdefault(sec, {}).update(ent.get(sec, {}))
    flat = {}
    for sec in _REFERENCE_SECTIONS[1:]:
        flat.update(merged[sec])
    model = merged["basic_specs"].get("model_type", "")
    if not model:
        raise ValueError(f"DER id {der_id!r}: basic_specs.model_type is missing")
    flat["phases"] = 1 if "SinglePhase" in model else 3
    flat["wte0"] = flat.pop("wte", 6.28)
    missing = [k for k in ("Np", "Ns", "Vdcmpp0", "Srated", "Ioverload", "Vrmsrated", "Rf_actual", "Lf_actual", "C_actual",
                           "R1_actual", "X1_actual", "Kp_GCC", "Ki_GCC", "Kp_DC", "Ki_DC", "Kp_Q", "Ki_Q", "wp")
               if k not in flat]
    if missing:
        raise ValueError(f"DER id {der_id!r}: missing parameters {missing}")
    return flat


def load_der_parameters(der_id, config_file=None):
    """Per-unit parameters of one DER as a filled ``pvder_params`` struct plus a dict of
    bases/extras (values of reference config_der.json:2-21, :66-84)."""
    raw = read_der_config(der_id, config_file)
    wbase = 2.0 * math.pi * 60.0
    Zbase = VBASE * VBASE / SBASE
    Lbase = Zbase / wbase
    Cbase = 1.0 / (Zbase * wbase)
    a = VGRID_RATED / (raw["Vrmsrated"] * math.sqrt(2.0))
    Zt = complex(raw["R1_actual"], raw["X1_actual"]) / Zbase + (Z2_ACTUAL / Zbase) / (a * a)
    phases = int(raw["phases"])
    Lf = raw["Lf_actual"] / Lbase
    Tact = 298.15
    Irated = (raw["Srated"] / (phases * raw["Vrmsrated"])) * math.sqrt(2.0)
    p = _cabi.Params()
    p.Rf = raw["Rf_actual"] / Zbase
    p.Rt, p.Xt = Zt.real, Zt.imag
    p.inv_Lf = 1.0 / Lf
    p.inv_wb = 1.0 / wbase
    for k in ("Kp_GCC", "Ki_GCC", "Kp_DC", "Ki_DC", "Kp_Q", "Ki_Q", "wp"):
        setattr(p, k, float(raw[k]))
    p.Kp_PLL, p.Ki_PLL = 180.0, 320.0
    p.inv_C = 1.0 / (raw["C_actual"] / Cbase)
    p.w0 = wbase
    p.dw = 0.0                       # w0 - w_grid: the env never changes the grid frequency
    p.vgs = (VGRID_RATED / VBASE) / a
    p.np_iph100 = raw["Np"] * (_ISCR + _KV * (Tact - _T0))
    p.np_irs = raw["Np"] * _IRS
    p.kappa = _QE * VBASE / (_KB * Tact * _A * raw["Ns"])
    p.pv_scale = VBASE / SBASE
    p.Vrms_ref = raw["Vrmsrated"] / VBASE
    p.iref_limit = raw["Ioverload"] * Irated / (SBASE / VBASE)
    p.m_limit10 = 10.0
    p.p_target = REWARD_SPEC["reference_values"]["P_ref"] / SBASE
    p.q_target = REWARD_SPEC["reference_values"]["Q_ref"] / SBASE
    p.Lf = Lf
    p.Rf_Rt = p.Rf + p.Rt
    extras = dict(phases=phases, Vbase=VBASE, Sbase=SBASE, Vdcbase=VBASE, Ibase=SBASE / VBASE, wbase=wbase,
                  Vdc_ref0=raw["Vdcmpp0"] / VBASE, wte0=raw["wte0"], a=a, Zbase=Zbase)
    return p, extras


def validate_n_sim(n):
    """PVDER_env.py:602-620."""
    if n is None:
        return N_SIM_DEFAULT
    if isinstance(n, bool) or not isinstance(n, int):
        raise ValueError("n_sim_time_steps_per_env_step must be an integer!")
    return n if n >= N_SIM_MIN else N_SIM_MIN


def validate_max_sim_time(t, n, max_episode_steps):
    """PVDER_env.py:561-575 (clamped to [1.0, max_episode_steps * n / 60])."""
    limit = max_episode_steps * SIM_TIME_STEP * n
    if t is None:
        return limit
    if isinstance(t, bool) or not isinstance(t, (int, float)):
        raise ValueError("max_sim_time must be a float!")
    if t < MIN_SIM_TIME:
        return MIN_SIM_TIME
    if t > limit:
        return limit
    return t


def validate_discrete(flag):
    """PVDER_env.py:590-600."""
    if flag is None:
        return REWARD_SPEC["DISCRETE_REWARD"]
    if not isinstance(flag, bool):
        raise ValueError("DISCRETE_REWARD must be a boolean!")
    return flag


def validate_goals(goals):
    """PVDER_env.py:577-588.  None and [] (which crash the reference) are rejected."""
    if goals is None or len(goals) == 0 or not set(goals).issubset(GOAL_SPEC.keys()):
        raise ValueError("Goal list:{} contains invalid elements, available elements are:{}".format(
            goals, list(GOAL_SPEC.keys())))
    return list(goals)


def validate_reward_list(goal, reward_list):
    """env_goal_spec[goal]['reward']['my_spec'] (PVDER_env.py:78-93, :249, :445-452): the goal's required terms first,
    then any of its optional ones.  'Vdc_error' is accepted only where the reference can evaluate it (its target is
    defined for power_regulation alone, :242-244; elsewhere the reference raises NameError)."""
    spec = GOAL_SPEC[goal]["reward"]
    if reward_list is None:
        return list(spec["required"])
    terms = list(reward_list)
    allowed = spec["required"] + spec["optional"]
    if (len(terms) > 4 or len(set(terms)) != len(terms) or not set(spec["required"]).issubset(terms)
            or not set(terms).issubset(allowed)):
        raise ValueError("Reward list:{} is invalid for goal {}: required {}, optional {}".format(
            terms, goal, spec["required"], spec["optional"]))
    if "Vdc_error" in terms and goal != "power_regulation":
        raise ValueError("Vdc_error has no target outside the power_regulation goal (reference PVDER_env.py:242-244)")
    return terms


def validate_events_spec(spec):
    out = copy.deepcopy(DEFAULT_EVENTS_SPEC)
    for kind, params in (spec or {}).items():
        if kind not in out:
            raise ValueError("{} is not a valid event!".format(kind))              # PVDER_env.py:433
        for k, v in params.items():
            if k not in out[kind]:
                raise ValueError("{} is not a valid paramter for {} event!".format(k, kind))   # :431
            out[kind][k] = v
    return out


def _grid_index(t, what):
    k = t * SUBSTEPS_PER_SEC
    if abs(k - round(k)) > 1e-9 or round(k) < 0:
        raise ValueError(f"{what}={t} s is not on the half-cycle (1/120 s) grid")
    return int(round(k))


@dataclass
class EnvConfig:
    """Frozen description of one (vector) environment; ``.c`` is the packed C struct."""

    model_type: str = "model_2"                 # reference default, PVDER_env.py:366
    n_sim_time_steps_per_env_step: int = 15
    max_sim_time: float = 40.0
    DISCRETE_REWARD: bool = True
    goals_list: list = field(default_factory=lambda: list(DEFAULT_GOAL))
    # env_goal_spec[goal]['reward']['my_spec'] (PVDER_env.py:249): None = the goal's required term (what update_env_goal
    # (None, None) installs, :445-452); a list = required + chosen optional terms, summed in list order
    reward_list: list | None = None
    events_spec: dict = field(default_factory=lambda: copy.deepcopy(DEFAULT_EVENTS_SPEC))
    event_mode: str = "philox"
    seed: int = 0
    auto_reset: bool = False
    max_episode_steps: int = 500
    micro: int = 1
    # three-phase only.  "auto" (default): each env whose stored state is a balanced set -- the only
    # kind the env itself creates (balanced grid, symmetric reset) -- is integrated on phase a alone,
    # any other state by the general 23-state model.  True / "balanced": always balanced; False /
    # "general": always the general model, one thread per env; "split": the general model with three
    # lanes per env (one per phase) -- the fast path for unbalanced work.
    balanced_three_phase: bool | str = "auto"
    # pvder Grid(unbalance_ratio_b, unbalance_ratio_c): magnitude of the phase-b/c grid voltage relative
    # to phase a.  The reference env always builds Grid(events=...) with 1.0 (PVDER_env.py:372).
    grid_unbalance_ratio: tuple = (1.0, 1.0)
    config_file: str | None = None       # DER parameter file: this project's flat layout or the reference's config_der.json
    der_id: str | None = None            # overrides the derId of env_model_spec (PVDER_env.py:56-58), e.g. "50_type1"
    # Fine steps on the fixed half-cycle grid (level L: 2^L integrator steps of h / 2^L for that sub-step, 0 = off).
    # The reference's LSODA (PVDER_env.py:166) adapts its step where the solution moves fast; the kernel refines exactly
    # those sub-steps: the one whose inputs changed at its start (an event instant, or -- refine_on_action -- an action
    # that moved Q_ref / Vdc_ref) and the PLL pull-in of the first startup_substeps sub-steps after reset (wte0 = 6.28
    # is ~90 degrees from lock, config_der.json:18).  Defaults: DESIGN.md "Fine steps".
    refine_input_level: int = 1
    refine_on_action: bool = False
    startup_substeps: int = 12
    startup_level: int = 3

    def __post_init__(self):
        if self.model_type not in MODEL_SPEC:
            raise ValueError(f"model_type must be one of {list(MODEL_SPEC)}")
        if self.event_mode not in _cabi.EVENT_MODES:
            raise ValueError(f"event_mode must be one of {list(_cabi.EVENT_MODES)}")
        self.n_sim_time_steps_per_env_step = validate_n_sim(self.n_sim_time_steps_per_env_step)
        self.max_sim_time = validate_max_sim_time(self.max_sim_time, self.n_sim_time_steps_per_env_step,
                                                  self.max_episode_steps)
        self.DISCRETE_REWARD = validate_discrete(self.DISCRETE_REWARD)
        self.goals_list = validate_goals(self.goals_list)
        self.events_spec = validate_events_spec(self.events_spec)
        self.par, self.extras = load_der_parameters(self.der_id or MODEL_SPEC[self.model_type]["derId"], self.config_file)
        self.phases = self.extras["phases"]
        if self.phases != (1 if self.model_type == "model_1" else 3):
            raise ValueError(f"DER id {self.der_id!r} has {self.phases} phase(s); {self.model_type} needs the other kind")
        self.n_state = 6 * self.phases + 5
        self.c = self._pack()

    @property
    def sim_time_per_env_step(self):
        return SIM_TIME_STEP * self.n_sim_time_steps_per_env_step          # PVDER_env.py:616

    @property
    def delQref(self):
        return ACTION_SPEC["delQref"] * self.n_sim_time_steps_per_env_step   # PVDER_env.py:617

    @property
    def delVdcref(self):
        return ACTION_SPEC["delVdcref"] * self.n_sim_time_steps_per_env_step  # PVDER_env.py:618

    def steady_state(self, Vgrid=1.0, Sinsol=100.0, Q_ref=0.0):
        """(y0[ns], ma0, ia0) through the library's host-side Newton solve (A.6)."""
        lib = _cabi.load()
        y0 = (C.c_double * _cabi.MAX_STATES)()
        ma0 = (C.c_double * 2)()
        ia0 = (C.c_double * 2)()
        _cabi.check(lib.pvder_steady_state(C.byref(self.par), self.phases, self.extras["Vdc_ref0"], Vgrid, Sinsol,
                                           Q_ref, self.extras["wte0"], y0, ma0, ia0))
        return list(y0)[:self.n_state], complex(ma0[0], ma0[1]), complex(ia0[0], ia0[1])

    def _pack(self):
        c = _cabi.EnvConfigC()
        c.par = self.par
        n = self.n_sim_time_steps_per_env_step
        c.phases = self.phases
        c.n_sub_per_step = 2 * n
        if self.micro not in (1, 2, 4, 8):
            raise ValueError("micro (integrator steps per half-cycle sub-step) must be 1, 2, 4 or 8")
        c.base_level = {1: 0, 2: 1, 4: 2, 8: 3}[self.micro]
        c.done_substep = int(math.ceil(self.max_sim_time * SUBSTEPS_PER_SEC - 1e-9))
        c.discrete_reward = int(self.DISCRETE_REWARD)
        c.goal = _cabi.GOALS[self.goals_list[0]]                          # PVDER_env.py:234
        self.reward_list = validate_reward_list(self.goals_list[0], self.reward_list)
        for t in range(4):
            c.reward_terms[t] = _cabi.REWARD_TERMS[self.reward_list[t]] if t < len(self.reward_list) else -1
        c.auto_reset = int(self.auto_reset)
        v, s = self.events_spec["voltage"], self.events_spec["insolation"]
        enabled = bool(v["ENABLE"] or s["ENABLE"])
        c.event_mode = _cabi.EVENT_MODES[self.event_mode] if (enabled or self.event_mode == "table") else 0
        # instants come from the *voltage* entry for both kinds (PVDER_env.py:408-411)
        c.ev_start_k = _grid_index(v["t_events_start"], "t_events_start")
        c.ev_step_k = max(1, _grid_index(v["t_events_step"], "t_events_step"))
        stop_k = _grid_index(v["t_events_stop"], "t_events_stop")
        c.ev_count = max(0, -(-(stop_k - c.ev_start_k) // c.ev_step_k)) if c.event_mode else 0
        c.ev_voltage_enable, c.ev_insol_enable = int(bool(v["ENABLE"])), int(bool(s["ENABLE"]))
        b3 = self.balanced_three_phase
        if isinstance(b3, bool):
            b3 = "balanced" if b3 else "general"
        if b3 not in _cabi.THREE_PHASE_MODES:
            raise ValueError("balanced_three_phase must be True, False, 'auto', 'balanced', 'general' or 'split'")
        rb, rc = (float(r) for r in self.grid_unbalance_ratio)
        if not (rb > 0.0 and rc > 0.0):
            raise ValueError("grid_unbalance_ratio entries must be positive")
        if (rb, rc) != (1.0, 1.0):
            if self.phases != 3:
                raise ValueError("grid_unbalance_ratio needs the three-phase model (model_2)")
            if b3 == "balanced":
                raise ValueError("an unbalanced grid cannot be integrated by the balanced reduction")
            if b3 == "auto":
                b3 = "split"          # every env is unbalanced: all of them take the general model
        self.three_phase_mode = b3 if self.phases == 3 else "single_phase"
        c.balanced3 = _cabi.THREE_PHASE_MODES[b3] if self.phases == 3 else 0
        c.vg_ratio_b, c.vg_ratio_c = rb, rc
        for name in ("refine_input_level", "startup_level"):
            lv = getattr(self, name)
            if isinstance(lv, bool) or not isinstance(lv, int) or not 0 <= lv <= _cabi.FINE_LEVELS:
                raise ValueError(f"{name} must be an integer in [0, {_cabi.FINE_LEVELS}]")
        if isinstance(self.startup_substeps, bool) or not isinstance(self.startup_substeps, int) or self.startup_substeps < 0:
            raise ValueError("startup_substeps must be a non-negative integer")
        c.refine_input_level, c.refine_on_action = self.refine_input_level, int(bool(self.refine_on_action))
        c.startup_substeps, c.startup_level = self.startup_substeps, self.startup_level
        c.ev_v_min, c.ev_v_max = float(v["min"]), float(v["max"])
        c.ev_s_min, c.ev_s_max = float(s["min"]), float(s["max"])
        c.delQ_pu = self.delQref / self.extras["Sbase"]                   # PVDER_env.py:225
        c.delVdc_pu = self.delVdcref / self.extras["Vdcbase"]             # PVDER_env.py:229
        c.max_sim_time = float(self.max_sim_time)
        c.substeps_per_sec = float(SUBSTEPS_PER_SEC)
        c.seed = int(self.seed) & 0xFFFFFFFFFFFFFFFF
        c.Q_ref0 = 0.0
        c.Vdc_ref0 = self.extras["Vdc_ref0"]
        self.c = c
        y0, self.ma0, self.ia0 = self.steady_state()
        for i, v_ in enumerate(y0):
            c.y0[i] = v_
        self.y0 = y0
        return c

    @property
    def episode_steps(self):
        return -(-self.c.done_substep // self.c.n_sub_per_step)
