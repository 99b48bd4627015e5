"""CPU tier: host logic of the drop-in package and the C-ABI library surface (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gym_pvder_b200 as G
from gym_pvder_b200 import _cabi, config as cfgmod
from gym_pvder_b200.sharding import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "pvder_b200.h")).read()
    declared = set(re.findall(r"\b(pvder_[a-z0-9_]+)\s*\(", header)) - {"pvder_env"}
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pvder_abi_version() == _cabi.ABI_VERSION
    assert lib.pvder_sd_fields(1) == 17 and lib.pvder_sd_fields(3) == 29 and lib.pvder_si_fields() == _cabi.SI_FIELDS
    assert lib.pvder_error_string(-1) == b"invalid argument"


def test_struct_layout_matches_header():
    """ctypes mirror and C struct agree (size is checked through a host-only call using the struct)."""
    cfg = G.EnvConfig(model_type="model_1")
    assert C.sizeof(_cabi.Params) == 29 * 8
    assert C.sizeof(_cabi.EnvConfigC) == 29 * 8 + 22 * 4 + 11 * 8 + 23 * 8 + 2 * 8
    assert _cabi.load().pvder_config_size() == C.sizeof(_cabi.EnvConfigC)
    assert list(cfg.c.y0)[:11] == cfg.y0 and cfg.c.y0[10] == 6.28


def test_missing_extension_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libpvder_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _cabi.load()


def test_vec_env_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.PVDERVecEnv(4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gym-solarpvder-environment_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "host_emul" not in txt or f in ("pvder_env_step.cuh", "pvder_common.cuh", "pvder_split3.cuh"), f


def test_make_and_registration():
    """reference gym_PVDER/__init__.py:3-10, tests:11-14."""
    env = G.make("PVDER-v0")
    assert env.spec.id == "PVDER-v0" and env.spec.max_episode_steps == 500
    assert isinstance(env.unwrapped, G.PVDER)
    assert env.n_sim_time_steps_per_env_step == 15 and env.max_sim_time == 40.0
    assert env.DISCRETE_REWARD is True and env.goals_list == ["voltage_regulation"]
    assert env.action_space == G.Discrete(5) and env.action_space.n == 5
    assert env.observation_space.shape == (11,) and env.observation_space.dtype == np.float32
    assert env.metadata == {"render.modes": ["vector", "human"]}
    assert len(env.observed_quantities) == 11
    with pytest.raises(KeyError):
        G.make("PVDER-v1")
    with pytest.raises(AssertionError):
        env.step(0)                      # TimeLimit: step before reset


def test_kwargs_validation():
    """PVDER_env.py:561-620."""
    assert cfgmod.validate_n_sim(None) == 60 and cfgmod.validate_n_sim(0) == 1 and cfgmod.validate_n_sim(7) == 7
    with pytest.raises(ValueError):
        cfgmod.validate_n_sim(1.5)
    assert cfgmod.validate_max_sim_time(None, 15, 500) == pytest.approx(125.0)
    assert cfgmod.validate_max_sim_time(0.5, 15, 500) == 1.0
    assert cfgmod.validate_max_sim_time(1e9, 15, 500) == pytest.approx(125.0)
    with pytest.raises(ValueError):
        cfgmod.validate_max_sim_time("40", 15, 500)
    assert cfgmod.validate_discrete(None) is True
    with pytest.raises(ValueError):
        cfgmod.validate_discrete(1)
    with pytest.raises(ValueError):
        cfgmod.validate_goals(["voltage"])
    with pytest.raises(ValueError):
        cfgmod.validate_goals([])
    with pytest.raises(ValueError):
        G.make("PVDER-v0", DISCRETE_REWARD="yes")
    with pytest.raises(ValueError):
        G.make("PVDER-v0", goals_list=["fly"])
    env = G.make("PVDER-v0", n_sim_time_steps_per_env_step=10, max_sim_time=25.0)
    u = env.unwrapped
    assert u._delQref == 250 and u._delVdcref == pytest.approx(0.2) and u._sim_time_per_env_step == pytest.approx(10 / 60)


def test_env_config_packing():
    c = G.EnvConfig(model_type="model_1", n_sim_time_steps_per_env_step=10, max_sim_time=25.0)
    assert c.c.phases == 1 and c.c.n_sub_per_step == 20 and c.c.done_substep == 3000 and c.episode_steps == 150
    assert (c.c.ev_start_k, c.c.ev_step_k, c.c.ev_count) == (120, 120, 38)          # arange(1, 39, 1)
    assert c.c.ev_voltage_enable == 1 and c.c.ev_insol_enable == 0 and c.c.event_mode == 1
    assert c.c.delQ_pu == 250 / 50e3 and c.c.delVdc_pu == pytest.approx(0.2 / 500)
    assert c.c.goal == 0 and c.c.discrete_reward == 1
    off = G.EnvConfig(events_spec={"voltage": {"ENABLE": False}})
    assert off.c.event_mode == 0 and off.c.ev_count == 0 and off.c.phases == 3 and off.n_state == 23
    with pytest.raises(ValueError):
        G.EnvConfig(events_spec={"voltage": {"t_events_start": 1.001}})
    with pytest.raises(ValueError):
        G.EnvConfig(events_spec={"wind": {"min": 1}})
    with pytest.raises(ValueError):
        G.EnvConfig(model_type="model_9")


def test_steady_state_through_the_c_abi():
    """pvder_steady_state (host-only) against the oracle's independent Newton solve."""
    from oracle.pvder_model import PVDERModel, load_der_params

    for mt, der in (("model_1", "10"), ("model_2", "50")):
        cfg = G.EnvConfig(model_type=mt)
        y0, ma0, ia0 = PVDERModel(load_der_params(der)).steady_state()
        np.testing.assert_allclose(cfg.y0, y0, rtol=0, atol=1e-14)
        assert abs(cfg.ma0 - ma0) < 1e-14 and abs(cfg.ia0 - ia0) < 1e-14


def test_spaces():
    d = G.Discrete(5, seed=0)
    assert all(0 <= d.sample() < 5 for _ in range(50)) and d.contains(4) and not d.contains(5) and not d.contains(1.0)
    assert 3 in d and np.int64(2) in d and True not in d
    b = G.Box(-10, 10, (11,), np.float32)
    assert b.contains(np.zeros(11, np.float32)) and not b.contains(np.full(11, 11.0, np.float32))
    assert not b.contains(np.zeros(10, np.float32)) and b.sample().dtype == np.float32


def test_update_env_events_is_per_instance():
    """the reference mutates a class-level dict; here the spec is per instance."""
    a, b = G.make("PVDER-v0"), G.make("PVDER-v0")
    a.update_env_events([{"voltage": {"min": 0.95}}])
    assert a.env_events_spec["voltage"]["min"] == 0.95 and b.env_events_spec["voltage"]["min"] == 0.98
    with pytest.raises(AssertionError):
        a.update_env_events({"voltage": {"min": 0.9}})
    with pytest.raises(ValueError):
        a.update_env_events([{"voltage": {"foo": 1}}])


def test_shard_bounds():
    for total, world in ((1 << 20, 8), (1000, 3), (5, 8)):
        spans = [shard_bounds(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) prints ONE JSON line with the
    contract's keys; a tiny run: 1 step of 64 env steps per core."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_bench_reference_arm_under_torchrun_prints_one_line():
    """The driver launches the reference arm like the GPU arm (torchrun, one rank per GPU): rank 0 alone runs and prints
    the ONE JSON line, the other ranks exit 0 without work; nothing else reaches stdout."""
    import json
    import socket
    import subprocess
    import sys

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0


def test_host_buffer_chunk_plan_invariants():
    """pvder_plan_chunks (pure host arithmetic behind pvder_env_step_host): the plan covers the batch exactly, starts with a
    one-wave chunk, never has more than 12 chunks or a chunk below one unit, puts the largest chunk second and shrinks
    from there by about the requested copy/kernel time ratio."""
    import ctypes as C
    from gym_pvder_b200 import _cabi
    lib = _cabi.load()
    sizes = (C.c_int64 * 12)()
    assert lib.pvder_plan_chunks(0, 0.5, sizes) == 1 and sizes[0] == 0          # small batches: one launch
    assert lib.pvder_plan_chunks(23, 0.5, sizes) == 1 and sizes[0] == 23
    for units in list(range(24, 400)) + [1000, 4321, 100000]:
        for q in (0.0, 0.3, 0.41, 0.5, 0.66, 0.8, 0.9, 1.7, float("nan")):
            m = lib.pvder_plan_chunks(units, q, sizes)
            s = [sizes[i] for i in range(m)]
            assert 2 <= m <= 12 and sum(s) == units and s[0] == 4 and min(s) >= 1, (units, q, s)
            assert all(s[i] >= s[i + 1] for i in range(1, m - 1)), (units, q, s)
    m = lib.pvder_plan_chunks(110, 0.66, sizes)                                  # 1 Mi single-phase envs on a B200
    s = [sizes[i] for i in range(m)]
    assert m >= 9 and all(abs(s[i + 1] / s[i] - 0.66) < 0.2 for i in range(1, 6)), s


def test_registers_with_a_real_gym_when_one_is_importable(monkeypatch):
    """the id is also registered with gym / gymnasium when importable (neither is in this image: a stand-in
    module records the call).  The entry point receives only its kwargs -- no `spec=` -- like under a real gym.make."""
    import sys
    import types

    from gym_pvder_b200 import registration

    calls = []
    fake = types.ModuleType("gymnasium")
    fake.envs = types.SimpleNamespace(registry={})
    fake.register = lambda **kw: calls.append(kw)
    monkeypatch.setitem(sys.modules, "gymnasium", fake)
    monkeypatch.setitem(sys.modules, "gym", None)          # "import gym" raises ImportError
    seen = {}

    class Probe:
        max_sim_time_user = None

        def __init__(self, **kw):
            seen.update(kw)

    registration.register(id="PVDER-probe-v0", entry_point=Probe, kwargs={"x": 1}, max_episode_steps=7)
    assert calls == [{"id": "PVDER-probe-v0", "entry_point": Probe, "kwargs": {"x": 1}, "max_episode_steps": 7}]
    env = registration.make("PVDER-probe-v0", y=2)
    assert seen == {"x": 1, "y": 2}                        # no spec= kwarg
    assert env.env.spec.id == "PVDER-probe-v0" and env._max_episode_steps == 7
    del registration.registry["PVDER-probe-v0"]


def test_generated_model_headers_are_reproducible(tmp_path):
    """The committed csrc/pvder_model_*.cuh are exactly what tools/gen_model.py emits with its default switches (pivot
    order Vdc, delta before the current pair; plain LU) -- nobody edits generated code by hand, and the study switches
    (PVDER_GEN_*) default to the committed configuration."""
    import subprocess
    import sys

    env = {k: v for k, v in os.environ.items() if not k.startswith("PVDER_GEN_")}
    env["PVDER_GEN_OUT"] = str(tmp_path)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "tools", "gen_model.py")], check=True, env=env, capture_output=True,
                   timeout=900)
    csrc = os.path.join(root, "gym-solarpvder-environment_b200", "csrc")
    for name in ("pvder_model_1ph.cuh", "pvder_model_3ph.cuh", "pvder_model_3ph_bal.cuh"):
        with open(os.path.join(csrc, name)) as fa, open(os.path.join(str(tmp_path), name)) as fb:
            assert fa.read() == fb.read(), f"{name} differs from the generator's output"
