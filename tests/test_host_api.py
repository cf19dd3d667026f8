"""Host-side pieces of the round-2 boundary that need no GPU: the Gymnasium registration table, the lazily built per-env
infos, the per-task info builders and the batch-options struct."""
import ctypes

import numpy as np
import pytest
import torch

from mujoco_gymnasium_environments_b200 import capi, registration
from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
from mujoco_gymnasium_environments_b200.vector_env import LazyInfos


def test_registration_table_matches_the_reference_ids():
    # quadruped_parkour_env/__init__.py:16-24, humanoid_soccer_env/__init__.py:18-26, humanoid_construction_env/__init__.py:18-26,
    # robotic_arm_assembly_env/__init__.py:12-17
    by_id = {r["id"]: r for r in registration.REGISTRY}
    assert by_id["QuadrupedParkour-v0"]["max_episode_steps"] == 6000 and by_id["QuadrupedParkour-v0"]["reward_threshold"] == 8000.0
    assert by_id["HumanoidSoccer-v0"]["max_episode_steps"] == 2500          # registered value; the class truncates at 5000
    assert by_id["HumanoidConstruction-v0"]["max_episode_steps"] == 3000 and by_id["RoboticArmAssembly-v0"]["max_episode_steps"] == 150000
    for r in registration.REGISTRY:
        mod, cls = r["entry_point"].split(":")
        assert mod == "mujoco_gymnasium_environments_b200.envs"
        import importlib
        assert hasattr(importlib.import_module(mod), cls)
    assert "QuadrupedParkour-v1" in registration.NOT_REGISTERED
    try:
        import gymnasium  # noqa: F401
        assert registration.register_all(prefix="B200Test/") == ["B200Test/" + r["id"] for r in registration.REGISTRY]
    except ImportError:
        assert registration.register_all() == []


def test_package_helpers_follow_the_reference_packages():
    # humanoid_soccer_env/__init__.py:31-56 (make_env, get_env_info), bipedal_rescue_env/rescue_env.py:803-814 (register_env)
    info = registration.get_env_info("humanoid_soccer")
    assert info["name"] == "HumanoidSoccer-v0" and info["max_episode_steps"] == 2500 and info["class_max_episode_steps"] == 5000
    assert info["action_space"].startswith("Box(33,)") and info["observation_space"] == "Box(80,)" and info["reward_threshold"] == 8000.0
    assert registration.get_env_info("bipedal_rescue")["name"] == "BipedalRescue-v0"
    assert registration.get_env_info("humanoid_martial_arts")["name"] is None            # the package registers nothing
    assert set(registration.ON_DEMAND) == {"bipedal_rescue", "humanoid_dancing"}
    assert registration.ON_DEMAND["humanoid_dancing"]["max_episode_steps"] == TASKS["humanoid_dancing"].max_episode_steps == 3600
    try:
        import gymnasium  # noqa: F401
        assert registration.register_env("bipedal_rescue", prefix="B200Test/") == "B200Test/BipedalRescue-v0"
        assert registration.register_env("bipedal_rescue", prefix="B200Test/") == "B200Test/BipedalRescue-v0"     # twice is fine
    except ImportError:
        assert registration.register_env("bipedal_rescue") == ""
    with pytest.raises(NotImplementedError):
        registration.make_env("humanoid_soccer", render_mode="human")                     # no renderer in the engine
    if not torch.cuda.is_available():
        with pytest.raises(capi.B2Error):
            registration.make_env("quadruped_parkour")                                    # no CPU fallback


@pytest.mark.parametrize("task", list(TASKS))
def test_vector_info_has_the_reference_keys_and_env_shaped_values(task):
    spec = TASKS[task]; t = load_tables(task); n = 5
    ti = torch.arange(n * 64, dtype=torch.int32).reshape(n, 64) % 7; tf = torch.rand((n, 64)); x = torch.rand((n, t.nbody, 3))
    info = spec.vector_info(torch, ti, tf, x, t)
    numeric = {"current_move", "task", "task_phase", "held_component", "assembly_progress"}     # strings in the reference, indices here
    skipped = {"combo_chain", "component_status", "ball_contact", "robot_upright"}               # not derivable from the task state rows alone
    for k in spec.info_keys:
        if k in skipped:
            continue
        assert k in info, (task, k)
        v = info[k]
        leaves = list(v.values()) if isinstance(v, dict) else [v]
        for leaf in leaves:
            assert leaf.shape[0] == n, (task, k)


def test_lazy_infos_builds_once_and_only_on_demand():
    calls = []

    class FakeSpec:
        info_keys = ["step_count"]

        @staticmethod
        def vector_info(torch_, ti, tf, xpos, tables):
            calls.append(1)
            return {"step_count": ti[:, 0]}

    class FakeEnv:
        spec = FakeSpec(); tables = None; torch = torch

        def task_state(self, with_xpos=False):
            return torch.ones((3, 4), dtype=torch.int32), torch.zeros((3, 4)), torch.zeros((3, 2, 3))

    infos = LazyInfos(FakeEnv(), {"final_obs": torch.zeros((3, 2)), "_final_obs": torch.zeros(3, dtype=torch.bool)})
    assert infos["final_obs"].shape == (3, 2) and not calls          # eager entries cost nothing
    assert torch.equal(infos["step_count"], torch.ones(3, dtype=torch.int32)) and len(calls) == 1
    assert "task_ti" in infos and "nonexistent" not in infos and len(calls) == 1
    with pytest.raises(KeyError):
        infos["nonexistent"]
    assert set(infos.keys()) >= {"final_obs", "_final_obs", "step_count", "task_ti", "task_tf"}


def _fake_vector_env(n=3):
    """A B200VectorEnv without its device batch: enough for the host-side AsyncVectorEnv surface."""
    from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
    env = B200VectorEnv.__new__(B200VectorEnv)
    env.torch = torch; env.spec = TASKS["bipedal_rescue"]; env.tables = load_tables("bipedal_rescue"); env.num_envs = n
    env.single_action_space = env.spec.action_space(env.tables); env.single_observation_space = env.spec.observation_space(env.tables)
    env._pending = None
    ti = torch.zeros((n, 64), dtype=torch.int32); tf = torch.zeros((n, 64)); ti[:, 11] = torch.arange(n); ti[1, 2] = 0b101
    env.task_state = lambda with_xpos=False: (ti, tf, torch.zeros((n, env.tables.nbody, 3))) if with_xpos else (ti, tf)
    written = []

    class FakeBatch:
        get_task_state = staticmethod(lambda: (ti, tf))
        xpos = staticmethod(lambda: torch.zeros((n, env.tables.nbody, 3)))
        set_task_state = staticmethod(lambda a, b: written.append((a.clone(), b.clone())))
    env.batch = FakeBatch(); env._written = written
    steps = []
    env.step = lambda a: steps.append(a) or ("obs", "rew", "term", "trunc", {})
    env.reset = lambda seed=None, options=None: ("obs0", {"seed": seed})
    return env, steps


def test_async_vector_env_surface_get_attr_call_and_async_pairs():
    # gymnasium.vector.AsyncVectorEnv's calling convention, as BASELINE.md section 3 harness B drives the reference's envs
    env, steps = _fake_vector_env()
    assert env.get_attr("max_episode_steps") == (10000, 10000, 10000) and env.get_attr("render_mode") == (None,) * 3
    assert env.get_attr("victims_remaining") == (5, 4, 3) and env.get_attr("victims_carried") == (0, 2, 0)
    stats = env.get_attr("episode_stats")
    assert len(stats) == 3 and [s["victims_rescued"] for s in stats] == [0, 1, 2]
    assert env.call("render") == (None,) * 3 and env.call("frame_skip") == (1, 1, 1)
    with pytest.raises(AttributeError):
        env.get_attr("no_such_attribute")
    with pytest.raises(AttributeError):
        env.set_attr("victims_rescued", [[], [], []])        # a Python list in the reference, a bit mask here: not an info column
    with pytest.raises(AttributeError):
        env.set_attr("victims_remaining", [1, 2, 3])         # derived (5 - rescued), not stored
    env.set_attr("energy_remaining", [10.0, 20.0, 30.0])
    assert env._written[-1][1][:, 1].tolist() == [10.0, 20.0, 30.0] and env.get_attr("energy_remaining") == (10.0, 20.0, 30.0)
    env.set_attr("energy_remaining", 500.0)
    assert env.get_attr("energy_remaining") == (500.0,) * 3
    env.set_attr("episode_stats", {"falls": 2})
    assert [s["falls"] for s in env.get_attr("episode_stats")] == [2, 2, 2] and len(env._written) == 3
    with pytest.raises(ValueError):
        env.set_attr("energy_remaining", [1.0, 2.0])
    env.step_async("a0")
    with pytest.raises(RuntimeError):
        env.step_async("a1")
    with pytest.raises(RuntimeError):
        env.reset_wait()
    assert env.step_wait()[0] == "obs" and steps == ["a0"]
    with pytest.raises(RuntimeError):
        env.step_wait()
    env.reset_async(seed=7)
    assert env.reset_wait() == ("obs0", {"seed": 7}) and env.unwrapped is env


def test_batch_opts_struct_is_eight_ints_with_the_documented_fields():
    names = [f[0] for f in capi.B2BatchOpts._fields_]
    assert names == ["envs_per_block", "arena_floats", "con_cap", "row_cap", "warps_per_env", "disable_wide", "warmstart_once_per_step", "fifo_queue"]
    assert ctypes.sizeof(capi.B2BatchOpts) == 32
    hdr = open(__file__.replace("tests/test_host_api.py", "include/b2env.h")).read()
    for n in names:
        assert n in hdr, n
