"""The oracle reproduces its committed golden vectors (tests/golden/, made by tools/make_golden.py), and the
quadruped task restatement keeps the reference's observable semantics."""
import os

import numpy as np
import pytest

from oracle import ref
from oracle.tasks_ref import QuadrupedParkourRef

GOLD = os.path.join(os.path.dirname(__file__), "golden", "quadruped_parkour.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_physics_golden(quad_tables, gold):
    om = ref.load_model(quad_tables)
    for k in range(gold["qpos"].shape[0]):
        d = ref.RefData(om)
        d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
        ref.mj_forward(om, d)
        assert d.ncon == gold["ncon"][k]
        assert [(c.geom1, c.geom2) for c in d.contact] == [tuple(p) for p in gold["pairs"][k][:d.ncon].tolist()]
        d.qacc_warmstart[:] = gold["warm"][k]       # mj_forward left qacc there (MuJoCo 3.x mj_fwdConstraint); step from the stored warm start
        ref.mj_step(om, d)
        assert np.allclose(d.qpos, gold["qpos1"][k], rtol=0, atol=1e-12)
        assert np.allclose(d.qvel, gold["qvel1"][k], rtol=0, atol=1e-10)
        ref.mj_step(om, d, 9)
        assert np.allclose(d.qpos, gold["qpos10"][k], rtol=0, atol=1e-9)


def test_task_golden(quad_tables, gold):
    for k in range(gold["task_inject"].shape[0]):
        env = QuadrupedParkourRef(quad_tables)
        o, _ = env.reset(randomize=tuple(float(x) for x in gold["task_inject"][k]))
        assert np.allclose(o, gold["task_obs0"][k], atol=1e-6)
        for s in range(gold["task_actions"].shape[0]):
            o, r, te, tr, info = env.step(gold["task_actions"][s, k])
            assert np.allclose(o, gold["task_obs"][s, k], atol=1e-6)
            assert r == pytest.approx(gold["task_rew"][s, k], abs=1e-8) and te == gold["task_term"][s, k]


def test_quadruped_observation_layout_and_quirks(quad_tables):
    env = QuadrupedParkourRef(quad_tables, seed=3)
    obs, info = env.reset()
    assert obs.shape == (95,) and obs.dtype == np.float32
    assert np.all(obs[45:49] == 0)                       # F8: foot body ids compared with contact geom ids
    assert np.all(obs[61:85] == 10.0) and obs[93] == 0.0 and obs[94] == np.float32(0.8)
    assert obs[85] == pytest.approx(8.0 - env.data.xpos[1][0], abs=1e-5) and obs[86] == 1.0
    # reset randomises bl_knee / bl_ankle (joint ids 17/18 used as qpos addresses), not the platform/pendulum
    env2 = QuadrupedParkourRef(quad_tables)
    env2.reset(randomize=(1.25, -0.75))
    env3 = QuadrupedParkourRef(quad_tables)
    env3.reset(randomize=(0.0, 0.0))
    assert env2.data.qpos[23] == pytest.approx(env3.data.qpos[23], abs=1e-9)     # platform_slide untouched
    assert abs(env2.data.qpos[17] - env3.data.qpos[17]) > 0.5
    assert set(info) == {"step_count", "episode_reward", "max_forward_progress", "checkpoints_reached", "fall_count", "course_completion"}


def test_quadruped_step_semantics(quad_tables):
    env = QuadrupedParkourRef(quad_tables)
    env.reset(randomize=(0.0, 0.0))
    a = np.full(16, 1e6, np.float32)
    o, r, te, tr, info = env.step(a)
    assert np.allclose(env.data.ctrl[:16], env.action_high)        # clipped to the action space
    assert env.data.ctrl[16] == 0.0                                 # 50 sin(0.5 * 0 * dt): pre-increment counter
    assert info["step_count"] == 1 and not tr
    env.step(np.zeros(16, np.float32))
    assert env.data.ctrl[16] == pytest.approx(50 * np.sin(0.5 * 0.01)) and env.data.ctrl[17] == pytest.approx(100 * np.sin(0.3 * 0.01))
    # ncon > 8 penalty: the two unstable platforms settle on both planes (16 contacts) -> always on
    env = QuadrupedParkourRef(quad_tables)
    env.reset(randomize=(0.0, 0.0))
    for _ in range(40):
        o, r, te, tr, info = env.step(np.zeros(16, np.float32))
    assert env.data.ncon > 8 and r < -400
    env.step_count = 6000
    assert env.step(np.zeros(16, np.float32))[3] is True
