"""Gymnasium ids of the reference packages, pointing at the B200 class mirrors.

The reference registers its envs on package import (quadruped_parkour_env/__init__.py:16-34, humanoid_soccer_env/__init__.py:18-26,
humanoid_construction_env/__init__.py:18-26, robotic_arm_assembly_env/__init__.py:12-17; the dancing, rescue and martial-arts
packages register nothing).  ``register_all()`` registers the same ids with the same ``max_episode_steps`` /
``reward_threshold`` when ``gymnasium`` is importable and reports what it did; without gymnasium it returns the table
only, so callers (and the CPU tests) can still inspect it.

Registered vs class values: the reference registers HumanoidSoccer-v0 with ``max_episode_steps=2500`` while the class
truncates at 5000 (soccer_env.py:36), so ``gymnasium.make`` wraps it in a 2500-step TimeLimit -- kept as is.
QuadrupedParkour-v1 differs from -v0 only by ``render_mode='human'``, which the engine rejects (rendering is outside the
hot path): it is listed but not registered.
"""
from __future__ import annotations

from typing import Dict, List

_PKG = "mujoco_gymnasium_environments_b200.envs"
REGISTRY: List[Dict] = [
    dict(id="QuadrupedParkour-v0", entry_point=f"{_PKG}:QuadrupedParkourEnv", max_episode_steps=6000, reward_threshold=8000.0, kwargs={"render_mode": None}),
    dict(id="HumanoidSoccer-v0", entry_point=f"{_PKG}:HumanoidSoccerEnv", max_episode_steps=2500, reward_threshold=8000.0, kwargs={"render_mode": None}),
    dict(id="HumanoidConstruction-v0", entry_point=f"{_PKG}:HumanoidConstructionEnv", max_episode_steps=3000, reward_threshold=10000.0, kwargs={"render_mode": None}),
    dict(id="RoboticArmAssembly-v0", entry_point=f"{_PKG}:RoboticArmAssemblyEnv", max_episode_steps=150000, reward_threshold=8000.0, kwargs={}),
]
# registered only when the caller asks: the reference's register_env() helpers (bipedal_rescue_env/rescue_env.py:803-814,
# humanoid_dancing_env/dancing_env.py:1308-1319); the martial-arts package registers nothing at all
ON_DEMAND: Dict[str, Dict] = {
    "bipedal_rescue": dict(id="BipedalRescue-v0", entry_point=f"{_PKG}:BipedalRescueEnv", max_episode_steps=10000, reward_threshold=20000.0, kwargs={}),
    "humanoid_dancing": dict(id="HumanoidDancing-v0", entry_point=f"{_PKG}:HumanoidDancingEnv", max_episode_steps=3600, reward_threshold=5000.0, kwargs={}),
}
_CLASS_OF = {"quadruped_parkour": "QuadrupedParkourEnv", "humanoid_dancing": "HumanoidDancingEnv", "humanoid_soccer": "HumanoidSoccerEnv",
             "bipedal_rescue": "BipedalRescueEnv", "humanoid_construction": "HumanoidConstructionEnv",
             "humanoid_martial_arts": "HumanoidMartialArtsEnv", "robotic_arm_assembly": "RoboticArmAssemblyEnv"}
NOT_REGISTERED = {"QuadrupedParkour-v1": "render_mode='human' (no renderer in the engine)"}


def register_all(prefix: str = "") -> List[str]:
    """Register the ids (optionally ``prefix``-ed, e.g. ``"B200/"``, to coexist with the reference packages).  Returns the ids
    registered; an empty list when gymnasium is not installed."""
    try:
        from gymnasium.envs.registration import register, registry
    except ImportError:
        return []
    done = []
    for spec in REGISTRY:
        gid = prefix + spec["id"]
        if gid not in registry:
            register(id=gid, entry_point=spec["entry_point"], max_episode_steps=spec["max_episode_steps"],
                     reward_threshold=spec["reward_threshold"], kwargs=dict(spec["kwargs"]))
        done.append(gid)
    return done


def register_env(task: str, prefix: str = "") -> str:
    """The reference's per-package ``register_env()`` (rescue, dancing): registers the id if it is not there yet and returns it;
    returns ``""`` without gymnasium.  Like the reference, a second call is not an error."""
    spec = ON_DEMAND[task]
    try:
        from gymnasium.envs.registration import register, registry
    except ImportError:
        return ""
    gid = prefix + spec["id"]
    if gid not in registry:
        register(id=gid, entry_point=spec["entry_point"], max_episode_steps=spec["max_episode_steps"],
                 reward_threshold=spec["reward_threshold"], kwargs=dict(spec["kwargs"]))
    return gid


def make_env(task: str, render_mode=None, **kwargs):
    """``make_env(render_mode=None, **kwargs)`` of the soccer / construction packages (humanoid_soccer_env/__init__.py:31-42,
    humanoid_construction_env/__init__.py:31-42), for any of the seven tasks: one instance of the class mirror."""
    import importlib
    return getattr(importlib.import_module(_PKG), _CLASS_OF[task])(render_mode=render_mode, **kwargs)


def get_env_info(task: str) -> Dict:
    """``get_env_info()`` of the soccer / construction packages (humanoid_soccer_env/__init__.py:44-56).  The spaces are the ones
    the env really has (the reference's text says Box(25,) / Box(85,) for soccer while the class builds 33 / 80)."""
    from .tasks import TASKS, load_tables
    spec = TASKS[task]; t = load_tables(task)
    reg = next((r for r in REGISTRY + list(ON_DEMAND.values()) if r["entry_point"].endswith(":" + _CLASS_OF[task])), None)
    a = spec.action_space(t); o = spec.observation_space(t)
    return {"name": reg["id"] if reg else None, "version": "1.0.0", "class": _CLASS_OF[task],
            "action_space": f"Box({a.shape[0]},) [{float(a.low.min()):g}, {float(a.high.max()):g}]", "observation_space": f"Box({o.shape[0]},)",
            "max_episode_steps": reg["max_episode_steps"] if reg else spec.max_episode_steps, "class_max_episode_steps": spec.max_episode_steps,
            "reward_threshold": reg["reward_threshold"] if reg else None, "frame_skip": spec.frame_skip, "render_fps": spec.render_fps}
