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
