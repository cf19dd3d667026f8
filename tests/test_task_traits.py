"""The step kernels are specialised per task at compile time (csrc/b2_tasks.cuh): solver, condim-6 rows and, since round 2, whether
the convex (MPR) path is compiled in at all.  b2_batch_create refuses a model that contradicts its task's traits; this host test
checks the shipped traits against the shipped model tables, so a table regenerated from a changed MJCF cannot silently lose
its cylinder contacts."""
import os
import re

import numpy as np
import pytest

from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables

HERE = os.path.dirname(os.path.abspath(__file__))
STRUCT = {"quadruped_parkour": "QuadrupedTask", "humanoid_dancing": "DancingTask", "humanoid_soccer": "SoccerTask", "bipedal_rescue": "RescueTask",
          "humanoid_construction": "ConstructionTask", "humanoid_martial_arts": "MartialArtsTask", "robotic_arm_assembly": "ArmTask"}


def traits(struct):
    src = open(os.path.join(HERE, "..", "mujoco_gymnasium_environments_b200", "csrc", "b2_tasks.cuh")).read()
    body = src[src.index("struct " + struct):]
    body = body[:body.index("\n};")] if "\n};" in body else body
    get = lambda name: re.search(r"\b" + name + r" = ([A-Za-z0-9_<> ()]+?)[,;]", body).group(1)
    return dict(solver=int(get("SOLVER")), condim6=get("CONDIM6") == "true", convex=get("CONVEX_PAIRS") == "true")


@pytest.mark.parametrize("task", sorted(TASKS))
def test_compile_time_traits_match_the_model_tables(task):
    t = load_tables(task); tr = traits(STRUCT[task])
    gt = np.asarray(t.geom_type); a = gt[np.asarray(t.pair_g1)]; b = gt[np.asarray(t.pair_g2)]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    nconvex = int((((lo == 3) & (hi == 5)) | ((lo == 5) & ((hi == 5) | (hi == 6)))).sum())      # capsule-cylinder, cylinder-cylinder, cylinder-box
    assert tr["convex"] == (nconvex > 0), (task, nconvex)
    assert tr["condim6"] == bool((np.asarray(t.pair_condim) == 6).any()), task
    assert tr["solver"] == int(t.arrays["solver"]) if hasattr(t, "arrays") else True
