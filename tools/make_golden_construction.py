"""Generate tests/golden/humanoid_construction.npz from the fp64 oracle (oracle/mjstep_ref.c + oracle/tasks_ref.py).

Same caveat as tools/make_golden.py: these vectors pin the oracle (RK4 at 2 ms, Newton, 99 dofs in 12 kinematic trees,
1 202 candidate pairs), they are not outputs of MuJoCo.  Run:  python tools/make_golden_construction.py
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from mujoco_gymnasium_environments_b200.tasks import load_tables
from oracle import ref
from oracle.tasks_ref import HumanoidConstructionRef

t = load_tables("humanoid_construction")
om = ref.load_model(t)
rng = np.random.default_rng(20261022)
N = 6
S = dict(qpos=[], qvel=[], ctrl=[], warm=[], qpos1=[], qvel1=[], qacc=[], qpos5=[], ncon=[], pairs=[], dist=[], nefc=[])
env = HumanoidConstructionRef(t); env.reset(draws=(3, 1.0, 0.1, 20.0)); d = env.data
k = 0
while len(S["qpos"]) < N:
    k += 1
    for _ in range(20 + 15 * (k % 3)):
        d.ctrl[:] = rng.uniform(-1, 1, 33) * 200 * 0.03
        ref.mj_step(om, d)
    q = d.qpos.astype(np.float32); v = d.qvel.astype(np.float32); c = d.ctrl.astype(np.float32); w = d.qacc_warmstart.astype(np.float32)
    e = ref.RefData(om)
    e.qpos[:] = q; e.qvel[:] = v; e.ctrl[:] = c; e.qacc_warmstart[:] = w
    ref.mj_forward(om, e)
    con = e.contact
    if len(con) > 60:
        continue
    pairs = np.full((64, 2), -1, np.int32); dist = np.zeros(64)
    for i, cc in enumerate(con):
        pairs[i] = (cc.geom1, cc.geom2); dist[i] = cc.dist
    S["nefc"].append(e.nefc); S["qacc"].append(e.qacc.copy())
    e.qacc_warmstart[:] = w                    # mj_forward left qacc there (MuJoCo 3.x); the fixture steps from the stored warm start
    ref.mj_step(om, e)
    S["qpos"].append(q); S["qvel"].append(v); S["ctrl"].append(c); S["warm"].append(w)
    S["qpos1"].append(e.qpos.copy()); S["qvel1"].append(e.qvel.copy())
    S["ncon"].append(len(con)); S["pairs"].append(pairs); S["dist"].append(dist)
    ref.mj_step(om, e, 4)
    S["qpos5"].append(e.qpos.copy())
out = {k: np.array(v) for k, v in S.items()}
print("physics fixture: ncon", out["ncon"], "nefc", out["nefc"])

M = 3; STEPS = 10
inject = np.array([[1, 2.0, 0.1, 20.0], [2, 4.0, 0.3, 30.0], [0, 0.5, 0.0, 16.0]], np.float32)
acts = (rng.uniform(-1, 1, (STEPS, M, 33)) * 200 * 0.03).astype(np.float32)
obs0 = np.zeros((M, 135), np.float32); obs = np.zeros((STEPS, M, 135), np.float32); rew = np.zeros((STEPS, M)); term = np.zeros((STEPS, M), bool)
for k in range(M):
    env = HumanoidConstructionRef(t)
    obs0[k], _ = env.reset(draws=tuple(inject[k]))
    for s in range(STEPS):
        obs[s, k], rew[s, k], term[s, k], _, _ = env.step(acts[s, k])
out.update(task_inject=inject, task_actions=acts, task_obs0=obs0, task_obs=obs, task_rew=rew, task_term=term)
print("task fixture rewards", rew[0], rew[-1])
path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "humanoid_construction.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
