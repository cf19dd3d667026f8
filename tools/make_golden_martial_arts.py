"""Generate tests/golden/humanoid_martial_arts.npz from the fp64 oracle (oracle/mjstep_ref.c + oracle/tasks_ref.py).

Same caveat as tools/make_golden.py: these vectors pin the oracle (Euler at 16.67 ms with implicit joint damping, Newton-50,
47 dofs in 4 kinematic trees, 294 candidate pairs incl. the dynamic cylinders of the two dummies), they are not outputs
of MuJoCo.  Run:  python tools/make_golden_martial_arts.py
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from mujoco_gymnasium_environments_b200.tasks import load_tables
from oracle import ref
from oracle.tasks_ref import HumanoidMartialArtsRef

t = load_tables("humanoid_martial_arts")
om = ref.load_model(t)
rng = np.random.default_rng(20261023)
N = 6
S = dict(qpos=[], qvel=[], ctrl=[], warm=[], qpos1=[], qvel1=[], qacc=[], qpos5=[], ncon=[], pairs=[], dist=[], nefc=[])
env = HumanoidMartialArtsRef(t); env.reset(draws=(0.3, -0.2)); d = env.data
k = 0
while len(S["qpos"]) < N:
    k += 1
    for _ in range(6 + 5 * (k % 3)):
        d.ctrl[:] = rng.uniform(-1, 1, 28) * env.ctrl_hi * 0.2
        ref.mj_step(om, d)
    q = d.qpos.astype(np.float32); v = d.qvel.astype(np.float32); c = d.ctrl.astype(np.float32); w = d.qacc_warmstart.astype(np.float32)
    e = ref.RefData(om)
    e.qpos[:] = q; e.qvel[:] = v; e.ctrl[:] = c; e.qacc_warmstart[:] = w
    ref.mj_forward(om, e)
    con = e.contact
    if len(con) > 44 or len(con) < 3:
        continue
    pairs = np.full((48, 2), -1, np.int32); dist = np.zeros(48)
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
gt = np.asarray(t.geom_type)
print("physics fixture: ncon", out["ncon"], "nefc", out["nefc"])
print("pair types seen:", sorted({(int(gt[a]), int(gt[b])) for p in out["pairs"] for a, b in p if a >= 0}))

M = 3; STEPS = 12
inject = np.array([[0.3, -0.2], [-0.45, 0.1], [0.0, 0.49]], np.float32)
acts = (rng.uniform(-1, 1, (STEPS, M, 28)) * 0.3).astype(np.float32)
obs0 = np.zeros((M, 113), np.float32); obs = np.zeros((STEPS, M, 113), np.float32); rew = np.zeros((STEPS, M)); term = np.zeros((STEPS, M), bool)
for k in range(M):
    env = HumanoidMartialArtsRef(t)
    obs0[k], _ = env.reset(draws=tuple(inject[k]))
    for s in range(STEPS):
        obs[s, k], rew[s, k], term[s, k], _, _ = env.step(acts[s, k])
out.update(task_inject=inject, task_actions=acts, task_obs0=obs0, task_obs=obs, task_rew=rew, task_term=term)
print("task fixture rewards", rew[0], rew[-1], "term", term.any(axis=0))
path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "humanoid_martial_arts.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
