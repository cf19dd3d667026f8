"""Generate tests/golden/bipedal_rescue.npz from the fp64 oracle (oracle/mjstep_ref.c + oracle/tasks_ref.py).

Same caveat as tools/make_golden.py: these vectors pin the oracle (RK4, PGS, 3 175 candidate pairs, explicit gripper
pairs, victims as six-joint bodies), they are not outputs of MuJoCo.  Run:  python tools/make_golden_rescue.py
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from mujoco_gymnasium_environments_b200.tasks import load_tables
from oracle import ref
from oracle.tasks_ref import BipedalRescueRef

t = load_tables("bipedal_rescue")
om = ref.load_model(t)
rng = np.random.default_rng(20261021)


def draws():
    d = np.zeros(12, np.float32); d[0:2] = rng.uniform(-5, 5, 2); d[2:] = rng.uniform(-1, 1, 10)
    return d


N = 8
S = dict(qpos=[], qvel=[], ctrl=[], warm=[], qpos1=[], qvel1=[], warm1=[], qpos5=[], ncon=[], pairs=[], dist=[], nefc=[], iters=[])
tries = 0
while len(S["qpos"]) < N:
    tries += 1
    env = BipedalRescueRef(t); env.reset(draws=draws()); d = env.data
    for _ in range(3 + 4 * (tries % 5)):
        d.ctrl[:] = rng.uniform(-1, 1, 26) * 100 * (0.02 if len(S["qpos"]) < 4 else 0.1)
        ref.mj_step(om, d)
    q = d.qpos.astype(np.float32); v = d.qvel.astype(np.float32); c = d.ctrl.astype(np.float32); w = d.qacc_warmstart.astype(np.float32)
    e = ref.RefData(om)
    e.qpos[:] = q; e.qvel[:] = v; e.ctrl[:] = c; e.qacc_warmstart[:] = w
    ref.mj_forward(om, e)
    con = e.contact
    if e.nefc > 116 or len(con) > 44:
        continue                                   # stay inside the engine's fixed capacities
    pairs = np.full((48, 2), -1, np.int32); dist = np.zeros(48)
    for i, cc in enumerate(con):
        pairs[i] = (cc.geom1, cc.geom2); dist[i] = cc.dist
    S["nefc"].append(e.nefc); S["iters"].append(e.solver_iter)
    ok = True
    e.qacc_warmstart[:] = w                    # mj_forward left qacc there (MuJoCo 3.x); the fixture steps from the stored warm start
    ref.mj_step(om, e)
    q1, v1, w1 = e.qpos.copy(), e.qvel.copy(), e.qacc_warmstart.copy()
    S["qpos"].append(q); S["qvel"].append(v); S["ctrl"].append(c); S["warm"].append(w)
    S["qpos1"].append(q1); S["qvel1"].append(v1); S["warm1"].append(w1)
    S["ncon"].append(len(con)); S["pairs"].append(pairs); S["dist"].append(dist)
    ref.mj_step(om, e, 4)
    S["qpos5"].append(e.qpos.copy())
out = {k: np.array(v) for k, v in S.items()}
print("physics fixture: ncon", out["ncon"], "nefc", out["nefc"], "iters", out["iters"])

M = 3; STEPS = 12
inject = np.stack([draws() for _ in range(M)])
acts = (rng.uniform(-1, 1, (STEPS, M, 26)) * 100 * 0.03).astype(np.float32)
obs0 = np.zeros((M, 102), np.float32); obs = np.zeros((STEPS, M, 102), np.float32); rew = np.zeros((STEPS, M)); term = np.zeros((STEPS, M), bool)
q0 = np.zeros((M, 63), np.float32); v0 = np.zeros((M, 63), np.float32); w0 = np.zeros((M, 63), np.float32)
ncon = np.zeros((STEPS, M), np.int32); nefc = np.zeros((STEPS, M), np.int32)
for k in range(M):
    env = BipedalRescueRef(t)
    obs0[k], _ = env.reset(draws=[float(x) for x in inject[k]])
    dd = env.data
    # both sides continue from the fp32 post-reset state (finger slides rest exactly on their limits after reset)
    q0[k] = dd.qpos; v0[k] = dd.qvel; w0[k] = dd.qacc_warmstart
    dd.qpos[:] = q0[k]; dd.qvel[:] = v0[k]; dd.qacc_warmstart[:] = w0[k]
    for s in range(STEPS):
        obs[s, k], rew[s, k], term[s, k], _, _ = env.step(acts[s, k])
        ncon[s, k] = dd.ncon; nefc[s, k] = dd.nefc
out.update(task_inject=inject, task_actions=acts, task_obs0=obs0, task_obs=obs, task_rew=rew, task_term=term,
           task_q0=q0, task_v0=v0, task_w0=w0, task_ncon=ncon, task_nefc=nefc)
print("task fixture: ncon max", ncon.max(axis=0), "nefc max", nefc.max(axis=0), "rewards", rew[0], rew[-1])
path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bipedal_rescue.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
