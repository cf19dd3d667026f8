"""Generate tests/golden/humanoid_soccer.npz from the fp64 oracle (oracle/mjstep_ref.c + oracle/tasks_ref.py).

Same caveat as tools/make_golden.py: these vectors pin the oracle (Euler + implicit damping, PGS, box-box / capsule-box
ground contacts, joint springs, qfrc_applied / xfrc_applied), they are not outputs of MuJoCo.
Run:  python tools/make_golden_soccer.py
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from mujoco_gymnasium_environments_b200.tasks import load_tables
from oracle import ref
from oracle.tasks_ref import HumanoidSoccerRef

t = load_tables("humanoid_soccer")
om = ref.load_model(t)
rng = np.random.default_rng(20261020)


def draws():
    d = np.zeros(36, np.float32)
    d[0] = rng.uniform(-15, -5); d[1] = rng.uniform(-10, 10); d[2] = rng.uniform(-.5, .5); d[3:32] = rng.uniform(-.1, .1, 29)
    d[32] = rng.uniform(-2, 2); d[33] = rng.uniform(0, 2); d[34] = rng.uniform(0, 2 * np.pi); d[35] = rng.uniform(.05, .15)
    return d


N = 10
env = HumanoidSoccerRef(t)
env.reset(draws=draws())
d = env.data
S = dict(qpos=[], qvel=[], ctrl=[], warm=[], qapp=[], qpos1=[], qvel1=[], warm1=[], qpos10=[], ncon=[], pairs=[], dist=[], nefc=[], iters=[])
k = 0
while len(S["qpos"]) < N:
    k += 1
    calm = len(S["qpos"]) < 5
    for _ in range(6 + 5 * (k % 4)):
        d.ctrl[:] = rng.uniform(-1, 1, 33) * 150 * (0.02 if calm else 0.1)
        ref.mj_step(om, d)
    q = d.qpos.astype(np.float32); v = d.qvel.astype(np.float32); c = d.ctrl.astype(np.float32); w = d.qacc_warmstart.astype(np.float32)
    e = ref.RefData(om)
    e.qpos[:] = q; e.qvel[:] = v; e.ctrl[:] = c; e.qacc_warmstart[:] = w
    ref.mj_forward(om, e)
    con = e.contact
    if e.nefc > 88 or len(con) > 30:
        continue                                   # stay inside the engine's fixed capacities
    pairs = np.full((32, 2), -1, np.int32); dist = np.zeros(32)
    for i, cc in enumerate(con):
        pairs[i] = (cc.geom1, cc.geom2); dist[i] = cc.dist
    S["nefc"].append(e.nefc); S["iters"].append(e.solver_iter)
    e.qacc_warmstart[:] = w                    # mj_forward left qacc there (MuJoCo 3.x); the fixture steps from the stored warm start
    ref.mj_step(om, e)
    S["qpos"].append(q); S["qvel"].append(v); S["ctrl"].append(c); S["warm"].append(w)
    S["qpos1"].append(e.qpos.copy()); S["qvel1"].append(e.qvel.copy()); S["warm1"].append(e.qacc_warmstart.copy())
    S["ncon"].append(len(con)); S["pairs"].append(pairs); S["dist"].append(dist)
    ref.mj_step(om, e, 9)
    S["qpos10"].append(e.qpos.copy())
S.pop("qapp")
out = {k: np.array(v) for k, v in S.items()}
print("physics fixture: ncon", out["ncon"], "nefc", out["nefc"], "iters", out["iters"])

M = 4; STEPS = 30
inject = np.stack([draws() for _ in range(M)])
inject[1, 0] = -11.5; inject[1, 33] = 1.5       # ball starts at x = -9.5 ... and one env with the ball behind x = -10 (goalkeeper reacts)
inject[2, 0] = -14.0
acts = (rng.uniform(-1, 1, (STEPS, M, 33)) * 150 * 0.05).astype(np.float32)
obs0 = np.zeros((M, 80), np.float32); obs = np.zeros((STEPS, M, 80), np.float32); rew = np.zeros((STEPS, M)); term = np.zeros((STEPS, M), bool)
ncon = np.zeros((STEPS, M), np.int32); qapp = np.zeros((STEPS, M))
for k in range(M):
    env = HumanoidSoccerRef(t)
    obs0[k], _ = env.reset(draws=[float(x) for x in inject[k]])
    for s in range(STEPS):
        obs[s, k], rew[s, k], term[s, k], _, _ = env.step(acts[s, k])
        ncon[s, k] = env.data.ncon; qapp[s, k] = env.data.qfrc_applied[0]
out.update(task_inject=inject, task_actions=acts, task_obs0=obs0, task_obs=obs, task_rew=rew, task_term=term, task_ncon=ncon, task_qapp=qapp)
print("task fixture: ncon max per env", ncon.max(axis=0), "goalkeeper force", qapp[-1], "rewards", rew[-1])
path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "humanoid_soccer.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
