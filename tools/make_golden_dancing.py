"""Generate tests/golden/humanoid_dancing.npz from the fp64 oracle (oracle/mjstep_ref.c + oracle/tasks_ref.py).

Same caveat as tools/make_golden.py: MuJoCo is importable nowhere, so these vectors pin the oracle (RK4, PGS, self
contacts of the pinned humanoid), they are not outputs of the reference.  Run:  python tools/make_golden_dancing.py
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from mujoco_gymnasium_environments_b200.tasks import load_tables, TASKS
from oracle import ref
from oracle.tasks_ref import HumanoidDancingRef

t = load_tables("humanoid_dancing")
om = ref.load_model(t)
rng = np.random.default_rng(20261019)
N = 12
env = HumanoidDancingRef(t)
env.reset(sequence=[(k % 10, 1.0 + 0.1 * k) for k in range(20)])
d = env.data
S = dict(qpos=[], qvel=[], ctrl=[], warm=[], qpos1=[], qvel1=[], warm1=[], qpos10=[], ncon=[], pairs=[], dist=[], nefc=[], iters=[])
k = 0
while len(S["qpos"]) < N:
    k += 1
    calm = len(S["qpos"]) < 6          # first half: |ctrl| <= 2 (the 10-step drift bound is asserted on these)
    for _ in range(4 + 3 * (k % 5)):
        d.ctrl[:] = rng.uniform(-1, 1, 29) * 200 * (0.01 if calm else 0.02 + 0.01 * (k % 5))
        ref.mj_step(om, d)
    q = d.qpos.astype(np.float32); v = d.qvel.astype(np.float32); c = d.ctrl.astype(np.float32); w = d.qacc_warmstart.astype(np.float32)
    e = ref.RefData(om)
    e.qpos[:] = q; e.qvel[:] = v; e.ctrl[:] = c; e.qacc_warmstart[:] = w
    ref.mj_forward(om, e)
    con = e.contact
    if len(S["qpos"]) not in (0, 6) and len(con) == 0 and k < 400:
        continue                                   # keep most of the fixture on self-contact states
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
out = {k: np.array(v) for k, v in S.items()}
print("physics fixture: ncon", out["ncon"], "nefc", out["nefc"], "iters", out["iters"])

# task level: injected dance sequences; the post-reset state is stored in fp32 and both sides continue from it (the knees
# rest exactly on their joint limit after reset, so the sign of a 1e-10 residue would otherwise decide a constraint row)
M = 4; STEPS = 40
inject = np.zeros((M, 40), np.float32)
inject[:, 0::2] = rng.integers(0, 10, (M, 20)); inject[:, 1::2] = rng.uniform(1, 3, (M, 20))
inject[0, 1] = 0.25; inject[0, 3] = 0.2; inject[0, 5] = 0.15       # short moves: transitions, history and completion bonus fire
acts = (rng.uniform(-1, 1, (STEPS, M, 29)) * 200 * 0.02).astype(np.float32)
acts[:, 3] *= 10.0                                                    # a vigorous dancer: self contacts, not-upright branch
obs0 = np.zeros((M, 94), np.float32); obs = np.zeros((STEPS, M, 94), np.float32); rew = np.zeros((STEPS, M)); term = np.zeros((STEPS, M), bool)
q0 = np.zeros((M, 29), np.float32); v0 = np.zeros((M, 29), np.float32); w0 = np.zeros((M, 29), np.float32)
ncon = np.zeros((STEPS, M), np.int32)
for k in range(M):
    env = HumanoidDancingRef(t)
    obs0[k], _ = env.reset(sequence=[(int(inject[k, 2 * i]), float(inject[k, 2 * i + 1])) for i in range(20)])
    dd = env.data
    q0[k] = dd.qpos; v0[k] = dd.qvel; w0[k] = dd.qacc_warmstart
    dd.qpos[:] = q0[k]; dd.qvel[:] = v0[k]; dd.qacc_warmstart[:] = w0[k]
    env.prev_joint_vel = dd.qvel[6:].copy()
    for s in range(STEPS):
        obs[s, k], rew[s, k], term[s, k], _, _ = env.step(acts[s, k])
        ncon[s, k] = dd.ncon
out.update(task_inject=inject, task_actions=acts, task_obs0=obs0, task_obs=obs, task_rew=rew, task_term=term,
           task_q0=q0, task_v0=v0, task_w0=w0, task_ncon=ncon)
print("task fixture: ncon max per env", ncon.max(axis=0), "rewards", rew[-1])
path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "humanoid_dancing.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
