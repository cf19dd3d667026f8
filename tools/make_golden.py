"""Generate tests/golden/quadruped_parkour.npz from the fp64 oracle (oracle/mjstep_ref.c + oracle/tasks_ref.py).

MuJoCo cannot be imported in the authoring container or on the GPU box, so these vectors are NOT outputs of the
reference itself: they freeze the oracle's behaviour (regression pin) and give the GPU tests a fixture that does not
depend on rebuilding the oracle.  Run:  python tools/make_golden.py
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from mujoco_gymnasium_environments_b200.tasks import load_tables, TASKS
from oracle import ref
from oracle.tasks_ref import QuadrupedParkourRef

t = load_tables("quadruped_parkour")
om = ref.load_model(t)
rng = np.random.default_rng(20261018)
N = 8
d = ref.RefData(om)
d.qpos[0:3] = [2, 0, 0.6]
S = dict(qpos=[], qvel=[], ctrl=[], warm=[], qpos1=[], qvel1=[], warm1=[], qpos10=[], qvel10=[], ncon=[], pairs=[], dist=[])
for k in range(N):
    d.ctrl[:16] = rng.uniform(-1, 1, 16) * (0.5 if k < 4 else 10.0)
    ref.mj_step(om, d, 11 + 17 * k)
    # states are stored in fp32 (what the device holds) and the oracle is advanced from exactly those values
    q = d.qpos.astype(np.float32); v = d.qvel.astype(np.float32); c = d.ctrl.astype(np.float32); w = d.qacc_warmstart.astype(np.float32)
    e = ref.RefData(om)
    e.qpos[:] = q; e.qvel[:] = v; e.ctrl[:] = c; e.qacc_warmstart[:] = w
    ref.mj_forward(om, e)
    con = e.contact
    pairs = np.full((32, 2), -1, np.int32); dist = np.zeros(32)
    for i, cc in enumerate(con):
        pairs[i] = (cc.geom1, cc.geom2); dist[i] = cc.dist
    e.qacc_warmstart[:] = w                    # mj_forward left qacc there (MuJoCo 3.x); the fixture steps from the stored warm start
    ref.mj_step(om, e)
    S["qpos"].append(q); S["qvel"].append(v); S["ctrl"].append(c); S["warm"].append(w)
    S["qpos1"].append(e.qpos.copy()); S["qvel1"].append(e.qvel.copy()); S["warm1"].append(e.qacc_warmstart.copy())
    S["ncon"].append(len(con)); S["pairs"].append(pairs); S["dist"].append(dist)
    ref.mj_step(om, e, 9)
    S["qpos10"].append(e.qpos.copy()); S["qvel10"].append(e.qvel.copy())
out = {k: np.array(v) for k, v in S.items()}

# task level: injected resets and 6 control steps with small actions
M = 4
inject = np.array([[0.3, -0.2], [1.0, 0.5], [-1.2, 0.9], [0.0, 0.0]], np.float32)
hi = TASKS["quadruped_parkour"].action_space(t).high
acts = (rng.uniform(-1, 1, (6, M, 16)) * 0.02 * hi).astype(np.float32)
obs0 = np.zeros((M, 95), np.float32); obs = np.zeros((6, M, 95), np.float32); rew = np.zeros((6, M)); term = np.zeros((6, M), bool)
for k in range(M):
    env = QuadrupedParkourRef(t)
    obs0[k], _ = env.reset(randomize=(float(inject[k, 0]), float(inject[k, 1])))
    for s in range(6):
        obs[s, k], rew[s, k], term[s, k], _, _ = env.step(acts[s, k])
out.update(task_inject=inject, task_actions=acts, task_obs0=obs0, task_obs=obs, task_rew=rew, task_term=term)
path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quadruped_parkour.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
