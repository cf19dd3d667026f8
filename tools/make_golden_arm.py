"""Generate tests/golden/robotic_arm_assembly.npz from the fp64 oracle (oracle/mjstep_ref.c + oracle/tasks_ref.py).

Same caveat as tools/make_golden.py: these vectors pin the oracle (10 Euler sub-steps of 2 ms, Newton-50, 63 dofs in 10
kinematic trees, 785 candidate pairs of which 18 are condim 6), they are not outputs of MuJoCo.  Random actions never bring
a gripper pad onto a component, so two of the physics states and the task case are crafted: a component is placed between
the pads.  Run:  python tools/make_golden_arm.py
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from mujoco_gymnasium_environments_b200.tasks import load_tables
from oracle import ref
from oracle.tasks_ref import RoboticArmAssemblyRef

t = load_tables("robotic_arm_assembly")
om = ref.load_model(t)
rng = np.random.default_rng(20261025)
CAP = 128
S = dict(qpos=[], qvel=[], ctrl=[], warm=[], qpos1=[], qvel1=[], qacc=[], qpos5=[], ncon=[], pairs=[], dist=[], dims=[], nefc=[])
env = RoboticArmAssemblyRef(t); d = env.data


def qmul(a, b):
    return np.array([a[0]*b[0]-a[1]*b[1]-a[2]*b[2]-a[3]*b[3], a[0]*b[1]+a[1]*b[0]+a[2]*b[3]-a[3]*b[2],
                     a[0]*b[2]-a[1]*b[3]+a[2]*b[0]+a[3]*b[1], a[0]*b[3]+a[1]*b[2]-a[2]*b[1]+a[3]*b[0]])


def craft(env, comp, opening=0.0085):
    """Put component `comp` between the gripper pads of the current pose, its thin (local z) axis along the closing
    direction, fingers closed to a 3 mm gap: the 4 mm thick CPU chip then touches both pads with 0.5 mm penetration."""
    d = env.data
    d.qpos[7] = opening; d.qpos[8] = opening; d.ctrl[7] = opening; d.ctrl[8] = opening
    ref.mj_forward(env.model, d)
    gl = t.name2id("geom", "gripper_left_pad"); gr = t.name2id("geom", "gripper_right_pad")
    mid = 0.5 * (d.geom_xpos[gl] + d.geom_xpos[gr])
    a = env.comp_qadr[comp]
    qg = d.xquat[t.name2id("body", "gripper_mount")]
    d.qpos[a:a + 3] = mid; d.qpos[a + 3:a + 7] = qmul(qg, np.array([np.sqrt(0.5), 0, np.sqrt(0.5), 0]))
    b = env.comp_body[comp]; da = int(t.jnt_dofadr[int(t.body_jntadr[b])])
    d.qvel[da:da + 6] = 0


def record(d):
    q = d.qpos.astype(np.float32); v = d.qvel.astype(np.float32); c = d.ctrl.astype(np.float32); w = d.qacc_warmstart.astype(np.float32)
    e = ref.RefData(om)
    e.qpos[:] = q; e.qvel[:] = v; e.ctrl[:] = c; e.qacc_warmstart[:] = w
    ref.mj_forward(om, e)
    con = e.contact
    assert len(con) <= CAP, len(con)
    pairs = np.full((CAP, 2), -1, np.int32); dist = np.zeros(CAP); dims = np.zeros(CAP, np.int32)
    for i, cc in enumerate(con):
        pairs[i] = (cc.geom1, cc.geom2); dist[i] = cc.dist; dims[i] = cc.dim
    S["nefc"].append(e.nefc); S["qacc"].append(e.qacc.copy())
    e.qacc_warmstart[:] = w                    # mj_forward left qacc there (MuJoCo 3.x); the fixture steps from the stored warm start
    ref.mj_step(om, e)
    S["qpos"].append(q); S["qvel"].append(v); S["ctrl"].append(c); S["warm"].append(w)
    S["qpos1"].append(e.qpos.copy()); S["qvel1"].append(e.qvel.copy())
    S["ncon"].append(len(con)); S["pairs"].append(pairs); S["dist"].append(dist); S["dims"].append(dims)
    ref.mj_step(om, e, 4)
    S["qpos5"].append(e.qpos.copy())


for k in range(2):
    for _ in range(4 + 3 * k):
        env.step(rng.uniform(env.action_low, env.action_high))
    record(d)
for comp, opening in (("cpu", 0.0085), ("battery", 0.0)):
    env.reset()
    craft(env, comp, opening)
    record(d)
out = {k: np.array(v) for k, v in S.items()}
print("physics fixture: ncon", out["ncon"], "nefc", out["nefc"], "condim-6 contacts per state", (out["dims"] == 6).sum(axis=1))

# task fixtures: (a) reset + random actions, (b) crafted pickup: the CPU between the pads, then the gripper opens
STEPS = 6
acts = rng.uniform(env.action_low, env.action_high, (STEPS, 9)).astype(np.float32)
env = RoboticArmAssemblyRef(t)
obs0, _ = env.reset()
obs = np.zeros((STEPS, 110), np.float32); rew = np.zeros(STEPS); term = np.zeros(STEPS, bool)
for s in range(STEPS):
    obs[s], rew[s], term[s], _, _ = env.step(acts[s])
out.update(task_actions=acts, task_obs0=obs0, task_obs=obs, task_rew=rew, task_term=term)
env = RoboticArmAssemblyRef(t); env.reset(); d = env.data
d.qvel[:] = 0; d.qacc_warmstart[:] = 0          # the arm is mounted through the table top and flails after reset: start the grasp from rest
craft(env, "cpu")
cq = d.qpos.astype(np.float32); cv = d.qvel.astype(np.float32); cw = d.qacc_warmstart.astype(np.float32)
d.qpos[:] = cq; d.qvel[:] = cv; d.qacc_warmstart[:] = cw
cacts = np.zeros((4, 9), np.float32); cacts[:, 7] = [8.5, 8.5, 0.0, 0.0]              # hold the 3 mm gap, then release (larger values close the fingers)
cobs = np.zeros((4, 110), np.float32); crew = np.zeros(4); cterm = np.zeros(4, bool); cheld = np.zeros(4, np.int32); cphase = np.zeros(4, np.int32)
for s in range(4):
    cobs[s], crew[s], cterm[s], _, _ = env.step(cacts[s])
    cheld[s] = env.SEQ.index(env.held_component) if env.held_component else -1; cphase[s] = env.PHASES[env.task_phase]
print("crafted pickup: held", cheld, "phase", cphase, "reward", crew.round(1), "status", env.component_status["cpu"])
out.update(craft_qpos=cq, craft_qvel=cv, craft_warm=cw, craft_actions=cacts, craft_obs=cobs, craft_rew=crew, craft_term=cterm, craft_held=cheld, craft_phase=cphase)
path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "robotic_arm_assembly.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
