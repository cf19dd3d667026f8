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
S = dict(qpos=[], qvel=[], ctrl=[], warm=[], qpos1=[], qvel1=[], qacc=[], qpos5=[], ncon=[], pairs=[], dist=[], dims=[], nefc=[],
         qacc_sens=[], qpos1_sens=[], qvel1_sens=[], qpos5_sens=[])
from oracle.twin import perturbed, spread
prng = np.random.default_rng(5)
NTWIN = 8
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
    # the oracle's own response to fp32-sized perturbations of this state (oracle/twin.py): flat cylinder caps resting on flat
    # faces make MPR's contact point a function of the last bits
    tw = dict(qacc=[], qpos1=[], qvel1=[], qpos5=[])
    for _ in range(NTWIN):
        g = ref.RefData(om)
        g.qpos[:] = perturbed(q, prng); g.qvel[:] = perturbed(v, prng); g.ctrl[:] = c; g.qacc_warmstart[:] = w
        ref.mj_forward(om, g); tw["qacc"].append(g.qacc.copy())
        g.qacc_warmstart[:] = w
        ref.mj_step(om, g); tw["qpos1"].append(g.qpos.copy()); tw["qvel1"].append(g.qvel.copy())
        ref.mj_step(om, g, 4); tw["qpos5"].append(g.qpos.copy())
    for k_ in tw:
        S[k_ + "_sens"].append(spread(tw[k_], S[k_][-1]))


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
# twins of the task rollout: same actions from a perturbed post-reset state
tobs = np.zeros((NTWIN, STEPS, 110)); trew = np.zeros((NTWIN, STEPS))
for k_ in range(NTWIN):
    g = RoboticArmAssemblyRef(t); g.reset()
    g.data.qpos[:] = perturbed(g.data.qpos, prng); g.data.qvel[:] = perturbed(g.data.qvel, prng)
    for s in range(STEPS):
        tobs[k_, s], trew[k_, s], _, _, _ = g.step(acts[s])
out.update(task_obs_sens=spread(tobs, obs), task_rew_sens=spread(trew, rew))
# twins of reset(): the ten settle steps from a perturbed copy of the state reset() writes (ctrl 0, as in reset())
t0 = np.zeros((NTWIN, 110))
for k_ in range(NTWIN):
    g = RoboticArmAssemblyRef(t); g.reset()
    ref.mj_resetData(g.model, g.data)
    q0 = np.array(g.data.qpos); q0[0:7] = [0, -0.5, 0.5, 0, 0.5, 0, 0]
    for c_ in g.SEQ:
        a_ = g.comp_qadr[c_]; q0[a_:a_ + 3] = g.INITIAL[c_]; q0[a_ + 3:a_ + 7] = [1, 0, 0, 0]
    g.data.qpos[:] = perturbed(q0, prng); g.data.qvel[:] = perturbed(np.zeros_like(np.array(g.data.qvel)), prng)
    ref.mj_step(g.model, g.data, 10)
    t0[k_] = g._get_observation()
out.update(task_obs0_sens=spread(t0, obs0))
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
tobs = np.zeros((NTWIN, 4, 110)); trew = np.zeros((NTWIN, 4))
for k_ in range(NTWIN):
    g = RoboticArmAssemblyRef(t); g.reset()
    g.data.qpos[:] = perturbed(cq, prng); g.data.qvel[:] = perturbed(cv, prng); g.data.qacc_warmstart[:] = cw
    for s in range(4):
        tobs[k_, s], trew[k_, s], _, _, _ = g.step(cacts[s])
out.update(craft_obs_sens=spread(tobs, cobs), craft_rew_sens=spread(trew, crew))
print("crafted pickup: held", cheld, "phase", cphase, "reward", crew.round(1), "status", env.component_status["cpu"])
out.update(craft_qpos=cq, craft_qvel=cv, craft_warm=cw, craft_actions=cacts, craft_obs=cobs, craft_rew=crew, craft_term=cterm, craft_held=cheld, craft_phase=cphase)
path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "robotic_arm_assembly.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
