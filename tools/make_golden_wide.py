"""Generate tests/golden/wide_states.npz: over-capacity physics states for the wide tier, from the fp64 oracle.

The reference puts no cap on data.ncon (quadruped_parkour_env/parkour_env.py:470-485 loops over all of it, :711 penalises
ncon > 8).  Each task is rolled out with full-range uniform actions (the bench's distribution) through the oracle's task
layer; after every control step the state is probed with mj_forward, and states whose contact / row counts exceed the
engine's on-chip capacities (32-64 contacts, 128 rows per PGS island) are kept: a belly-down quadruped, the rescue robot on
top of its victims, the construction humanoid fallen into the materials.  Same caveat as tools/make_golden.py: these
vectors pin the oracle, they are not outputs of MuJoCo.  Run:  python tools/make_golden_wide.py
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle import ref
from oracle.tasks_ref import TASKS

# task -> (min contacts, min rows, states to keep, control steps to search)
WANT = {"humanoid_soccer": (33, 134, 4, 6000), "bipedal_rescue": (56, 230, 4, 600),
        "humanoid_dancing": (0, 69, 3, 4000), "humanoid_construction": (62, 260, 3, 400), "humanoid_martial_arts": (50, 200, 3, 20000)}
VMAX = {"humanoid_martial_arts": 1500, "humanoid_dancing": 2e3}
SUBSTEPS = {}      # probe every physics sub-step of a control step (same ctrl)
CAP = 160
out = {}
for task, (min_con, min_rows, keep, steps) in WANT.items():
    env = TASKS[task](seed=7); rng = np.random.default_rng(7); env.reset()
    om = env.model; lo, hi = env.action_low, env.action_high
    S = dict(qpos=[], qvel=[], ctrl=[], warm=[], ncon=[], nefc=[], pairs=[], dist=[], qacc=[], qpos1=[], qvel1=[])
    last = -100
    for s in range(steps):
        _, _, term, trunc, _ = env.step(rng.uniform(lo, hi))
        d = env.data
        probe = ref.RefData(om)
        probe.qpos[:] = d.qpos; probe.qvel[:] = d.qvel; probe.ctrl[:] = d.ctrl; probe.qacc_warmstart[:] = d.qacc_warmstart
        for sub in range(SUBSTEPS.get(task, 1)):
          if sub:
            ref.mj_step(om, probe)
          if probe.nwarn == 0 and s - last >= 3 and np.all(np.isfinite(probe.qpos)) and np.max(np.abs(probe.qvel)) < VMAX.get(task, 500):
            q = probe.qpos.astype(np.float32); v = probe.qvel.astype(np.float32); c = probe.ctrl.astype(np.float32); w = probe.qacc_warmstart.astype(np.float32)
            e = ref.RefData(om)
            e.qpos[:] = q; e.qvel[:] = v; e.ctrl[:] = c; e.qacc_warmstart[:] = w
            ref.mj_forward(om, e)
            con = e.contact
            if len(con) >= min_con and e.nefc >= min_rows and len(con) <= CAP and e.nwarn == 0:
                pairs = np.full((CAP, 2), -1, np.int32); dist = np.zeros(CAP)
                for i, cc in enumerate(con):
                    pairs[i] = (cc.geom1, cc.geom2); dist[i] = cc.dist
                qacc = e.qacc.copy(); nefc = e.nefc
                e2 = ref.RefData(om)
                e2.qpos[:] = q; e2.qvel[:] = v; e2.ctrl[:] = c; e2.qacc_warmstart[:] = w
                ref.mj_step(om, e2)
                if e2.nwarn == 0:
                    for k, x in (("qpos", q), ("qvel", v), ("ctrl", c), ("warm", w), ("ncon", len(con)), ("nefc", nefc), ("pairs", pairs),
                                 ("dist", dist), ("qacc", qacc), ("qpos1", e2.qpos.copy()), ("qvel1", e2.qvel.copy())):
                        S[k].append(x)
                    last = s
        if term or trunc or env.data.nwarn:
            if env.data.nwarn:
                env = TASKS[task](seed=1000 + s)
            env.reset(); last = -100
        if len(S["qpos"]) >= keep:
            break
    print(task, "kept", len(S["qpos"]), "ncon", S["ncon"], "nefc", S["nefc"])
    for k, x in S.items():
        out[f"{task}__{k}"] = np.array(x)

# The quadruped only reaches > 32 contacts under full-range actions in states that are about to blow up (|qvel| ~ 1e6,
# followed by mj_checkAcc's reset), which say nothing at 1e-4.  Its wide states are authored instead: the robot dropped
# belly-down into the floor with splayed legs (torso 10-17 cm above the plane, random roll / pitch and joint angles), taken
# as is and after one and two physics steps -- 36-46 contacts, 150-190 rows, velocities of a few rad/s.
def quadruped_states():
    from mujoco_gymnasium_environments_b200.tasks import load_tables
    t = load_tables("quadruped_parkour"); om = ref.load_model(t); rng = np.random.default_rng(3)
    S = dict(qpos=[], qvel=[], ctrl=[], warm=[], ncon=[], nefc=[], pairs=[], dist=[], qacc=[], qpos1=[], qvel1=[])
    while len(S["qpos"]) < 6:
        d = ref.RefData(om); ref.mj_resetData(om, d)
        q = d.qpos.copy(); q[2] = rng.uniform(0.10, 0.16)
        a = rng.uniform(-0.3, 0.3, 2); cr, sr, cp, sp = np.cos(a[0] / 2), np.sin(a[0] / 2), np.cos(a[1] / 2), np.sin(a[1] / 2)
        q[3:7] = [cr * cp, sr * cp, cr * sp, -sr * sp]; q[7:19] = rng.uniform(-1.2, 1.2, 12)
        d.qpos[:] = q
        for sub in range(1 + len(S["qpos"]) % 3):
            if sub:
                ref.mj_step(om, d)
        qf = d.qpos.astype(np.float32); vf = d.qvel.astype(np.float32); cf = d.ctrl.astype(np.float32); wf = d.qacc_warmstart.astype(np.float32)
        e = ref.RefData(om); e.qpos[:] = qf; e.qvel[:] = vf; e.ctrl[:] = cf; e.qacc_warmstart[:] = wf
        ref.mj_forward(om, e); con = e.contact
        if len(con) < 36 or e.nefc < 150:
            continue
        pairs = np.full((CAP, 2), -1, np.int32); dist = np.zeros(CAP)
        for i, cc in enumerate(con):
            pairs[i] = (cc.geom1, cc.geom2); dist[i] = cc.dist
        e2 = ref.RefData(om); e2.qpos[:] = qf; e2.qvel[:] = vf; e2.ctrl[:] = cf; e2.qacc_warmstart[:] = wf
        ref.mj_step(om, e2)
        for k, x in (("qpos", qf), ("qvel", vf), ("ctrl", cf), ("warm", wf), ("ncon", len(con)), ("nefc", e.nefc), ("pairs", pairs),
                     ("dist", dist), ("qacc", e.qacc.copy()), ("qpos1", e2.qpos.copy()), ("qvel1", e2.qvel.copy())):
            S[k].append(x)
    print("quadruped_parkour (authored) ncon", S["ncon"], "nefc", S["nefc"])
    return S


for k, x in quadruped_states().items():
    out[f"quadruped_parkour__{k}"] = np.array(x)
p = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "wide_states.npz")
np.savez_compressed(p, **out)
print("wrote", p, os.path.getsize(p), "bytes")
