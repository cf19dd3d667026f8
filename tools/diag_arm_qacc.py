"""Diagnostic: arm golden states on the GPU vs the oracle, per-dof qacc differences and the oracle's convex-pair contacts."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from mujoco_gymnasium_environments_b200 import capi
from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
from oracle import ref
t = load_tables("robotic_arm_assembly")
gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "robotic_arm_assembly.npz"))
m = capi.DeviceModel(t, 0); n = gold["qpos"].shape[0]
b = capi.Batch(m, TASKS["robotic_arm_assembly"].describe(t), n, 0, 0)
f = lambda k: torch.tensor(gold[k], dtype=torch.float32)
b.set_state(f("qpos"), f("qvel"), f("ctrl"), f("warm"), torch.zeros(n))
dbg = b.debug_forward(); torch.cuda.synchronize()
om = ref.load_model(t)
gt = np.asarray(t.geom_type)
dof_body = np.asarray(t.dof_bodyid)
for k in range(n):
    d = ref.RefData(om)
    d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
    ref.mj_forward(om, d)
    qa = dbg["qacc"][k].cpu().numpy(); dq = np.abs(qa - d.qacc); sc = np.abs(d.qacc).max()
    print("state", k, "ncon", d.ncon, "nefc", d.nefc, "max|qacc|", sc, "rel", dq.max() / sc, "iters gpu", int(dbg["solver_iter"][k]), "oracle", d.solver_iter)
    for i in np.argsort(-dq)[:8]:
        print("   dof %2d body %-18s gpu %12.5f oracle %12.5f  diff %.3e" % (i, t.id2name("body", int(dof_body[i])) if hasattr(t, "id2name") else dof_body[i], qa[i], d.qacc[i], dq[i]))
    ep = dbg["efc_pos"][k].cpu().numpy()[:d.nefc]; dp = np.abs(ep - d.efc_pos[:d.nefc])
    print("   max |efc_pos diff| %.3e at row %d" % (dp.max(), dp.argmax()))
    for c in d.contact:
        if gt[c.geom1] == 5 or (gt[c.geom2] == 5 and gt[c.geom1] != 0 and gt[c.geom1] != 2):
            print("   convex %-20s %-20s dist %+.6f pos %s n %s" % (t.id2name("geom", c.geom1) if hasattr(t, "id2name") else c.geom1, t.id2name("geom", c.geom2) if hasattr(t, "id2name") else c.geom2, c.dist, np.round(c.pos, 5), np.round(c.frame[:3], 5)))
