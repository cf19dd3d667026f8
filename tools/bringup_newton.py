"""GPU bring-up of the Newton solver: small known-answer models and the construction model, CUDA vs fp64 oracle."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
from mujoco_gymnasium_environments_b200 import capi, mjcf
from oracle import ref
import kat_models as K


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


def run(name, xml, nsteps_list, ctrl=None, perturb=None):
    t = mjcf.compile_mjcf(xml, name=name)
    dm = capi.DeviceModel(t, 0)
    b = capi.Batch(dm, None, 1, 0, 0)
    om = ref.load_model(t); d = ref.RefData(om)
    if perturb is not None:
        perturb(d)
    if ctrl is not None:
        d.ctrl[:] = ctrl
    q = torch.tensor(d.qpos.copy()[None], dtype=torch.float32); v = torch.tensor(d.qvel.copy()[None], dtype=torch.float32)
    c = torch.tensor(d.ctrl.copy()[None], dtype=torch.float32) if t.nu else None
    b.set_state(q, v, c, torch.zeros((1, int(t.nv))), torch.zeros(1))
    d.qpos[:] = q[0].numpy(); d.qvel[:] = v[0].numpy()
    done = 0
    for n in nsteps_list:
        b.physics_step(n - done); ref.mj_step(om, d, n - done); done = n
        torch.cuda.synchronize()
        st = b.get_state()
        print(f"{name}: after {n} steps qpos {relerr(st['qpos'][0].cpu(), d.qpos):.2e} qvel {relerr(st['qvel'][0].cpu(), d.qvel):.2e} "
              f"(abs {float(np.max(np.abs(st['qvel'][0].cpu().numpy() - d.qvel))):.2e}) ncon {d.ncon} nefc {d.nefc} iters {d.solver_iter}")
    dbg = b.debug_forward(); torch.cuda.synchronize(); ref.mj_forward(om, d)
    print(f"   forward: qacc {relerr(dbg['qacc'][0].cpu(), d.qacc):.2e} qfrc_c {relerr(dbg['qfrc_constraint'][0].cpu(), d.qfrc_constraint):.2e} "
          f"nefc {int(dbg['nefc'][0])}/{d.nefc} iters {int(dbg['solver_iter'][0])}/{d.solver_iter} stats {b.stats().cpu().numpy()[3:7]}")
    b.close(); dm.close()


run("sphere_on_plane", K.SPHERE_ON_PLANE.format(solver="Newton"), [1, 10, 200])
run("box_on_plane", K.BOX_ON_PLANE.format(solver="Newton"), [1, 10, 300])
run("limited_hinge", K.LIMITED_HINGE.format(solver="Newton"), [1, 50, 400], ctrl=[1.0])
if len(sys.argv) > 1:
    from mujoco_gymnasium_environments_b200.tasks import load_tables
    xml = open(sys.argv[1]).read()
    run("construction", xml, [1, 5, 20])
