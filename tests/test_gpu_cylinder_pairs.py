"""The cylinder pair functions on the device (plane-cylinder, cylinder-box, cylinder-cylinder, capsule-cylinder,
sphere-cylinder) against the fp64 oracle on small hand-built scenes: contact lists bit-exact, distances to 5e-6, one
`mj_step` (Newton) within 1e-4.  The host build of the same code is swept in tests/test_collide_host.py; this test runs
the compiled device code, including plane-cylinder which no task golden happens to contain."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCENE = """<mujoco><option timestep="0.002" solver="Newton" iterations="50" tolerance="1e-10"/>
<default><joint armature="0.01" damping="0.5"/><geom contype="1" conaffinity="1" friction="0.9 0.05 0.05" margin="0.005"/></default>
<worldbody>
  <geom name="floor" type="plane" size="5 5 0.1"/>
  <geom name="slab" type="box" pos="2 0 0.05" size="0.6 0.6 0.05"/>
  <body name="c1" pos="{p1}" euler="{e1}"><freejoint/><geom name="cyl1" type="cylinder" size="0.2 0.5" mass="20"/></body>
  <body name="c2" pos="{p2}" euler="{e2}"><freejoint/><geom name="cyl2" type="cylinder" size="0.15 0.3" mass="5"/></body>
  <body name="cap" pos="{p3}" euler="{e3}"><freejoint/><geom name="capsule" type="capsule" size="0.05 0.3" mass="2"/></body>
  <body name="ball" pos="{p4}"><freejoint/><geom name="sphere" type="sphere" size="0.1" mass="1"/></body>
</worldbody></mujoco>"""

CASES = [
    # upright cylinder resting in the floor, second one lying on its side on the floor, capsule across the first, ball on its cap
    dict(p1="0 0 0.497", e1="0 0 0", p2="1 0 0.148", e2="90 0 0", p3="0.22 0 0.6", e3="90 0 0", p4="0 0 1.095"),
    # tilted cylinder touching the floor with its rim, the other standing on the slab, capsule end-on on the slab cylinder's cap
    dict(p1="0 0 0.52", e1="20 10 0", p2="2 0 0.398", e2="0 0 0", p3="2 0 1.04", e3="0 0 0", p4="-1 0 0.098"),
    # two cylinders side by side (parallel axes) and a capsule leaning on one; ball wedged between cylinder and floor
    dict(p1="0 0 0.499", e1="0 0 0", p2="0.345 0 0.299", e2="0 0 0", p3="-0.24 0 0.5", e3="0 15 0", p4="0 0.29 0.099"),
    # crossed cylinders (one lying on the other), capsule on the slab, ball on the floor
    dict(p1="0 0 0.198", e1="90 0 0", p2="0 0 0.545", e2="0 90 0", p3="2 0.2 0.148", e3="90 0 0", p4="1 1 0.0999"),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_cylinder_scene_matches_oracle(case):
    import torch
    from mujoco_gymnasium_environments_b200 import capi, mjcf
    from oracle import ref
    t = mjcf.compile_mjcf(SCENE.format(**CASES[case]), name="cyl%d" % case)
    om = ref.load_model(t); d = ref.RefData(om)
    ref.mj_forward(om, d)
    opairs = [(c.geom1, c.geom2) for c in d.contact]; odist = np.array([c.dist for c in d.contact])
    m = capi.DeviceModel(t, 0); b = capi.Batch(m, None, 2, 0, 0, con_cap=48)
    ncon, geom, dist = b.contacts(48)
    dbg = b.debug_forward()
    torch.cuda.synchronize()
    n = int(ncon[0])
    gt = np.asarray(t.geom_type)
    kinds = sorted({(int(gt[a]), int(gt[b_])) for a, b_ in opairs})
    print("case", case, "contacts", n, "pair types", kinds)
    assert n == len(opairs) and n >= 4
    assert [tuple(x) for x in geom[0, :n].cpu().numpy().tolist()] == opairs
    assert np.allclose(dist[0, :n].cpu().numpy(), odist, atol=5e-6)
    assert int(dbg["nefc"][0]) == d.nefc
    qa = dbg["qacc"][0].cpu().numpy()
    assert np.max(np.abs(qa - d.qacc)) / (np.max(np.abs(d.qacc)) + 1e-12) < 1e-3
    # MPR stops on an absolute 1e-6 test, so its normal on curved or flat-on-flat pairs answers to the last bits of the poses:
    # the step is compared up to the oracle's own response to an fp32-sized perturbation of the state (oracle/twin.py)
    from oracle.twin import SLACK, perturbed
    q0, v0 = d.qpos.copy(), d.qvel.copy(); prng = np.random.default_rng(4); sq = 0.0; sv = 0.0
    b.physics_step(1)
    ref.mj_step(om, d)
    for _ in range(6):
        g = ref.RefData(om); g.qpos[:] = perturbed(q0, prng); g.qvel[:] = perturbed(v0, prng)
        ref.mj_step(om, g); sq = np.maximum(sq, SLACK * np.abs(g.qpos - d.qpos)); sv = np.maximum(sv, SLACK * np.abs(g.qvel - d.qvel))
    st = b.get_state()
    assert np.max(np.maximum(np.abs(st["qpos"][0].cpu().numpy() - d.qpos) - sq, 0)) / np.max(np.abs(d.qpos)) < 1e-4
    assert np.max(np.maximum(np.abs(st["qvel"][0].cpu().numpy() - d.qvel) - sv, 0)) / (np.max(np.abs(d.qvel)) + 1e-12) < 1e-4
    b.close()
