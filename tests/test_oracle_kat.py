"""Analytic known-answer tests that pin the fp64 oracle (oracle/mjstep_ref.c).

The reference repository holds no golden vector for mj_step and MuJoCo itself cannot be imported here or on the GPU
box ("parity unpinned"), so the oracle is pinned against closed-form results of the published
MuJoCo model instead: discrete free fall, pendulum period, conservation laws, soft-contact equilibrium from
solref/solimp, joint-limit equilibrium, and the implicit-damping Euler update.
"""
import math

import numpy as np
import pytest

from mujoco_gymnasium_environments_b200.mjcf import compile_mjcf
from oracle import ref
import kat_models as K


def make(xml):
    t = compile_mjcf(xml)
    m = ref.load_model(t)
    return t, m, ref.RefData(m)


def energy(t, d):
    ke = 0.5 * d.qvel @ d.M @ d.qvel
    pe = -sum(t.body_mass[b] * np.dot(t.gravity, d.xipos[b]) for b in range(t.nbody))
    return ke + pe


def test_free_fall_semi_implicit_euler():
    t, m, d = make(K.FREE_SPHERE.format(integ="Euler"))
    n, h = 1000, 0.001
    ref.mj_step(m, d, n)
    assert d.qpos[2] == pytest.approx(5 - 9.81 * h * h * n * (n + 1) / 2, abs=1e-10)
    assert d.qvel[2] == pytest.approx(-9.81 * n * h, abs=1e-10)
    assert d.time == pytest.approx(1.0)


def test_free_fall_rk4_is_exact_parabola():
    t, m, d = make(K.FREE_SPHERE.format(integ="RK4"))
    ref.mj_step(m, d, 500)
    assert d.qpos[2] == pytest.approx(5 - 0.5 * 9.81 * 0.25, abs=1e-10)


def test_pendulum_period_and_energy():
    t, m, d = make(K.PENDULUM)
    d.qpos[0] = 0.1
    ref.mj_forward(m, d); e0 = energy(t, d)
    prev, cross = d.qpos[0], []
    for i in range(10000):
        ref.mj_step(m, d)
        if prev > 0 >= d.qpos[0]:
            cross.append((i + 1) * 0.0005)
        prev = d.qpos[0]
    I = t.body_inertia[1][1] + 1.0          # about the pivot: I_com + m l^2, m = 1, l = 1
    T = 2 * math.pi * math.sqrt(I / 9.81) * (1 + 0.1 ** 2 / 16)
    assert np.diff(cross).mean() == pytest.approx(T, abs=1e-3)
    ref.mj_forward(m, d)
    assert energy(t, d) == pytest.approx(e0, abs=1e-9)


@pytest.mark.parametrize("g", ["0 0 0", "0 0 -9.81"])
def test_branched_floating_chain_conserves_energy_and_momentum(g):
    t, m, d = make(K.FLOATING_CHAIN.format(g=g))
    d.qvel[:] = np.random.default_rng(0).normal(size=t.nv)
    ref.mj_forward(m, d); e0 = energy(t, d); c0 = d.subtree_com[1].copy()
    ref.mj_step(m, d, 1000); c1 = d.subtree_com[1].copy()
    ref.mj_step(m, d, 1000); c2 = d.subtree_com[1].copy()
    ref.mj_forward(m, d)
    assert energy(t, d) == pytest.approx(e0, rel=1e-7)
    gz = float(g.split()[2])
    # centre of mass: uniform motion plus free fall; second difference isolates the acceleration
    acc = (c2 - 2 * c1 + c0) / 0.25
    assert np.allclose(acc, [0, 0, gz], atol=1e-6)
    assert np.all(np.linalg.eigvalsh(d.M) > 0) and np.allclose(d.M, d.M.T)


def _contact_equilibrium(mass, mu, tran, g=9.81, nrows=4):
    """Penetration r solving nrows * K*imp(r)*r / R(r) = m g for the default solref/solimp at timestep 0.002."""
    dmin, dmax, width, mid, power = 0.9, 0.95, 0.001, 0.5, 2.0
    tc, dr = 0.02, 1.0
    Kk = 1.0 / (dmax * dmax * tc * tc * dr * dr)

    def imp(r):
        x = min(r / width, 1.0)
        y = x * x / mid if x <= mid else 1 - (1 - x) ** 2 / (1 - mid)
        return dmin + y * (dmax - dmin)

    def total(r):
        i = imp(r)
        R = (1 - i) / i * tran * (1 + mu * mu)
        R = 2 * mu * mu * R
        return nrows * Kk * i * r / R - mass * g

    lo, hi = 0.0, 0.05
    for _ in range(200):
        mid_ = 0.5 * (lo + hi)
        lo, hi = (mid_, hi) if total(mid_) < 0 else (lo, mid_)
    return 0.5 * (lo + hi)


@pytest.mark.parametrize("solver", ["PGS", "Newton"])
def test_sphere_on_plane_rest_penetration(solver):
    t, m, d = make(K.SPHERE_ON_PLANE.format(solver=solver))
    ref.mj_step(m, d, 3000)
    assert d.ncon == 1 and d.nefc == 4
    c = d.contact[0]
    assert (c.geom1, c.geom2) == (0, 1)                     # plane first (lower geom type)
    assert np.allclose(c.frame[0:3], [0, 0, 1])
    mass = t.body_mass[1]
    r = _contact_equilibrium(mass, mu=1.0, tran=1.0 / mass)   # friction = max(1.0 plane default, 0.8)
    assert -c.dist == pytest.approx(r, rel=1e-4)
    assert np.abs(d.qvel).max() < 1e-6
    f = d.efc_force
    assert np.allclose(f, mass * 9.81 / 4, rtol=1e-4)         # four pyramid rows share the load
    assert d.qfrc_constraint[2] == pytest.approx(mass * 9.81, rel=1e-4)


def test_pgs_and_newton_agree_when_converged():
    out = {}
    for solver in ("PGS", "Newton"):
        t, m, d = make(K.BOX_ON_PLANE.format(solver=solver).replace('iterations="100"', 'iterations="2000" tolerance="1e-14"'))
        ref.mj_step(m, d, 400)
        out[solver] = (d.qpos.copy(), d.qvel.copy(), d.ncon)
    assert out["PGS"][2] == out["Newton"][2]
    assert np.allclose(out["PGS"][0], out["Newton"][0], atol=2e-5)


def test_box_and_capsule_rest_contacts():
    t, m, d = make(K.BOX_ON_PLANE.format(solver="Newton"))
    ref.mj_step(m, d, 2500)
    pairs = [(c.geom1, c.geom2) for c in d.contact]
    assert pairs.count((0, 1)) == 4 and pairs.count((0, 2)) == 2     # box: 4 corners; capsule: both end spheres
    assert pairs == sorted(pairs)                                     # body-major contact order
    assert all(c.dist < 0 for c in d.contact)
    # the capsule's tangent hint is its axis: frame row 1 parallel to the (horizontal) capsule axis
    cap_axis = d.geom_xmat[2].reshape(3, 3)[:, 2]
    con = [c for c in d.contact if c.geom2 == 2][0]
    assert abs(abs(np.dot(con.frame[3:6], cap_axis)) - 1) < 1e-3
    assert np.abs(d.qvel).max() < 1e-4


def test_joint_limit_equilibrium():
    t, m, d = make(K.LIMITED_HINGE.format(solver="Newton"))
    ref.mj_step(m, d, 6000)
    assert d.nefc == 1 and np.abs(d.qvel).max() < 1e-7
    q = d.qpos[0]
    assert q > 0.5                                       # gravity pushes the arm into the upper limit
    r = q - 0.5
    dmin, dmax, width, midp = 0.9, 0.95, 0.001, 0.5
    x = min(r / width, 1.0); y = x * x / midp if x <= midp else 1 - (1 - x) ** 2 / (1 - midp)
    imp = dmin + y * (dmax - dmin)
    Kk = 1 / (dmax ** 2 * 0.02 ** 2)
    R = (1 - imp) / imp * t.dof_invweight0[0]
    force = Kk * imp * r / R
    com = t.body_ipos[1][0]
    torque = t.body_mass[1] * 9.81 * com * math.cos(q)
    assert force == pytest.approx(torque, rel=1e-5)
    assert d.efc_force[0] == pytest.approx(force, rel=1e-5)


def test_euler_implicit_damping_and_ctrl_clamp():
    t, m, d = make(K.LIMITED_HINGE.format(solver="Newton").replace('gravity="0 0 -9.81"', 'gravity="0 0 0"'))
    d.qvel[0] = 0.3; d.ctrl[0] = 5.0                      # clamped to ctrlrange 1 -> torque gear*1 = 2
    ref.mj_forward(m, d)
    I = d.M[0, 0]; h, damp = 0.002, 0.5
    assert d.qfrc_actuator[0] == pytest.approx(2.0)
    assert d.ctrl[0] == 5.0                               # the clamp does not modify data.ctrl
    ref.mj_step(m, d)
    assert d.qvel[0] == pytest.approx(0.3 + h * (2.0 - damp * 0.3) / (I + h * damp), rel=1e-12)
    assert d.qpos[0] == pytest.approx(h * d.qvel[0], rel=1e-12)   # position uses the new velocity
    assert d.qacc_warmstart[0] == pytest.approx((2.0 - damp * 0.3) / I, rel=1e-12)   # un-damped qacc is what is saved


def test_reset_data_and_bad_state_autoreset(quad_tables):
    m = ref.load_model(quad_tables); d = ref.RefData(m)
    d.qpos[2] = 0.6; d.ctrl[:] = 1.0; d.qvel[:] = 0.1
    ref.mj_step(m, d, 3)
    ref.mj_resetData(m, d)
    assert np.array_equal(d.qpos, quad_tables.qpos0) and not d.qvel.any() and not d.ctrl.any() and d.time == 0
    d.qvel[3] = float("nan")
    ref.mj_step(m, d)
    assert d.nwarn == 1 and np.isfinite(d.qpos).all() and np.isfinite(d.qvel).all()


def test_quadruped_two_coincident_planes_double_contacts(quad_tables):
    # `floor` and `course_floor` coincide, every ground touch yields two contacts
    m = ref.load_model(quad_tables); d = ref.RefData(m)
    d.qpos[0:3] = [2, 0, 0.6]
    ref.mj_step(m, d, 400)
    pairs = [(c.geom1, c.geom2) for c in d.contact]
    feet = [quad_tables.name2id("geom", n) for n in ("fl_foot", "fr_foot", "bl_foot", "br_foot")]
    for g in feet:
        assert (0, g) in pairs and (1, g) in pairs
    assert pairs[0][0] == 0 and pairs[1][0] == 1 and pairs[0][1] == pairs[1][1]


def test_warmstart_is_saved_by_every_forward_pass():
    """MuJoCo 3.x mj_fwdConstraint ends with qacc_warmstart <- qacc (qacc_smooth when there are no rows), so mj_forward moves
    the warm start and RK4 stages 2-4 start from the previous stage; the switch restores the once-per-step reading."""
    xml = K.BOX_ON_PLANE.format(solver="PGS").replace('iterations="100"', 'iterations="3" integrator="RK4"')
    t, m, d = make(xml)
    ref.mj_step(m, d, 150)                                  # box and capsule resting on the plane: contact rows, PGS-3 unconverged
    assert d.nefc > 0
    q, v, w = d.qpos.copy(), d.qvel.copy(), d.qacc_warmstart.copy()
    ref.mj_forward(m, d)
    assert np.array_equal(d.qacc_warmstart, d.qacc) and not np.array_equal(d.qacc_warmstart, w)
    out = []
    for once in (False, True):
        e = ref.RefData(m); e.set_warmstart_once_per_step(once)
        e.qpos[:] = q; e.qvel[:] = v; e.qacc_warmstart[:] = w
        ref.mj_step(m, e)
        assert np.array_equal(e.qacc_warmstart, e.qacc)     # both readings leave the last stage's qacc behind
        out.append(e.qvel.copy())
    assert not np.allclose(out[0], out[1], rtol=0, atol=1e-9)   # the stages' warm starts differ, and PGS-3 does not forget it
    e = ref.RefData(m); e.set_warmstart_once_per_step(True)
    e.qpos[:] = q; e.qvel[:] = v; e.qacc_warmstart[:] = w
    ref.mj_forward(m, e)
    assert np.array_equal(e.qacc_warmstart, w)
    # no constraint rows: the warm start becomes qacc_smooth
    t2, m2, d2 = make(K.FREE_SPHERE.format(integ="Euler"))
    d2.qacc_warmstart[:] = 7.0
    ref.mj_forward(m2, d2)
    assert d2.nefc == 0 and np.array_equal(d2.qacc_warmstart, d2.qacc_smooth)
