"""Model compiler (MJCF subset -> tables): dimensions, id ordering, inertia rules, collision filtering."""
import math
import os

import numpy as np
import pytest

from mujoco_gymnasium_environments_b200 import mjcf
from mujoco_gymnasium_environments_b200.device_pack import build_device_tables, pack_device_model
from mujoco_gymnasium_environments_b200.model_pack import pack_model, derived_tables
import kat_models as K


def test_quadruped_dimensions(quad_tables):
    t = quad_tables
    # nq 38 / nv 37 / nu 31 / nbody 26 / njnt 32 / ngeom 108
    assert (t.nq, t.nv, t.nu, t.nbody, t.njnt, t.ngeom) == (38, 37, 31, 26, 32, 108)
    assert t.npair == 48                      # F7: only plane pairs survive the bitmask
    assert t.names["geom"][0] == "floor" and t.names["geom"][1] == "course_floor"
    assert [t.name2id("body", n) for n in ("torso", "fl_foot", "fr_foot", "bl_foot", "br_foot")] == [1, 5, 9, 13, 17]
    assert t.name2id("joint", "platform_slide") == 17 and t.name2id("joint", "pendulum_swing") == 18
    assert t.name2id("actuator", "platform_motor") == 16 and t.name2id("actuator", "pendulum_motor") == 17
    assert t.name2id("body", "nope") == -1
    assert t.timestep == pytest.approx(0.001) and t.solver == mjcf.SOLVER_PGS and t.iterations == 50
    assert t.integrator == mjcf.INT_EULER


def test_quadruped_pairs_are_plane_pairs_in_contact_order(quad_tables):
    t = quad_tables
    assert set(t.geom_type[t.pair_g1].tolist()) == {mjcf.GEOM_PLANE}
    counts = {}
    for g in t.pair_g2:
        counts[int(t.geom_type[g])] = counts.get(int(t.geom_type[g]), 0) + 1
    assert counts == {mjcf.GEOM_BOX: 14, mjcf.GEOM_CAPSULE: 24, mjcf.GEOM_SPHERE: 10}
    # body-major, then plane 0 before plane 1 (naive MuJoCo order)
    b = t.geom_bodyid[t.pair_g2]
    assert np.all(np.diff(b) >= 0)
    assert t.pair_g1[:4].tolist() == [0, 1, 0, 1]


def test_inertia_from_geom_overrides_inertial(quad_tables):
    t = quad_tables
    # torso box 0.4x0.15x0.1 half sizes at density 1000 -> 48 kg, not the <inertial mass="25">
    assert t.body_mass[1] == pytest.approx(8 * 0.4 * 0.15 * 0.1 * 1000)
    r = 0.03
    assert t.body_mass[5] == pytest.approx(4 / 3 * math.pi * r ** 3 * 1000)
    # joint defaults of the base file apply to appended joints (armature 0.01)
    assert np.allclose(t.dof_armature, 0.01)


def test_primitive_inertia_formulas():
    m, I = mjcf._geom_mass_inertia(mjcf.GEOM_BOX, np.array([0.1, 0.2, 0.3]), 1000.0, None)
    assert m == pytest.approx(48.0) and I[0] == pytest.approx(48 / 3 * (0.04 + 0.09))
    m, I = mjcf._geom_mass_inertia(mjcf.GEOM_SPHERE, np.array([0.5, 0, 0]), 1.0, "2.0")
    assert m == 2.0 and np.allclose(I, 0.4 * 2.0 * 0.25)
    # capsule: numeric integration of the solid of revolution
    r, h = 0.05, 0.2
    m, I = mjcf._geom_mass_inertia(mjcf.GEOM_CAPSULE, np.array([r, h, 0]), 1000.0, None)
    z = np.linspace(-h - r, h + r, 200001)
    rho = np.where(np.abs(z) <= h, r, np.sqrt(np.maximum(r * r - (np.abs(z) - h) ** 2, 0)))
    dm = 1000.0 * math.pi * rho ** 2
    assert m == pytest.approx(np.trapezoid(dm, z), rel=1e-6)
    assert I[2] == pytest.approx(np.trapezoid(0.5 * dm * rho ** 2, z), rel=1e-5)
    assert I[0] == pytest.approx(np.trapezoid(dm * (0.25 * rho ** 2 + z ** 2), z), rel=1e-5)


def test_fromto_points_z_from_to_towards_from():
    t = mjcf.compile_mjcf("""<mujoco><worldbody><body><joint/><geom type="capsule" size="0.1" fromto="0 0 0 0 0.4 0"/></body></worldbody></mujoco>""")
    R = mjcf.quat_to_mat(t.geom_quat[0])
    assert np.allclose(R[:, 2], [0, -1, 0], atol=1e-12)
    assert np.allclose(t.geom_pos[0], [0, 0.2, 0]) and t.geom_size[0][1] == pytest.approx(0.2)


def test_dense_mass_matrix_matches_oracle_crb():
    from oracle import ref
    t = mjcf.compile_mjcf(K.FLOATING_CHAIN.format(g="0 0 -9.81"))
    m = ref.load_model(t); d = ref.RefData(m)
    ref.mj_forward(m, d)
    assert np.allclose(d.M, t.M0, rtol=1e-10, atol=1e-12)
    assert np.all(np.linalg.eigvalsh(d.M) > 0)


def test_setconst_invweight(quad_tables):
    t = quad_tables
    Minv = np.linalg.inv(t.M0)
    assert t.dof_invweight0[0] == pytest.approx(np.mean(np.diag(Minv)[0:3]))
    assert t.dof_invweight0[10] == pytest.approx(Minv[10, 10])
    assert t.body_invweight0[0].tolist() == [0, 0] and t.body_invweight0[19].tolist() == [0, 0]   # world, static anchor
    assert t.meaninertia == pytest.approx(np.mean(np.diag(t.M0)))


def test_pack_roundtrip_and_layout(quad_tables):
    ints, flts = pack_model(quad_tables)
    from mujoco_gymnasium_environments_b200.model_pack import B2_MAGIC, INT_FIELDS, FLT_FIELDS
    assert ints[0] == B2_MAGIC and ints[1] == len(INT_FIELDS) and ints[2] == len(FLT_FIELDS)
    k = INT_FIELDS.index("body_parentid"); off, cnt = ints[4 + 2 * k], ints[5 + 2 * k]
    assert np.array_equal(ints[off:off + cnt], quad_tables.body_parentid) and off % 4 == 0
    di, df = pack_device_model(quad_tables)
    assert di.size % 4 == 0 and df.size % 4 == 0
    T = build_device_tables(quad_tables)
    assert T["tree_island"].tolist() if "tree_island" in T else True
    assert T["dims"][11] == 8          # eight islands: nothing but planes can touch anything
    # generated headers are in sync with the field lists
    from mujoco_gymnasium_environments_b200 import model_pack, device_pack
    root = os.path.join(os.path.dirname(__file__), "..", "include")
    assert open(os.path.join(root, "b2_model_layout.h")).read() == model_pack.emit_header()
    assert open(os.path.join(root, "b2_device_layout.h")).read() == device_pack.emit_header()


def test_compile_errors_are_value_errors():
    with pytest.raises(ValueError):
        mjcf.compile_mjcf("<mujoco><worldbody><body><joint type='ball'/><geom size='1'/></body></worldbody></mujoco>")
    with pytest.raises(ValueError):
        mjcf.compile_mjcf("<notmujoco/>")
