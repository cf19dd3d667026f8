"""Known-answer tests of the oracle's restatement of libccd's Minkowski Portal Refinement (oracle/mjstep_ref.c::mpr_convex,
the path MuJoCo's mjc_Convex takes for capsule-cylinder, cylinder-cylinder and cylinder-box pairs).  There is no MuJoCo to
compare with, so the algorithm is pinned by what it must satisfy whatever the portal sequence:

* sphere-sphere through MPR is the segment case (origin on the v0-v1 line) and equals the closed form to round-off;
* translating geom 2 by depth * normal leaves the two geoms touching: the closed-form distance of the translated pair is
  zero to within mpr_tolerance (1e-6), for every pair type that also has a closed form;
* separated pairs (closed-form distance above the margin) give no contact, and a pair exactly touching gives none either
  (libccd treats a zero support distance as "no intersection");
* a cylinder standing on a box reports the face penetration along the face normal; two crossed cylinders report the
  side-to-side overlap along the common perpendicular.
"""
import ctypes

import numpy as np
import pytest

from mujoco_gymnasium_environments_b200.mjcf import quat_to_mat
from oracle import ref

DP = ctypes.POINTER(ctypes.c_double)
SPHERE, CAPSULE, CYLINDER, BOX = 2, 3, 5, 6
EYE = np.eye(3).ravel()


@pytest.fixture(scope="module")
def R():
    L = ref.lib()
    for f in (L.ref_collide_raw, L.ref_mpr_raw):
        f.argtypes = [ctypes.c_int, ctypes.c_int, DP, DP, DP, DP, DP, DP, ctypes.c_double, DP]
    return L


def call(f, t1, t2, p1, m1, s1, p2, m2, s2, margin):
    a = [np.ascontiguousarray(x, np.float64) for x in (p1, m1, s1, p2, m2, s2)]
    o = np.zeros(80)
    n = f(t1, t2, *[x.ctypes.data_as(DP) for x in a], float(margin), o.ctypes.data_as(DP))
    return n, o[0], o[1:4].copy(), o[4:7].copy()


SIZES = {SPHERE: lambda r: [r.uniform(.03, .2), 0, 0], CAPSULE: lambda r: [r.uniform(.03, .1), r.uniform(.05, .3), 0],
         CYLINDER: lambda r: [r.uniform(.1, .3), r.uniform(.1, .5), 0], BOX: lambda r: list(r.uniform(.03, .25, 3))}


def rq(rng):
    q = rng.normal(size=4); return q / np.linalg.norm(q)


def test_sphere_sphere_is_the_segment_case(R):
    rng = np.random.default_rng(0); n = 0
    for _ in range(500):
        s1, s2 = SIZES[SPHERE](rng), SIZES[SPHERE](rng); p2 = rng.normal(size=3) * 0.15
        ka, da, pa, na = call(R.ref_collide_raw, SPHERE, SPHERE, np.zeros(3), EYE, s1, p2, EYE, s2, 0.01)
        km, dm, pm, nm = call(R.ref_mpr_raw, SPHERE, SPHERE, np.zeros(3), EYE, s1, p2, EYE, s2, 0.01)
        assert ka == km
        if ka:
            n += 1
            assert abs(da - dm) < 1e-14 and np.abs(pa - pm).max() < 1e-14 and np.abs(na - nm).max() < 1e-12
    assert n > 100


@pytest.mark.parametrize("pair", [(SPHERE, CAPSULE), (CAPSULE, CAPSULE), (SPHERE, BOX), (SPHERE, CYLINDER)])
def test_translating_by_the_result_separates_the_pair(R, pair):
    t1, t2 = pair; rng = np.random.default_rng(10 * t1 + t2); n = 0; margin = 0.01
    for _ in range(1500):
        s1, s2 = SIZES[t1](rng), SIZES[t2](rng); m1, m2 = quat_to_mat(rq(rng)).ravel(), quat_to_mat(rq(rng)).ravel()
        p2 = rng.normal(size=3) * 0.25
        ka, da, _, _ = call(R.ref_collide_raw, t1, t2, np.zeros(3), m1, s1, p2, m2, s2, margin)
        km, dm, pm, nm = call(R.ref_mpr_raw, t1, t2, np.zeros(3), m1, s1, p2, m2, s2, margin)
        if not ka:
            assert not km                      # closed-form distance above the margin: MPR finds no intersection
            continue
        if not km:
            assert da > margin - 1e-6          # only a pair within tolerance of the margin may be missed
            continue
        n += 1
        assert abs(np.linalg.norm(nm) - 1) < 1e-12 and dm <= margin
        assert dm <= da + 1e-6                 # MPR's depth is along its own direction, never less than the minimum penetration
        kb, db, _, _ = call(R.ref_collide_raw, t1, t2, np.zeros(3), m1, s1, p2 + nm * (margin - dm), m2, s2, 1.0)
        assert kb and -1.001e-6 < db - margin < 1e-9, (pair, db - margin)
    assert n > 200


def test_exact_touch_is_no_contact(R):
    # stacked coaxial cylinders, faces exactly in one plane (the arm's base plate and shoulder link, complete_model.xml:51-57)
    k, *_ = call(R.ref_mpr_raw, CYLINDER, CYLINDER, [0, 0, 0.05], EYE, [0.15, 0.05, 0], [0, 0, 0.25], EYE, [0.08, 0.15, 0], 0.0)
    assert k == 0
    k, d, p, n = call(R.ref_mpr_raw, CYLINDER, CYLINDER, [0, 0, 0.05], EYE, [0.15, 0.05, 0], [0, 0, 0.249], EYE, [0.08, 0.15, 0], 0.0)
    assert k == 1 and abs(d + 0.001) < 1e-12 and np.allclose(n, [0, 0, 1]) and np.allclose(p, [0, 0, 0.0995])


def test_cylinder_on_box_and_crossed_cylinders(R):
    # upright cylinder sunk 2 mm into the top face of a box, off-centre
    k, d, p, n = call(R.ref_mpr_raw, CYLINDER, BOX, [0.03, -0.02, 0.298], EYE, [0.05, 0.2, 0], [0, 0, 0], EYE, [0.3, 0.3, 0.1], 0.0)
    assert k == 1 and abs(d + 0.002) < 1e-6 and np.abs(n - [0, 0, -1]).max() < 1e-6 and abs(p[2] - 0.099) < 1e-3
    assert np.hypot(p[0] - 0.03, p[1] + 0.02) <= 0.05 + 1e-9          # somewhere under the cap (the barycentric blend pulls it 1/3 mm towards
    # the centres): one contact, as MuJoCo reports
    # crossed cylinders (axes along x and y), 3 mm side-to-side overlap along z
    mx = quat_to_mat(np.array([np.sqrt(.5), 0, np.sqrt(.5), 0])).ravel(); my = quat_to_mat(np.array([np.sqrt(.5), np.sqrt(.5), 0, 0])).ravel()
    k, d, p, n = call(R.ref_mpr_raw, CYLINDER, CYLINDER, [0, 0, 0], mx, [0.1, 0.4, 0], [0.01, 0.02, 0.217], my, [0.12, 0.4, 0], 0.0)
    assert k == 1 and abs(d + 0.003) < 2e-6 and np.abs(n - [0, 0, 1]).max() < 3e-3 and np.abs(p - [0.01, 0.0, 0.0985]).max() < 1e-3


def test_flat_contact_point_can_be_decided_by_rounding(R):
    """A flat cap on a flat face: depth and normal are well defined, but the contact *point* is wherever the ray from the
    centre difference through the origin leaves the portal, and for a thin cap far from the other geom's centre that answers to
    the last bits of the poses (libccd behaves the same; MuJoCo's multiccd option exists for this reason).  Shown on the arm's
    own scene: screw 1 standing on its 2 mm shaft cap in the screw bin (complete_model.xml:211-212), first golden state.
    Parity tests on scenes with such contacts compare up to the oracle's own spread (oracle/twin.py)."""
    import os
    from mujoco_gymnasium_environments_b200.tasks import load_tables
    t = load_tables("robotic_arm_assembly"); om = ref.load_model(t)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "robotic_arm_assembly.npz"))
    d = ref.RefData(om); d.qpos[:] = gold["qpos"][0]; d.qvel[:] = gold["qvel"][0]
    ref.mj_forward(om, d)
    g1 = t.name2id("geom", "screw1_shaft"); g2 = t.name2id("geom", "screw_bin_base")
    size = np.asarray(t.geom_size, np.float64).reshape(-1, 3)
    p1, m1 = np.array(d.geom_xpos[g1]), np.array(d.geom_xmat[g1]).ravel(); p2, m2 = np.array(d.geom_xpos[g2]), np.array(d.geom_xmat[g2]).ravel()
    rng = np.random.default_rng(3); pts = []
    for _ in range(12):
        k, dist, p, n = call(R.ref_mpr_raw, CYLINDER, BOX, p1 + rng.normal(size=3) * 1e-8, m1 + rng.normal(size=9) * 1e-8, size[g1], p2, m2, size[g2], 0.0)
        assert k == 1 and -1e-3 < dist < 0 and np.abs(n - [0, 0, -1]).max() < 1e-6
        pts.append(p[:2])
    pts = np.array(pts)
    assert np.all(np.hypot(pts[:, 0] - p1[0], pts[:, 1] - p1[1]) <= size[g1][0] + 1e-6)          # always under the cap ...
    assert max(np.ptp(pts[:, 0]), np.ptp(pts[:, 1])) > 0.25 * size[g1][0]                          # ... but not at a repeatable point of it
