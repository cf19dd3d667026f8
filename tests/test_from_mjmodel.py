"""from_mjmodel: driven by a stand-in object that exposes the committed tables under MjModel's attribute
names (the tables already use MuJoCo's type codes); the tables it returns must equal the ones the MJCF compiler produced."""
import types

import numpy as np
import pytest

from mujoco_gymnasium_environments_b200.from_mjmodel import from_mjmodel
from mujoco_gymnasium_environments_b200.tasks import load_tables



def fake_mjmodel(t):
    m = types.SimpleNamespace()
    for k, v in t.arrays.items():
        setattr(m, k, v.item() if v.ndim == 0 else v.copy())
    m.opt = types.SimpleNamespace(timestep=t.timestep, gravity=t.gravity, iterations=t.iterations, tolerance=t.tolerance, ls_iterations=t.ls_iterations,
                                  ls_tolerance=t.ls_tolerance, solver=t.solver, integrator=t.integrator, impratio=t.impratio, cone=0)
    m.stat = types.SimpleNamespace(meaninertia=t.meaninertia)
    ng = t.ngeom
    m.geom_solref = np.tile([0.02, 1.0], (ng, 1)); m.geom_solimp = np.tile([0.9, 0.95, 0.001, 0.5, 2.0], (ng, 1))
    m.geom_solmix = np.ones(ng); m.geom_priority = np.zeros(ng, int)
    nu = t.nu
    m.actuator_trntype = np.zeros(nu, int)
    jnt_of_dof = {int(t.jnt_dofadr[j]): j for j in range(t.njnt)}
    m.actuator_trnid = np.array([[jnt_of_dof[int(d)], -1] for d in t.act_dofid]).reshape(nu, 2)
    m.actuator_gear = np.concatenate([t.act_gear.reshape(nu, 1), np.zeros((nu, 5))], axis=1)
    m.actuator_ctrllimited = t.act_ctrllimited; m.actuator_ctrlrange = t.act_ctrlrange
    m.actuator_forcelimited = t.act_forcelimited; m.actuator_forcerange = t.act_forcerange
    m.actuator_gainprm = np.concatenate([t.act_gainprm.reshape(nu, 1), np.zeros((nu, 9))], axis=1)
    m.actuator_biasprm = np.concatenate([t.act_biasprm, np.zeros((nu, 7))], axis=1)
    m.npair = 0; m.ntendon = 0; m.neq = 0; m.nmesh = 0; m.nhfield = 0
    for kind, key in (("body", "body"), ("joint", "joint"), ("geom", "geom"), ("site", "site"), ("actuator", "actuator")):
        setattr(m, kind, (lambda names: (lambda i: types.SimpleNamespace(name=names[i])))(t.names[key]))
    return m


@pytest.mark.parametrize("task", ["quadruped_parkour", "humanoid_soccer", "humanoid_martial_arts"])     # tasks without explicit <pair>s
def test_round_trip_through_a_stand_in_mjmodel(task):
    t = load_tables(task)
    r = from_mjmodel(fake_mjmodel(t), name=task)
    assert set(r.arrays) == set(t.arrays)
    for k, v in t.arrays.items():
        if v.dtype.kind == "f":
            assert np.allclose(r.arrays[k], v, rtol=0, atol=1e-12), k
        else:
            assert np.array_equal(r.arrays[k], v), k
    assert r.names == t.names


def test_unsupported_features_are_refused():
    t = load_tables("quadruped_parkour")
    m = fake_mjmodel(t); m.ntendon = 1
    with pytest.raises(NotImplementedError):
        from_mjmodel(m)
    m = fake_mjmodel(t); m.opt.cone = 1
    with pytest.raises(NotImplementedError):
        from_mjmodel(m)
