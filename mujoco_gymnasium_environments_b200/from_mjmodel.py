"""``ModelTables`` straight from a compiled ``mujoco.MjModel`` -- the fast path of SURVEY.md section 8(f)1 for boxes where
MuJoCo is importable: no MJCF parsing of our own, and MuJoCo's own ``mj_setConst`` constants (``dof_invweight0``,
``body_invweight0``, ``stat.meaninertia``) instead of the restated ones, so any compiler difference is out of the picture.

    model = mujoco.MjModel.from_xml_string(xml)          # what every reference env does (parkour_env.py:100 ...)
    tables = from_mjmodel(model, name="quadruped_parkour")
    env = B200VectorEnv.from_tables(tables, ...)           # or capi.DeviceModel(tables, device)

Only attribute access on ``model`` is used (``model.nq``, ``model.body_pos`` ...), so the function never imports mujoco itself
and the CPU test drives it with a stand-in object built from the committed tables.  Unsupported features are refused loudly
(ball joints, tendons, equality constraints, mesh / hfield / ellipsoid geoms, non-joint transmissions, elliptic cones).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .mjcf import GEOM_PLANE, ModelTables, _candidate_pairs, _set_const

_SUPPORTED_GEOMS = {0, 2, 3, 5, 6}     # mjGEOM_PLANE, SPHERE, CAPSULE, CYLINDER, BOX (the tables keep MuJoCo's own type codes)


def _names(model, kind: str, n: int) -> List[str]:
    """Names through ``model.<kind>(i).name`` (MuJoCo >= 2.3 named access); empty strings when unavailable."""
    out = []
    for i in range(n):
        try:
            out.append(str(getattr(model, kind)(i).name))
        except Exception:
            out.append("")
    return out


def from_mjmodel(model, name: str = "") -> ModelTables:
    def _arr(x, dt):
        a = np.array(x, dt)
        return a if a.ndim == 0 else np.ascontiguousarray(a)
    i32 = lambda x: _arr(x, np.int32)
    f64 = lambda x: _arr(x, np.float64)
    nq, nv, nu, nbody, njnt, ngeom = (int(getattr(model, k)) for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom"))
    nsite = int(getattr(model, "nsite", 0))
    for attr, what in (("ntendon", "tendons"), ("neq", "equality constraints"), ("nmesh", "meshes"), ("nhfield", "height fields")):
        if int(getattr(model, attr, 0)) > 0:
            raise NotImplementedError(f"from_mjmodel: the model has {what}, which the engine does not implement")
    jt = np.array(model.jnt_type, int)
    if np.any(jt == 1):
        raise NotImplementedError("from_mjmodel: ball joints are not implemented")
    gt = np.array(model.geom_type, int)
    if not set(gt.tolist()) <= _SUPPORTED_GEOMS:
        raise NotImplementedError(f"from_mjmodel: unsupported geom types {sorted(set(gt.tolist()) - _SUPPORTED_GEOMS)}")
    opt = model.opt
    if int(getattr(opt, "cone", 0)) != 0:
        raise NotImplementedError("from_mjmodel: elliptic friction cones are not implemented (pyramidal only)")
    if int(opt.integrator) not in (0, 1) or int(opt.solver) not in (0, 2):
        raise NotImplementedError("from_mjmodel: integrator must be Euler or RK4 and solver PGS or Newton")
    A: Dict[str, np.ndarray] = {}
    parent = i32(model.body_parentid); weld = i32(model.body_weldid)
    depth = np.zeros(nbody, np.int32)
    for b in range(1, nbody):
        depth[b] = depth[parent[b]] + 1
    # kinematic trees: bodies whose weld root hangs off the world, numbered in body order; static bodies -1
    dof_body = i32(model.dof_bodyid); treeid = np.full(nbody, -1, np.int32); ntree = 0
    body_dofnum = i32(model.body_dofnum); body_dofadr = i32(model.body_dofadr)
    for b in range(1, nbody):
        if weld[b] == 0:
            continue
        p = parent[weld[b]] if weld[b] == b else b
        up = parent[b]
        if treeid[up] >= 0 and weld[up] != 0:
            treeid[b] = treeid[up]
        else:
            treeid[b] = ntree; ntree += 1
    A.update(nq=i32(nq), nv=i32(nv), nu=i32(nu), nbody=i32(nbody), njnt=i32(njnt), ngeom=i32(ngeom), nsite=i32(nsite), ntree=i32(ntree))
    A.update(timestep=f64(opt.timestep), gravity=f64(opt.gravity), iterations=i32(opt.iterations), tolerance=f64(opt.tolerance),
             ls_iterations=i32(getattr(opt, "ls_iterations", 50)), ls_tolerance=f64(getattr(opt, "ls_tolerance", 0.01)),
             solver=i32(opt.solver), integrator=i32(opt.integrator), impratio=f64(opt.impratio))
    A.update(body_parentid=parent, body_weldid=weld, body_rootid=i32(model.body_rootid), body_depth=depth, body_treeid=treeid,
             body_jntadr=i32(model.body_jntadr), body_jntnum=i32(model.body_jntnum), body_dofadr=body_dofadr, body_dofnum=body_dofnum,
             body_geomadr=i32(model.body_geomadr), body_geomnum=i32(model.body_geomnum), body_pos=f64(model.body_pos),
             body_quat=f64(model.body_quat), body_ipos=f64(model.body_ipos), body_iquat=f64(model.body_iquat), body_mass=f64(model.body_mass),
             body_subtreemass=f64(model.body_subtreemass), body_inertia=f64(model.body_inertia))
    A.update(jnt_type=i32(jt), jnt_qposadr=i32(model.jnt_qposadr), jnt_dofadr=i32(model.jnt_dofadr),
             jnt_bodyid=i32(model.jnt_bodyid), jnt_pos=f64(model.jnt_pos), jnt_axis=f64(model.jnt_axis), jnt_limited=i32(model.jnt_limited),
             jnt_range=f64(model.jnt_range), jnt_stiffness=f64(model.jnt_stiffness), jnt_margin=f64(model.jnt_margin),
             jnt_solref=f64(model.jnt_solref), jnt_solimp=f64(model.jnt_solimp), qpos0=f64(model.qpos0), qpos_spring=f64(model.qpos_spring))
    A.update(dof_bodyid=dof_body, dof_jntid=i32(model.dof_jntid), dof_parentid=i32(model.dof_parentid), dof_treeid=i32([treeid[b] for b in dof_body]),
             dof_armature=f64(model.dof_armature), dof_damping=f64(model.dof_damping))
    g = dict(type=[int(t) for t in gt], bodyid=i32(model.geom_bodyid), contype=i32(model.geom_contype), conaffinity=i32(model.geom_conaffinity),
             condim=i32(model.geom_condim), friction=f64(model.geom_friction), margin=f64(model.geom_margin), gap=f64(model.geom_gap),
             solref=f64(model.geom_solref), solimp=f64(model.geom_solimp), solmix=f64(model.geom_solmix), priority=i32(model.geom_priority))
    A.update(geom_type=i32(g["type"]), geom_bodyid=g["bodyid"], geom_size=f64(model.geom_size), geom_pos=f64(model.geom_pos), geom_quat=f64(model.geom_quat),
             geom_contype=g["contype"], geom_conaffinity=g["conaffinity"], geom_condim=g["condim"], geom_friction=g["friction"],
             geom_margin=g["margin"], geom_gap=g["gap"], geom_rbound=f64(model.geom_rbound))
    A.update(site_bodyid=i32(getattr(model, "site_bodyid", [])), site_pos=f64(getattr(model, "site_pos", np.zeros((0, 3)))).reshape(nsite, 3),
             site_quat=f64(getattr(model, "site_quat", np.zeros((0, 4)))).reshape(nsite, 4))
    # actuators: joint transmissions only; gain = gainprm[0] * ctrl, bias = biasprm[0] + biasprm[1] q + biasprm[2] qdot
    trn = np.array(getattr(model, "actuator_trntype", np.zeros(nu)), int)
    if np.any(trn != 0):
        raise NotImplementedError("from_mjmodel: only joint transmissions are implemented")
    jid = np.array(model.actuator_trnid, int).reshape(nu, 2)[:, 0] if nu else np.zeros(0, int)
    A.update(act_dofid=i32([model.jnt_dofadr[j] for j in jid]), act_gear=f64(np.array(model.actuator_gear).reshape(nu, -1)[:, 0] if nu else []),
             act_ctrllimited=i32(model.actuator_ctrllimited), act_ctrlrange=f64(model.actuator_ctrlrange).reshape(nu, 2),
             act_forcelimited=i32(model.actuator_forcelimited), act_forcerange=f64(model.actuator_forcerange).reshape(nu, 2),
             act_gainprm=f64(np.array(model.actuator_gainprm).reshape(nu, -1)[:, 0] if nu else []),
             act_biasprm=f64(np.array(model.actuator_biasprm).reshape(nu, -1)[:, :3] if nu else np.zeros((0, 3))))
    # explicit <pair>s and <exclude>s, then the static collision filter in MuJoCo's contact order
    explicit = {}
    for k in range(int(getattr(model, "npair", 0))):
        g1, g2 = int(model.pair_geom1[k]), int(model.pair_geom2[k])
        explicit[(min(g1, g2), max(g1, g2))] = dict(g1=g1, g2=g2, condim=int(model.pair_dim[k]), friction=f64(model.pair_friction[k]),
                                                    margin=float(model.pair_margin[k]), gap=float(model.pair_gap[k]),
                                                    solref=f64(model.pair_solref[k]), solimp=f64(model.pair_solimp[k]))
    excludes = set()
    for sig in np.array(getattr(model, "exclude_signature", []), np.int64).tolist():
        b1, b2 = sig >> 16, sig & 0xffff
        excludes.add((min(b1, b2), max(b1, b2)))
    pairs = _candidate_pairs(nbody, parent, weld, A["body_geomadr"], A["body_geomnum"], g, explicit, excludes)
    npair = len(pairs["g1"])
    A.update(npair=i32(npair), pair_g1=i32(pairs["g1"]), pair_g2=i32(pairs["g2"]), pair_condim=i32(pairs["condim"]),
             pair_friction=f64(pairs["friction"]).reshape(npair, 5), pair_margin=f64(pairs["margin"]), pair_gap=f64(pairs["gap"]),
             pair_solref=f64(pairs["solref"]).reshape(npair, 2), pair_solimp=f64(pairs["solimp"]).reshape(npair, 5))
    t = ModelTables(name=name, arrays=A, names=dict(body=_names(model, "body", nbody), joint=_names(model, "joint", njnt), geom=_names(model, "geom", ngeom),
                                                    site=_names(model, "site", nsite), actuator=_names(model, "actuator", nu)))
    _set_const(t)                                   # sparse-M addressing and M0; the weights below are MuJoCo's own mj_setConst values
    A["dof_invweight0"] = f64(model.dof_invweight0); A["body_invweight0"] = f64(model.body_invweight0).reshape(nbody, 2)
    A["meaninertia"] = np.array(float(model.stat.meaninertia))
    return t
