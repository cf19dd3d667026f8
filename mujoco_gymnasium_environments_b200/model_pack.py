"""Pack :class:`ModelTables` into the two flat buffers the C-ABI takes (``b2_model_create``).

Layout (declared for C in ``include/b2_model_layout.h``, generated from the lists below by
``python -m mujoco_gymnasium_environments_b200.model_pack --emit-header``):

  ints  : [B2_MAGIC, n_int_fields, n_flt_fields, 0,
           (offset,count) * n_int_fields,   -- offsets into ``ints``
           (offset,count) * n_flt_fields,   -- offsets into ``floats``
           payload...]
  floats: payload (fp64 for the oracle, converted to fp32 by the library for the device)

Every field starts on a 4-element boundary so device copies can use 128-bit loads.
"""
from __future__ import annotations

import sys
from typing import Dict, List, Tuple

import numpy as np

from .mjcf import ModelTables, JNT_FREE

B2_MAGIC = 0x42324D4A  # "B2MJ"

DIM_NAMES = ["nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "ntree", "npair", "nM", "iterations",
             "ls_iterations", "solver", "integrator", "maxdepth", "maxcon", "nlimited", "nisland"]
OPT_NAMES = ["timestep", "gravity_x", "gravity_y", "gravity_z", "tolerance", "ls_tolerance", "impratio",
             "meaninertia"]

INT_FIELDS = [
    "dims", "body_parentid", "body_weldid", "body_rootid", "body_depth", "body_treeid", "body_jntadr",
    "body_jntnum", "body_dofadr", "body_dofnum", "body_geomadr", "body_geomnum", "body_lastdof",
    "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_limited",
    "dof_bodyid", "dof_jntid", "dof_parentid", "dof_treeid", "dof_Madr", "dof_depth",
    "geom_type", "geom_bodyid", "site_bodyid",
    "act_dofid", "act_ctrllimited", "act_forcelimited",
    "pair_g1", "pair_g2", "pair_condim", "pair_maxcon",
    "level_adr", "level_body",          # bodies grouped by depth: level l = level_body[level_adr[l]:level_adr[l+1]]
    "body_childadr", "body_childnum", "body_child",
    "tree_dofadr", "tree_dofnum", "tree_island", "limited_jnt",
]
FLT_FIELDS = [
    "opt", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_subtreemass", "body_inertia",
    "body_invweight0", "jnt_pos", "jnt_axis", "jnt_range", "jnt_stiffness", "jnt_margin", "jnt_solref",
    "jnt_solimp", "qpos0", "qpos_spring", "dof_armature", "dof_damping", "dof_invweight0",
    "geom_size", "geom_pos", "geom_quat", "geom_rbound", "site_pos", "site_quat",
    "act_gear", "act_ctrlrange", "act_forcerange", "act_gainprm", "act_biasprm",
    "pair_friction", "pair_margin", "pair_gap", "pair_solref", "pair_solimp",
]

# static upper bound on contacts one geom pair can generate, by (type1,type2) with type1<=type2
# (plane0 sphere2 capsule3 cylinder5 box6); SURVEY App. B.4
_MAXCON = {(0, 2): 1, (0, 3): 2, (0, 5): 4, (0, 6): 4, (2, 2): 1, (2, 3): 1, (2, 5): 1, (2, 6): 1,
           (3, 3): 2, (3, 5): 1, (3, 6): 2, (5, 5): 1, (5, 6): 1, (6, 6): 8}


def derived_tables(m: ModelTables) -> Dict[str, np.ndarray]:
    """Integer tables the kernels need that are pure functions of the tree structure."""
    A = m.arrays
    nbody, nv, njnt, ntree, npair = int(A["nbody"]), int(A["nv"]), int(A["njnt"]), int(A["ntree"]), int(A["npair"])
    D: Dict[str, np.ndarray] = {}
    depth = A["body_depth"]
    maxdepth = int(depth.max()) if nbody > 1 else 0
    level_adr = [0]; level_body = []
    for l in range(maxdepth + 1):
        level_body += [b for b in range(nbody) if depth[b] == l]
        level_adr.append(len(level_body))
    D["level_adr"] = np.array(level_adr, np.int32); D["level_body"] = np.array(level_body, np.int32)
    childadr = np.zeros(nbody, np.int32); childnum = np.zeros(nbody, np.int32); child = []
    for b in range(nbody):
        childadr[b] = len(child)
        ch = [c for c in range(1, nbody) if A["body_parentid"][c] == b]
        childnum[b] = len(ch); child += ch
    D["body_childadr"] = childadr; D["body_childnum"] = childnum; D["body_child"] = np.array(child, np.int32)
    lastdof = np.full(nbody, -1, np.int32)
    for b in range(1, nbody):
        if A["body_dofnum"][b] > 0:
            lastdof[b] = A["body_dofadr"][b] + A["body_dofnum"][b] - 1
        else:
            lastdof[b] = lastdof[A["body_parentid"][b]]
    D["body_lastdof"] = lastdof
    dof_depth = np.zeros(nv, np.int32)
    for d in range(nv):
        p = A["dof_parentid"][d]
        dof_depth[d] = 0 if p < 0 else dof_depth[p] + 1
    D["dof_depth"] = dof_depth
    tree_dofadr = np.zeros(ntree, np.int32); tree_dofnum = np.zeros(ntree, np.int32)
    for t in range(ntree):
        idx = np.nonzero(A["dof_treeid"] == t)[0]
        tree_dofadr[t] = idx[0]; tree_dofnum[t] = len(idx)
        assert np.all(idx == np.arange(idx[0], idx[0] + len(idx))), "tree dofs must be contiguous"
    D["tree_dofadr"] = tree_dofadr; D["tree_dofnum"] = tree_dofnum
    # static islands: trees that can ever exchange a contact force (union over candidate pairs)
    label = np.arange(ntree)
    def find(x):
        while label[x] != x:
            x = label[x]
        return x
    for p in range(npair):
        t1 = A["body_treeid"][A["geom_bodyid"][A["pair_g1"][p]]]
        t2 = A["body_treeid"][A["geom_bodyid"][A["pair_g2"][p]]]
        if t1 >= 0 and t2 >= 0:
            a, b = find(t1), find(t2)
            if a != b:
                label[max(a, b)] = min(a, b)
    roots = sorted({find(t) for t in range(ntree)})
    D["tree_island"] = np.array([roots.index(find(t)) for t in range(ntree)], np.int32)
    nisland = len(roots)
    pmax = np.zeros(npair, np.int32)
    for p in range(npair):
        key = (int(A["geom_type"][A["pair_g1"][p]]), int(A["geom_type"][A["pair_g2"][p]]))
        pmax[p] = _MAXCON.get(key, 0)
    D["pair_maxcon"] = pmax
    lim = [j for j in range(njnt) if A["jnt_limited"][j] and A["jnt_type"][j] != JNT_FREE]
    D["limited_jnt"] = np.array(lim, np.int32)
    D["dims"] = np.array([int(A[k]) for k in DIM_NAMES[:14]] + [maxdepth, int(pmax.sum()), len(lim), nisland], np.int32)
    return D


def pack_model(m: ModelTables) -> Tuple[np.ndarray, np.ndarray]:
    A = dict(m.arrays)
    A.update(derived_tables(m))
    g = A["gravity"]
    A["opt"] = np.array([float(A["timestep"]), g[0], g[1], g[2], float(A["tolerance"]), float(A["ls_tolerance"]),
                         float(A["impratio"]), float(A["meaninertia"])])
    ni, nf = len(INT_FIELDS), len(FLT_FIELDS)
    head = 4 + 2 * (ni + nf)
    ints: List[np.ndarray] = []; flts: List[np.ndarray] = []
    table = np.zeros((ni + nf, 2), np.int32)
    off = (head + 3) // 4 * 4
    for k, name in enumerate(INT_FIELDS):
        a = np.ascontiguousarray(A[name], np.int32).ravel()
        table[k] = (off, a.size)
        pad = (-a.size) % 4
        ints.append(np.concatenate([a, np.zeros(pad, np.int32)])); off += a.size + pad
    foff = 0
    for k, name in enumerate(FLT_FIELDS):
        a = np.ascontiguousarray(A[name], np.float64).ravel()
        table[ni + k] = (foff, a.size)
        pad = (-a.size) % 4
        flts.append(np.concatenate([a, np.zeros(pad)])); foff += a.size + pad
    headarr = np.zeros((head + 3) // 4 * 4, np.int32)
    headarr[0:4] = (B2_MAGIC, ni, nf, 0)
    headarr[4:4 + 2 * (ni + nf)] = table.ravel()
    return np.concatenate([headarr] + ints), np.concatenate(flts)


def emit_header() -> str:
    lines = ["/* GENERATED by mujoco_gymnasium_environments_b200/model_pack.py --emit-header; do not edit. */",
             "#ifndef B2_MODEL_LAYOUT_H", "#define B2_MODEL_LAYOUT_H", "",
             f"#define B2_MAGIC 0x{B2_MAGIC:08X}", f"#define B2_N_INT_FIELDS {len(INT_FIELDS)}",
             f"#define B2_N_FLT_FIELDS {len(FLT_FIELDS)}", "", "enum b2_int_field {"]
    lines += [f"  B2I_{n} = {k}," for k, n in enumerate(INT_FIELDS)]
    lines += ["};", "", "enum b2_flt_field {"]
    lines += [f"  B2F_{n} = {k}," for k, n in enumerate(FLT_FIELDS)]
    lines += ["};", "", "enum b2_dim {"]
    lines += [f"  B2D_{n} = {k}," for k, n in enumerate(DIM_NAMES)]
    lines += ["};", "", "enum b2_opt {"]
    lines += [f"  B2O_{n} = {k}," for k, n in enumerate(OPT_NAMES)]
    lines += ["};", "",
              "/* joint / geom / solver / integrator enums (values follow MuJoCo's mjtJoint, mjtGeom, ...) */",
              "enum { B2_JNT_FREE = 0, B2_JNT_BALL = 1, B2_JNT_SLIDE = 2, B2_JNT_HINGE = 3 };",
              "enum { B2_GEOM_PLANE = 0, B2_GEOM_HFIELD = 1, B2_GEOM_SPHERE = 2, B2_GEOM_CAPSULE = 3,",
              "       B2_GEOM_ELLIPSOID = 4, B2_GEOM_CYLINDER = 5, B2_GEOM_BOX = 6 };",
              "enum { B2_SOLVER_PGS = 0, B2_SOLVER_CG = 1, B2_SOLVER_NEWTON = 2 };",
              "enum { B2_INT_EULER = 0, B2_INT_RK4 = 1 };", "",
              "/* field k of the int buffer: offset = ints[4+2k], count = ints[5+2k];",
              "   field k of the float buffer: offset = ints[4+2(B2_N_INT_FIELDS+k)], count = ints[5+2(...)] */",
              "#define B2_INT_OFF(ints, k) ((ints)[4 + 2 * (k)])", "#define B2_INT_CNT(ints, k) ((ints)[5 + 2 * (k)])",
              "#define B2_FLT_OFF(ints, k) ((ints)[4 + 2 * (B2_N_INT_FIELDS + (k))])",
              "#define B2_FLT_CNT(ints, k) ((ints)[5 + 2 * (B2_N_INT_FIELDS + (k))])", "",
              "#endif", ""]
    return "\n".join(lines)


if __name__ == "__main__":
    if "--emit-header" in sys.argv:
        sys.stdout.write(emit_header())
