"""Device-side model tables: what ``b2_model_create`` uploads and the step kernel stages into shared memory.

Same container format as :mod:`model_pack` (header of (offset,count) pairs, 4-element aligned fields) but a
different, trimmed field list: only what the CUDA kernels read, plus tables that are pure functions of the
tree structure and exist to make the warp/CTA-parallel passes branch-free (level lists, child lists, sparse-M
column ids, descendant lists for the level-synchronous L'DL solves, chain bitmasks for Jacobian rows,
collision-geom subset with precomputed local rotation matrices, de-duplicated contact parameter rows).

``python -m mujoco_gymnasium_environments_b200.device_pack --emit-header`` regenerates
``include/b2_device_layout.h``.
"""
from __future__ import annotations

import sys
from typing import Dict, List, Tuple

import numpy as np

from .mjcf import ModelTables, JNT_FREE, JNT_HINGE, JNT_SLIDE, quat_to_mat
from .model_pack import derived_tables

B2DEV_MAGIC = 0x42324456  # "B2DV"

DIMS = ["nq", "nv", "nu", "nbody", "njnt", "ntree", "nroot", "ncg", "npair", "nM", "nlim", "nisland", "maxdepth",
        "maxdofdepth", "iterations", "integrator", "solver", "nprm", "nmaskw", "ndesc", "maxraw", "nsite",
        "ls_iterations", "maxspan", "npairtab", "nfacstep"]
OPTS = ["timestep", "gx", "gy", "gz", "tolerance", "impratio", "meaninertia", "pgs_scale", "ls_tolerance"]

INT_FIELDS = [
    "dims",
    "body_parentid", "body_rootidx", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum", "body_lastdof",
    "body_childadr", "body_childnum", "body_child", "body_island", "body_tree", "level_adr", "level_body",
    "root_bodyadr", "root_bodynum",
    "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid",
    "dof_bodyid", "dof_Madr", "dof_depth", "dof_isrot", "dof_island", "dof_tree", "dof_actadr", "dof_actnum",
    "dof_jnt", "dofact", "Mcol", "dlevel_adr", "dlevel_dof", "dof_descadr", "dof_descnum", "desc_pack", "pair_mn",
    "bw_pack", "fw_pack", "fac_step", "fac_ops",
    "tree_dofadr", "tree_dofnum", "island_dofadr", "island_dofnum",
    "act_dofid", "act_ctrllimited", "act_forcelimited",
    "lim_jnt",
    "cg_type", "cg_body", "cg_geomid",
    "body_chainmask", "site_bodyid",
    # candidate-pair tables come last: kernels whose task sets COLD_PAIRS leave them in global memory (L2) instead of
    # staging them into shared memory (rescue: 3 175 pairs = 64 KB); header word 3 holds the offset where they start
    "pair_cg1", "pair_cg2", "pair_prm", "pair_rawadr", "pair_maxcon",
]
N_COLD_INT_FIELDS = 5
FLT_FIELDS = [
    "opt",
    "body_pos", "body_quat", "body_ipos", "body_imat", "body_mass", "body_inertia", "body_invweight0",
    "root_invmass",
    "jnt_pos", "jnt_axis", "jnt_range", "jnt_stiffness", "jnt_margin", "jnt_solprm",
    "qpos0", "qpos_spring",
    "dof_armature", "dof_damping", "dof_invweight0",
    "cg_size", "cg_pos", "cg_mat", "cg_rbound",
    "prm",
    "act_gear", "act_ctrlrange", "act_forcerange", "act_gain", "act_bias",
    "site_pos",
]
PRM_STRIDE = 16  # margin, gap, mu[5], solref[2], solimp[5], condim, pad


def build_device_tables(m: ModelTables) -> Dict[str, np.ndarray]:
    A = m.arrays
    D = derived_tables(m)
    nbody, nv, njnt, ntree, npair, nu = (int(A[k]) for k in ("nbody", "nv", "njnt", "ntree", "npair", "nu"))
    T: Dict[str, np.ndarray] = {}
    i32 = lambda x: np.ascontiguousarray(np.array(x, np.int32).ravel())
    f64 = lambda x: np.ascontiguousarray(np.array(x, np.float64).ravel())

    # roots: direct children of the world; the com of each root's subtree is the spatial-algebra origin of its bodies
    roots = [b for b in range(1, nbody) if A["body_parentid"][b] == 0]
    rootidx = np.full(nbody, -1, np.int32)
    root_adr, root_num, root_invmass = [], [], []
    for r, rb in enumerate(roots):
        members = [b for b in range(1, nbody) if A["body_rootid"][b] == rb]
        assert members == list(range(rb, rb + len(members))), "root subtree must be contiguous in body order"
        rootidx[members] = r
        root_adr.append(rb); root_num.append(len(members))
        sm = A["body_subtreemass"][rb]
        root_invmass.append(1.0 / sm if sm > 1e-15 else 0.0)
    nroot = len(roots)
    island_of_tree = D["tree_island"]
    body_island = np.array([island_of_tree[t] if t >= 0 else -1 for t in A["body_treeid"]], np.int32)

    T["body_parentid"] = i32(A["body_parentid"]); T["body_rootidx"] = i32(rootidx)
    for k in ("body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum"):
        T[k] = i32(A[k])
    for k in ("body_lastdof", "body_childadr", "body_childnum", "body_child", "level_adr", "level_body"):
        T[k] = i32(D[k])
    T["body_island"] = i32(body_island)
    T["body_tree"] = i32(A["body_treeid"])          # -1 for static bodies; islands are formed per step from the active contacts
    T["root_bodyadr"] = i32(root_adr); T["root_bodynum"] = i32(root_num)
    T["jnt_type"] = i32(A["jnt_type"]); T["jnt_qposadr"] = i32(A["jnt_qposadr"]); T["jnt_dofadr"] = i32(A["jnt_dofadr"])
    T["jnt_bodyid"] = i32(A["jnt_bodyid"])

    # dofs
    T["dof_bodyid"] = i32(A["dof_bodyid"]); T["dof_Madr"] = i32(A["dof_Madr"]); T["dof_depth"] = i32(D["dof_depth"])
    T["dof_jnt"] = i32(A["dof_jntid"])
    isrot = np.zeros(nv, np.int32)
    for d in range(nv):
        j = A["dof_jntid"][d]; t = A["jnt_type"][j]
        if t == JNT_HINGE or (t == JNT_FREE and d - A["jnt_dofadr"][j] >= 3):
            isrot[d] = 1
    T["dof_isrot"] = isrot
    T["dof_tree"] = i32(A["dof_treeid"]); T["dof_island"] = i32([island_of_tree[t] for t in A["dof_treeid"]])
    nM = int(A["nM"])
    Mcol = np.zeros(nM, np.int32)
    for i in range(nv):
        j = i; k = 0
        while j >= 0:
            Mcol[A["dof_Madr"][i] + k] = j; k += 1; j = A["dof_parentid"][j]
    T["Mcol"] = Mcol
    maxdd = int(D["dof_depth"].max()) if nv else 0
    dl_adr = [0]; dl = []
    for l in range(maxdd + 1):
        dl += [d for d in range(nv) if D["dof_depth"][d] == l]
        dl_adr.append(len(dl))
    T["dlevel_adr"] = i32(dl_adr); T["dlevel_dof"] = i32(dl)
    # descendants of each dof j with the address of L(i,j)
    desc_adr = np.zeros(nv, np.int32); desc_num = np.zeros(nv, np.int32); desc_dof = []; desc_madr = []
    for j in range(nv):
        desc_adr[j] = len(desc_dof)
        for i in range(j + 1, nv):
            a = A["dof_Madr"][i]; n = D["dof_depth"][i] + 1
            for k in range(1, n):
                if Mcol[a + k] == j:
                    desc_dof.append(i); desc_madr.append(a + k)
        desc_num[j] = len(desc_dof) - desc_adr[j]
    T["dof_descadr"] = desc_adr; T["dof_descnum"] = desc_num
    # one int per (descendant dof, address of L(i,j)) pair: dof | madr << 16
    assert nM < 65536 and nv < 65536
    T["desc_pack"] = i32([(d | (a << 16)) for d, a in zip(desc_dof, desc_madr)] if desc_dof else [0])
    # (m, n) with 1 <= m <= n <= maxdepth enumerated n-major: pair p of the rank-1 update of the L'DL factorisation
    pair_mn = []
    for n in range(1, maxdd + 1):
        for mm in range(1, n + 1):
            pair_mn.append(mm | (n << 8))
    T["pair_mn"] = i32(pair_mn if pair_mn else [0])
    # level-ordered packed per-dof records for the L'DL solves (same order as dlevel_dof):
    #   bw_pack = dof | n_desc << 8 | desc_adr << 16      fw_pack = dof | Madr << 8
    assert nv < 256 and len(desc_dof) < 65536 and max(desc_num, default=0) < 256
    T["bw_pack"] = i32([(d | (int(desc_num[d]) << 8) | (int(desc_adr[d]) << 16)) for d in dl] if dl else [0])
    T["fw_pack"] = i32([(d | (int(A["dof_Madr"][d]) << 8)) for d in dl] if dl else [0])
    # flat program of the L'DL factorisation: one step per eliminated dof k (leaves first) with depth > 0,
    #   fac_step = diag_adr | n_ops << 10 | op_adr << 18 ; fac_ops = tgt | a << 10 | b << 20  (LD[tgt] -= LD[a]*LD[b]/LD[diag])
    assert nM < 1024
    fac_step, fac_ops = [], []
    for k in range(nv - 1, -1, -1):
        ak = int(A["dof_Madr"][k]); dk = int(D["dof_depth"][k])
        if dk == 0:
            continue
        start = len(fac_ops)
        for n in range(1, dk + 1):
            for mm in range(1, n + 1):
                aa = int(A["dof_Madr"][Mcol[ak + mm]])
                fac_ops.append((aa + (n - mm)) | ((ak + mm) << 10) | ((ak + n) << 20))
        cnt = len(fac_ops) - start
        assert cnt < 256 and start < 8192
        fac_step.append(ak | (cnt << 10) | (start << 18))
    T["fac_step"] = i32(fac_step if fac_step else [0]); T["fac_ops"] = i32(fac_ops if fac_ops else [0])
    ndesc = len(desc_dof)
    T["tree_dofadr"] = i32(D["tree_dofadr"]); T["tree_dofnum"] = i32(D["tree_dofnum"])
    nisland = int(island_of_tree.max()) + 1 if ntree else 0
    isl_adr = np.zeros(max(nisland, 1), np.int32); isl_num = np.zeros(max(nisland, 1), np.int32)
    for k in range(nisland):
        ds = np.nonzero(T["dof_island"] == k)[0]
        isl_adr[k] = ds.min(); isl_num[k] = ds.max() - ds.min() + 1
    T["island_dofadr"] = isl_adr; T["island_dofnum"] = isl_num

    # actuators per dof
    T["act_dofid"] = i32(A["act_dofid"]) if nu else i32([0])
    T["act_ctrllimited"] = i32(A["act_ctrllimited"]) if nu else i32([0])
    T["act_forcelimited"] = i32(A["act_forcelimited"]) if nu else i32([0])
    actadr = np.zeros(nv, np.int32); actnum = np.zeros(nv, np.int32); dofact = []
    for d in range(nv):
        actadr[d] = len(dofact)
        ids = [a for a in range(nu) if A["act_dofid"][a] == d]
        dofact += ids; actnum[d] = len(ids)
    T["dof_actadr"] = actadr; T["dof_actnum"] = actnum; T["dofact"] = i32(dofact if dofact else [0])
    T["lim_jnt"] = i32(D["limited_jnt"]) if len(D["limited_jnt"]) else i32([0])
    nlim = len(D["limited_jnt"])

    # collision geoms: only geoms that occur in a candidate pair
    used = sorted(set(A["pair_g1"].tolist()) | set(A["pair_g2"].tolist()))
    cgidx = {g: k for k, g in enumerate(used)}
    ncg = len(used)
    cg_type, cg_body, cg_size, cg_pos, cg_mat, cg_rb = [], [], [], [], [], []
    for g in used:
        cg_type.append(A["geom_type"][g]); cg_body.append(A["geom_bodyid"][g]); cg_size.append(A["geom_size"][g])
        cg_pos.append(A["geom_pos"][g]); cg_mat.append(quat_to_mat(A["geom_quat"][g]).ravel())
        cg_rb.append(A["geom_rbound"][g])
        # geoms of static non-world bodies (goal posts, ...) keep their body: the kinematics pass computes every
        # body frame, static ones included; their island is -1 and their chain mask empty, so they only push back
    T["cg_type"] = i32(cg_type if ncg else [0]); T["cg_body"] = i32(cg_body if ncg else [0])
    T["cg_geomid"] = i32(used if ncg else [0])
    T["cg_size"] = f64(cg_size if ncg else [0] * 3); T["cg_pos"] = f64(cg_pos if ncg else [0] * 3)
    T["cg_mat"] = f64(cg_mat if ncg else [0] * 9); T["cg_rbound"] = f64(cg_rb if ncg else [0])
    # contact parameter rows, de-duplicated
    prm_rows: List[Tuple] = []; pair_prm = []
    for p in range(npair):
        row = tuple(np.concatenate([[A["pair_margin"][p], A["pair_gap"][p]], A["pair_friction"][p], A["pair_solref"][p],
                                    A["pair_solimp"][p], [float(A["pair_condim"][p]), 0.0]]).tolist())
        if row not in prm_rows:
            prm_rows.append(row)
        pair_prm.append(prm_rows.index(row))
    T["pair_cg1"] = i32([cgidx[g] for g in A["pair_g1"]] if npair else [0])
    T["pair_cg2"] = i32([cgidx[g] for g in A["pair_g2"]] if npair else [0])
    T["pair_prm"] = i32(pair_prm if npair else [0])
    pmax = D["pair_maxcon"]
    T["pair_maxcon"] = i32(pmax if npair else [0])
    T["pair_rawadr"] = i32(np.concatenate([[0], np.cumsum(pmax)[:-1]]) if npair else [0])
    maxraw = int(pmax.sum())
    T["prm"] = f64(prm_rows if prm_rows else [[0.0] * PRM_STRIDE])
    nprm = max(len(prm_rows), 1)

    # chain masks: bit d of body b set when dof d lies on the path from b to the world
    nmaskw = max((nv + 31) // 32, 1)
    mask = np.zeros((nbody, nmaskw), np.uint32)
    parent_dof = A["dof_parentid"]
    for b in range(1, nbody):
        d = D["body_lastdof"][b]
        while d >= 0:
            mask[b, d // 32] |= np.uint32(1 << (d % 32)); d = parent_dof[d]
    T["body_chainmask"] = mask.view(np.int32).ravel()
    T["site_bodyid"] = i32(A["site_bodyid"]) if int(A["nsite"]) else i32([0])
    T["site_pos"] = f64(A["site_pos"]) if int(A["nsite"]) else f64([0, 0, 0])

    # floats
    T["body_pos"] = f64(A["body_pos"]); T["body_quat"] = f64(A["body_quat"]); T["body_ipos"] = f64(A["body_ipos"])
    T["body_imat"] = f64([quat_to_mat(q).ravel() for q in A["body_iquat"]])
    T["body_mass"] = f64(A["body_mass"]); T["body_inertia"] = f64(A["body_inertia"])
    T["body_invweight0"] = f64(A["body_invweight0"]); T["root_invmass"] = f64(root_invmass if nroot else [0])
    T["jnt_pos"] = f64(A["jnt_pos"]); T["jnt_axis"] = f64(A["jnt_axis"]); T["jnt_range"] = f64(A["jnt_range"])
    T["jnt_stiffness"] = f64(A["jnt_stiffness"]); T["jnt_margin"] = f64(A["jnt_margin"])
    T["jnt_solprm"] = f64(np.concatenate([A["jnt_solref"], A["jnt_solimp"], np.zeros((njnt, 1))], axis=1))
    T["qpos0"] = f64(A["qpos0"]); T["qpos_spring"] = f64(A["qpos_spring"])
    T["dof_armature"] = f64(A["dof_armature"]); T["dof_damping"] = f64(A["dof_damping"])
    T["dof_invweight0"] = f64(A["dof_invweight0"])
    T["act_gear"] = f64(A["act_gear"]) if nu else f64([0]); T["act_ctrlrange"] = f64(A["act_ctrlrange"]) if nu else f64([0, 0])
    T["act_forcerange"] = f64(A["act_forcerange"]) if nu else f64([0, 0])
    T["act_gain"] = f64(A["act_gainprm"]) if nu else f64([0]); T["act_bias"] = f64(A["act_biasprm"]) if nu else f64([0, 0, 0])

    g = A["gravity"]
    mi = float(A["meaninertia"])
    T["opt"] = f64([float(A["timestep"]), g[0], g[1], g[2], float(A["tolerance"]), float(A["impratio"]), mi,
                    1.0 / (mi * max(1, nv)), float(A["ls_tolerance"])])
    T["dims"] = i32([int(A["nq"]), nv, nu, nbody, njnt, ntree, nroot, ncg, npair, nM, nlim, nisland,
                     int(D["dims"][14]), maxdd, int(A["iterations"]), int(A["integrator"]), int(A["solver"]), nprm,
                     nmaskw, ndesc, maxraw, int(A["nsite"]), int(A["ls_iterations"]), int(isl_num.max()) if nisland else 0,
                     len(pair_mn), len(fac_step)])
    return T


def pack_device_model(m: ModelTables) -> Tuple[np.ndarray, np.ndarray]:
    T = build_device_tables(m)
    ni, nf = len(INT_FIELDS), len(FLT_FIELDS)
    head = (4 + 2 * (ni + nf) + 3) // 4 * 4
    table = np.zeros((ni + nf, 2), np.int32)
    ints, flts = [], []
    off = head
    for k, name in enumerate(INT_FIELDS):
        a = np.ascontiguousarray(T[name], np.int32).ravel()
        table[k] = (off, a.size); pad = (-a.size) % 4
        ints.append(np.concatenate([a, np.zeros(pad, np.int32)])); off += a.size + pad
    foff = 0
    for k, name in enumerate(FLT_FIELDS):
        a = np.ascontiguousarray(T[name], np.float64).ravel()
        table[ni + k] = (foff, a.size); pad = (-a.size) % 4
        flts.append(np.concatenate([a, np.zeros(pad)])); foff += a.size + pad
    h = np.zeros(head, np.int32)
    h[0:4] = (B2DEV_MAGIC, ni, nf, int(table[ni - N_COLD_INT_FIELDS, 0]))
    h[4:4 + 2 * (ni + nf)] = table.ravel()
    return np.concatenate([h] + ints), np.concatenate(flts)


def emit_header() -> str:
    L = ["/* GENERATED by mujoco_gymnasium_environments_b200/device_pack.py --emit-header; do not edit. */",
         "#ifndef B2_DEVICE_LAYOUT_H", "#define B2_DEVICE_LAYOUT_H", "",
         f"#define B2DEV_MAGIC 0x{B2DEV_MAGIC:08X}", f"#define B2DEV_N_INT_FIELDS {len(INT_FIELDS)}",
         f"#define B2DEV_N_FLT_FIELDS {len(FLT_FIELDS)}", f"#define B2DEV_PRM_STRIDE {PRM_STRIDE}", "",
         "enum b2dev_int_field {"]
    L += [f"  DI_{n} = {k}," for k, n in enumerate(INT_FIELDS)]
    L += ["};", "", "enum b2dev_flt_field {"]
    L += [f"  DF_{n} = {k}," for k, n in enumerate(FLT_FIELDS)]
    L += ["};", "", "enum b2dev_dim {"]
    L += [f"  DD_{n} = {k}," for k, n in enumerate(DIMS)]
    L += [f"  DD_COUNT = {len(DIMS)}", "};", "", "enum b2dev_opt {"]
    L += [f"  DO_{n} = {k}," for k, n in enumerate(OPTS)]
    L += [f"  DO_COUNT = {len(OPTS)}", "};", "", "#endif", ""]
    return "\n".join(L)


if __name__ == "__main__":
    if "--emit-header" in sys.argv:
        sys.stdout.write(emit_header())
