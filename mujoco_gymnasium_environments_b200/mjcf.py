"""MJCF-subset model compiler: MJCF text -> flat model tables (numpy, fp64).

This is the Python-side compiler named by BASELINE.json's north_star ("flattens each
task's model into SoA device tables").  It replaces ``mujoco.MjModel.from_xml_string``
at the reference call sites

    quadruped_parkour_env/parkour_env.py:100   humanoid_dancing_env/dancing_env.py:115
    humanoid_soccer_env/soccer_env.py:80       humanoid_martial_arts_env/martial_arts_env.py:103
    humanoid_construction_env/construction_env.py:105   bipedal_rescue_env/rescue_env.py:87
    robotic_arm_assembly_env/assembly_env.py:61

for the feature subset those seven models use: one top-level
<default>, free/hinge/slide joints, plane/sphere/capsule/cylinder/box geoms,
``inertiafromgeom``, motor/position actuators with joint transmission, explicit
contact <pair>s, and the compile-time constants of ``mj_setConst``
(``dof_invweight0``, ``body_invweight0``, ``stat.meaninertia``).

MuJoCo itself is an un-vendored dependency that is importable neither in the
authoring container nor on the GPU box, so every rule below restates the published
MuJoCo compile semantics; nothing here is checked against a live ``MjModel``
("parity unpinned").
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# enums shared with csrc/b2_model.h and oracle/mjstep_ref.c
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX = range(7)
GEOM_NAMES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
              "cylinder": GEOM_CYLINDER, "box": GEOM_BOX, "ellipsoid": GEOM_ELLIPSOID}
SOLVER_PGS, SOLVER_CG, SOLVER_NEWTON = 0, 1, 2
INT_EULER, INT_RK4 = 0, 1
MJMINVAL = 1e-15


# ----------------------------------------------------------------------------- quaternion helpers
def quat_mul(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return np.array([
        a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3],
        a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2],
        a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1],
        a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0]])


def quat_from_axis_angle(axis, angle):
    axis = np.asarray(axis, float)
    n = np.linalg.norm(axis)
    if n < 1e-14:
        return np.array([1.0, 0, 0, 0])
    s = math.sin(angle / 2) / n
    return np.array([math.cos(angle / 2), axis[0]*s, axis[1]*s, axis[2]*s])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w*w + x*x - y*y - z*z, 2*(x*y - w*z), 2*(x*z + w*y)],
        [2*(x*y + w*z), w*w - x*x + y*y - z*z, 2*(y*z - w*x)],
        [2*(x*z - w*y), 2*(y*z + w*x), w*w - x*x - y*y + z*z]])


def quat_normalize(q):
    q = np.asarray(q, float)
    n = np.linalg.norm(q)
    return np.array([1.0, 0, 0, 0]) if n < 1e-14 else q / n


def mat_to_quat(m):
    """Rotation matrix -> unit quaternion (w>=0 branch not enforced)."""
    m = np.asarray(m, float)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    return quat_normalize(q)


def quat_z_to_vec(vec):
    """Quaternion rotating +z onto ``vec`` (the rule MJCF ``fromto`` uses)."""
    v = np.asarray(vec, float)
    n = np.linalg.norm(v)
    if n < 1e-14:
        return np.array([1.0, 0, 0, 0])
    v = v / n
    axis = np.cross([0.0, 0, 1], v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1.0, 0, 0])
    else:
        axis = axis / s
    ang = math.atan2(s, v[2])
    return quat_from_axis_angle(axis, ang)


# ----------------------------------------------------------------------------- model container
@dataclass
class ModelTables:
    """Flat model tables. Every array is C-contiguous; float arrays fp64, ids int32."""
    name: str = ""
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)
    names: Dict[str, List[str]] = field(default_factory=dict)

    def __getattr__(self, k):
        arrays = object.__getattribute__(self, "arrays")
        if k in arrays:
            v = arrays[k]
            if v.ndim == 0:
                return v.item()
            return v
        raise AttributeError(k)

    def name2id(self, kind: str, name: str) -> int:
        """``mj_name2id`` stand-in: -1 when absent (the reference wraps lookups in try/except)."""
        try:
            return self.names[kind].index(name)
        except ValueError:
            return -1

    def id2name(self, kind: str, i: int) -> str:
        return self.names[kind][i]

    def save(self, path: str) -> None:
        meta = {f"names_{k}": np.array(v, dtype=object if False else "U") if v else np.zeros(0, "U1")
                for k, v in self.names.items()}
        np.savez_compressed(path, _model_name=np.array(self.name), **self.arrays, **meta)

    @staticmethod
    def load(path: str) -> "ModelTables":
        z = np.load(path, allow_pickle=False)
        m = ModelTables(name=str(z["_model_name"]))
        for k in z.files:
            if k == "_model_name":
                continue
            if k.startswith("names_"):
                m.names[k[6:]] = [str(s) for s in z[k]]
            else:
                m.arrays[k] = z[k]
        return m


# ----------------------------------------------------------------------------- parsing helpers
def _floats(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, float)
    v = np.array([float(x) for x in s.replace(",", " ").split()], float)
    if n is not None and len(v) < n and default is not None:
        out = np.array(default, float)
        out[:len(v)] = v
        return out
    return v


def _bool(s, default=None):
    if s is None:
        return default
    return s.strip().lower() == "true"


class _Defaults:
    """The single top-level <default> block (App. C.1: no nested classes are used)."""

    def __init__(self, root):
        self.d: Dict[str, Dict[str, str]] = {}
        dn = root.find("default")
        if dn is not None:
            for ch in dn:
                if ch.tag == "default":
                    continue  # nested classes are not used by the seven models
                self.d.setdefault(ch.tag, {}).update(ch.attrib)
        # actuator shortcuts alias the "general" default
        for k in ("motor", "position", "velocity", "general"):
            self.d.setdefault(k, {})

    def get(self, tag, elem, key, fallback=None):
        if key in elem.attrib:
            return elem.attrib[key]
        if tag in ("motor", "position", "velocity", "general"):
            for t in (tag, "general", "motor", "position"):
                if key in self.d.get(t, {}):
                    return self.d[t][key]
            return fallback
        return self.d.get(tag, {}).get(key, fallback)


class MjcfCompileError(ValueError):
    """Mirrors the ``ValueError`` MuJoCo raises on an MJCF compile error."""


def _orientation(elem, angle_scale, eulerseq="xyz"):
    """quat / euler / axisangle / xyaxes / zaxis -> unit quaternion."""
    if "quat" in elem.attrib:
        return quat_normalize(_floats(elem.get("quat")))
    if "euler" in elem.attrib:
        e = _floats(elem.get("euler")) * angle_scale
        q = np.array([1.0, 0, 0, 0])
        for ch, a in zip(eulerseq, e):
            ax = {"x": [1, 0, 0], "y": [0, 1, 0], "z": [0, 0, 1]}[ch.lower()]
            r = quat_from_axis_angle(ax, a)
            q = quat_mul(q, r) if ch.islower() else quat_mul(r, q)
        return quat_normalize(q)
    if "axisangle" in elem.attrib:
        a = _floats(elem.get("axisangle"))
        return quat_from_axis_angle(a[:3], a[3] * angle_scale)
    if "xyaxes" in elem.attrib:
        a = _floats(elem.get("xyaxes"))
        x = a[:3] / np.linalg.norm(a[:3])
        y = a[3:] - x * np.dot(x, a[3:])
        y /= np.linalg.norm(y)
        z = np.cross(x, y)
        return mat_to_quat(np.stack([x, y, z], axis=1))
    if "zaxis" in elem.attrib:
        return quat_z_to_vec(_floats(elem.get("zaxis")))
    return np.array([1.0, 0, 0, 0])


def _geom_mass_inertia(gtype, size, density, mass_attr):
    """Volume-based mass and principal inertia of a primitive (MuJoCo user_objects.cc rules)."""
    if gtype == GEOM_SPHERE:
        vol = 4.0 / 3.0 * math.pi * size[0] ** 3
    elif gtype == GEOM_CAPSULE:
        vol = math.pi * size[0] ** 2 * (2 * size[1]) + 4.0 / 3.0 * math.pi * size[0] ** 3
    elif gtype == GEOM_CYLINDER:
        vol = math.pi * size[0] ** 2 * (2 * size[1])
    elif gtype == GEOM_BOX:
        vol = 8.0 * size[0] * size[1] * size[2]
    elif gtype == GEOM_ELLIPSOID:
        vol = 4.0 / 3.0 * math.pi * size[0] * size[1] * size[2]
    else:  # plane
        return 0.0, np.zeros(3)
    mass = float(mass_attr) if mass_attr is not None else density * vol
    if gtype == GEOM_SPHERE:
        I = np.full(3, 2.0 * mass * size[0] ** 2 / 5.0)
    elif gtype == GEOM_CAPSULE:
        r, h = size[0], 2 * size[1]
        sm = mass * 4 * r / (4 * r + 3 * h)
        cm = mass - sm
        I = np.zeros(3)
        I[0] = I[1] = cm * (3 * r * r + h * h) / 12.0
        I[2] = cm * r * r / 2.0
        si = 2 * sm * r * r / 5.0
        I[0] += si + sm * h * (3 * r + 2 * h) / 8.0
        I[1] += si + sm * h * (3 * r + 2 * h) / 8.0
        I[2] += si
    elif gtype == GEOM_CYLINDER:
        r, h = size[0], 2 * size[1]
        I = np.array([mass * (3 * r * r + h * h) / 12.0, mass * (3 * r * r + h * h) / 12.0, mass * r * r / 2.0])
    elif gtype == GEOM_BOX:
        I = mass / 3.0 * np.array([size[1] ** 2 + size[2] ** 2, size[0] ** 2 + size[2] ** 2,
                                   size[0] ** 2 + size[1] ** 2])
    else:  # ellipsoid
        I = mass / 5.0 * np.array([size[1] ** 2 + size[2] ** 2, size[0] ** 2 + size[2] ** 2,
                                   size[0] ** 2 + size[1] ** 2])
    return mass, I


def _geom_rbound(gtype, size):
    if gtype == GEOM_SPHERE:
        return size[0]
    if gtype == GEOM_CAPSULE:
        return size[0] + size[1]
    if gtype == GEOM_CYLINDER:
        return math.sqrt(size[0] ** 2 + size[1] ** 2)
    if gtype in (GEOM_BOX, GEOM_ELLIPSOID):
        return float(np.linalg.norm(size)) if gtype == GEOM_BOX else float(max(size))
    return 0.0  # plane: unbounded, handled by the plane test


# ----------------------------------------------------------------------------- compiler
class _Body:
    def __init__(self):
        self.name = ""; self.parent = 0
        self.pos = np.zeros(3); self.quat = np.array([1.0, 0, 0, 0])
        self.joints = []; self.geoms = []; self.sites = []
        self.inertial = None


def compile_mjcf(xml_text: str, name: str = "") -> ModelTables:
    """Compile one MJCF string into :class:`ModelTables`."""
    root = ET.fromstring(xml_text)
    if root.tag != "mujoco":
        raise MjcfCompileError("root element must be <mujoco>")
    comp = root.find("compiler")
    angle_deg = True
    inertiafromgeom = "auto"
    eulerseq = "xyz"
    autolimits = True  # MuJoCo >= 3.0 default (App. C.7)
    if comp is not None:
        angle_deg = comp.get("angle", "degree") == "degree"
        inertiafromgeom = comp.get("inertiafromgeom", "auto")
        eulerseq = comp.get("eulerseq", "xyz")
        autolimits = _bool(comp.get("autolimits"), True)
    asc = math.pi / 180.0 if angle_deg else 1.0

    # ---- option
    opt = {"timestep": 0.002, "gravity": np.array([0, 0, -9.81]), "iterations": 100, "tolerance": 1e-8,
           "solver": SOLVER_NEWTON, "integrator": INT_EULER, "impratio": 1.0, "cone": 0, "ls_iterations": 50,
           "ls_tolerance": 0.01}
    for o in root.findall("option"):
        if "timestep" in o.attrib: opt["timestep"] = float(o.get("timestep"))
        if "gravity" in o.attrib: opt["gravity"] = _floats(o.get("gravity"))
        if "iterations" in o.attrib: opt["iterations"] = int(o.get("iterations"))
        if "tolerance" in o.attrib: opt["tolerance"] = float(o.get("tolerance"))
        if "ls_iterations" in o.attrib: opt["ls_iterations"] = int(o.get("ls_iterations"))
        if "ls_tolerance" in o.attrib: opt["ls_tolerance"] = float(o.get("ls_tolerance"))
        if "impratio" in o.attrib: opt["impratio"] = float(o.get("impratio"))
        if "solver" in o.attrib:
            opt["solver"] = {"PGS": SOLVER_PGS, "CG": SOLVER_CG, "Newton": SOLVER_NEWTON}[o.get("solver")]
        if "integrator" in o.attrib:
            it = o.get("integrator")
            if it not in ("Euler", "RK4"):
                raise MjcfCompileError(f"integrator {it} not in the supported subset")
            opt["integrator"] = INT_EULER if it == "Euler" else INT_RK4
        if o.get("cone", "pyramidal") != "pyramidal":
            raise MjcfCompileError("only pyramidal cones are in the supported subset")

    dfl = _Defaults(root)

    # ---- body tree, depth-first pre-order (App. C.3)
    bodies: List[_Body] = []
    world = _Body(); world.name = "world"; world.parent = 0
    bodies.append(world)

    def parse_body(elem, bidx):
        b = bodies[bidx]
        for ch in elem:
            if ch.tag in ("joint", "freejoint"):
                b.joints.append(ch)
            elif ch.tag == "geom":
                b.geoms.append(ch)
            elif ch.tag == "site":
                b.sites.append(ch)
            elif ch.tag == "inertial":
                b.inertial = ch
        for ch in elem:
            if ch.tag == "body":
                nb = _Body()
                nb.name = ch.get("name", f"body{len(bodies)}")
                nb.parent = bidx
                nb.pos = _floats(ch.get("pos"), 3, [0, 0, 0])
                nb.quat = _orientation(ch, asc, eulerseq)
                bodies.append(nb)
                parse_body(ch, len(bodies) - 1)

    wb = root.find("worldbody")
    if wb is None:
        raise MjcfCompileError("missing <worldbody>")
    parse_body(wb, 0)
    nbody = len(bodies)

    # ---- joints
    jnt = dict(type=[], qposadr=[], dofadr=[], bodyid=[], pos=[], axis=[], limited=[], range=[], stiffness=[],
               damping=[], armature=[], margin=[], name=[])
    body_jntadr = np.full(nbody, -1, np.int32); body_jntnum = np.zeros(nbody, np.int32)
    body_dofadr = np.full(nbody, -1, np.int32); body_dofnum = np.zeros(nbody, np.int32)
    nq = nv = 0
    qpos0 = []
    qpos_spring = []
    for bi, b in enumerate(bodies):
        for je in b.joints:
            if je.tag == "freejoint":
                jt = JNT_FREE
            else:
                jt = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[
                    dfl.get("joint", je, "type", "hinge")]
            if jt == JNT_BALL:
                raise MjcfCompileError("ball joints are not in the supported subset")
            if body_jntnum[bi] == 0:
                body_jntadr[bi] = len(jnt["type"]); body_dofadr[bi] = nv
            body_jntnum[bi] += 1
            jnt["type"].append(jt); jnt["qposadr"].append(nq); jnt["dofadr"].append(nv); jnt["bodyid"].append(bi)
            jnt["name"].append(je.get("name", f"joint{len(jnt['name'])}"))
            tag = "joint"
            jnt["pos"].append(_floats(dfl.get(tag, je, "pos"), 3, [0, 0, 0]) if je.tag == "joint" else np.zeros(3))
            ax = _floats(dfl.get(tag, je, "axis"), 3, [0, 0, 1]) if je.tag == "joint" else np.array([0, 0, 1.0])
            n = np.linalg.norm(ax)
            jnt["axis"].append(ax / n if n > 0 else np.array([0, 0, 1.0]))
            if jt == JNT_FREE:
                # free joints ignore limits, armature/damping defaults apply only when written on the element
                jnt["limited"].append(0); jnt["range"].append(np.zeros(2))
                jnt["stiffness"].append(0.0)
                jnt["damping"].append(float(je.get("damping", 0.0)) if je.tag == "joint" else 0.0)
                jnt["armature"].append(float(je.get("armature", 0.0)) if je.tag == "joint" else 0.0)
                jnt["margin"].append(0.0)
                if je.tag == "joint":
                    # MuJoCo applies joint defaults to free joints as well (armature/damping)
                    jnt["damping"][-1] = float(dfl.get(tag, je, "damping", 0.0))
                    jnt["armature"][-1] = float(dfl.get(tag, je, "armature", 0.0))
                body_dofnum[bi] += 6
                qpos0.extend(list(b.pos) + list(b.quat))
                qpos_spring.extend(list(b.pos) + list(b.quat))
                nq += 7; nv += 6
            else:
                rng = _floats(dfl.get(tag, je, "range"), 2, [0, 0])
                if jt == JNT_HINGE:
                    rng = rng * asc
                lim_s = dfl.get(tag, je, "limited", "auto")
                if lim_s == "auto":
                    limited = 1 if (autolimits and dfl.get(tag, je, "range") is not None) else 0
                else:
                    limited = 1 if lim_s == "true" else 0
                jnt["limited"].append(limited); jnt["range"].append(rng)
                jnt["stiffness"].append(float(dfl.get(tag, je, "stiffness", 0.0)))
                jnt["damping"].append(float(dfl.get(tag, je, "damping", 0.0)))
                jnt["armature"].append(float(dfl.get(tag, je, "armature", 0.0)))
                jnt["margin"].append(float(dfl.get(tag, je, "margin", 0.0)))
                ref = float(dfl.get(tag, je, "ref", 0.0)) * (asc if jt == JNT_HINGE else 1.0)
                body_dofnum[bi] += 1
                qpos0.append(ref)
                qpos_spring.append(float(dfl.get(tag, je, "springref", 0.0)) * (asc if jt == JNT_HINGE else 1.0))
                nq += 1; nv += 1
    njnt = len(jnt["type"])
    for bi in range(nbody):
        if body_jntnum[bi] > 1 and any(jnt["type"][body_jntadr[bi] + k] == JNT_FREE for k in range(body_jntnum[bi])):
            raise MjcfCompileError("free joint must be the only joint of its body")

    # ---- dofs
    dof_bodyid = np.zeros(nv, np.int32); dof_jntid = np.zeros(nv, np.int32); dof_parentid = np.full(nv, -1, np.int32)
    dof_armature = np.zeros(nv); dof_damping = np.zeros(nv)
    last_dof_of_body = np.full(nbody, -1, np.int32)
    for j in range(njnt):
        bi = jnt["bodyid"][j]; da = jnt["dofadr"][j]
        nd = 6 if jnt["type"][j] == JNT_FREE else 1
        for k in range(nd):
            d = da + k
            dof_bodyid[d] = bi; dof_jntid[d] = j
            dof_armature[d] = jnt["armature"][j]; dof_damping[d] = jnt["damping"][j]
            if last_dof_of_body[bi] >= 0:
                dof_parentid[d] = last_dof_of_body[bi]
            else:
                p = bodies[bi].parent
                while p > 0 and last_dof_of_body[p] < 0:
                    p = bodies[p].parent
                dof_parentid[d] = last_dof_of_body[p] if p > 0 else -1
            last_dof_of_body[bi] = d

    # ---- body bookkeeping: weld ids, root ids, trees
    body_parentid = np.array([b.parent for b in bodies], np.int32)
    body_weldid = np.zeros(nbody, np.int32); body_rootid = np.zeros(nbody, np.int32)
    body_depth = np.zeros(nbody, np.int32)
    for bi in range(1, nbody):
        p = body_parentid[bi]
        body_weldid[bi] = bi if body_jntnum[bi] > 0 else body_weldid[p]
        body_rootid[bi] = bi if p == 0 else body_rootid[p]
        body_depth[bi] = body_depth[p] + 1
    # kinematic trees = connected sets of dofs; body_treeid = -1 for static bodies
    dof_treeid = np.full(nv, -1, np.int32)
    ntree = 0
    for d in range(nv):
        if dof_parentid[d] < 0:
            dof_treeid[d] = ntree; ntree += 1
        else:
            dof_treeid[d] = dof_treeid[dof_parentid[d]]
    body_treeid = np.full(nbody, -1, np.int32)
    for bi in range(1, nbody):
        w = body_weldid[bi]
        if w > 0:
            # last dof of the weld body identifies the tree
            body_treeid[bi] = dof_treeid[last_dof_of_body[w]]

    # ---- geoms: world geoms first automatically (world body is id 0, pre-order)
    g = dict(type=[], bodyid=[], size=[], pos=[], quat=[], contype=[], conaffinity=[], condim=[], friction=[],
             margin=[], gap=[], solref=[], solimp=[], solmix=[], priority=[], rbound=[], mass=[], inertia=[],
             name=[])
    body_geomadr = np.full(nbody, -1, np.int32); body_geomnum = np.zeros(nbody, np.int32)
    for bi, b in enumerate(bodies):
        for ge in b.geoms:
            tname = dfl.get("geom", ge, "type", "sphere")
            if tname not in GEOM_NAMES:
                raise MjcfCompileError(f"geom type {tname} not in the supported subset")
            gt = GEOM_NAMES[tname]
            size = _floats(dfl.get("geom", ge, "size"), 3, [0, 0, 0])
            if size is None:
                size = np.zeros(3)
            pos = _floats(ge.get("pos"), 3, [0, 0, 0])
            quat = _orientation(ge, asc, eulerseq)
            if "fromto" in ge.attrib:
                ft = _floats(ge.get("fromto"))
                vec = ft[0:3] - ft[3:6]          # points from `to` towards `from`
                half = 0.5 * np.linalg.norm(vec)
                pos = 0.5 * (ft[0:3] + ft[3:6])
                quat = quat_z_to_vec(vec)
                if gt in (GEOM_CAPSULE, GEOM_CYLINDER):
                    size = np.array([size[0], half, 0.0])
                else:
                    size = np.array([size[0], size[1], half])
            if body_geomnum[bi] == 0:
                body_geomadr[bi] = len(g["type"])
            body_geomnum[bi] += 1
            g["type"].append(gt); g["bodyid"].append(bi); g["size"].append(size); g["pos"].append(pos)
            g["quat"].append(quat)
            g["contype"].append(int(dfl.get("geom", ge, "contype", 1)))
            g["conaffinity"].append(int(dfl.get("geom", ge, "conaffinity", 1)))
            g["condim"].append(int(dfl.get("geom", ge, "condim", 3)))
            g["friction"].append(_floats(dfl.get("geom", ge, "friction"), 3, [1, 0.005, 0.0001]))
            g["margin"].append(float(dfl.get("geom", ge, "margin", 0.0)))
            g["gap"].append(float(dfl.get("geom", ge, "gap", 0.0)))
            g["solref"].append(_floats(dfl.get("geom", ge, "solref"), 2, [0.02, 1.0]))
            g["solimp"].append(_floats(dfl.get("geom", ge, "solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]))
            g["solmix"].append(float(dfl.get("geom", ge, "solmix", 1.0)))
            g["priority"].append(int(dfl.get("geom", ge, "priority", 0)))
            g["rbound"].append(_geom_rbound(gt, size))
            density = float(dfl.get("geom", ge, "density", 1000.0))
            m_, I_ = _geom_mass_inertia(gt, size, density, dfl.get("geom", ge, "mass"))
            g["mass"].append(m_); g["inertia"].append(I_)
            g["name"].append(ge.get("name", ""))
    ngeom = len(g["type"])

    # ---- sites
    s_body, s_pos, s_quat, s_name = [], [], [], []
    for bi, b in enumerate(bodies):
        for se in b.sites:
            s_body.append(bi); s_pos.append(_floats(se.get("pos"), 3, [0, 0, 0]))
            s_quat.append(_orientation(se, asc, eulerseq)); s_name.append(se.get("name", ""))
    nsite = len(s_body)

    # ---- body inertial properties
    body_mass = np.zeros(nbody); body_ipos = np.zeros((nbody, 3)); body_iquat = np.zeros((nbody, 4))
    body_iquat[:, 0] = 1.0
    body_inertia = np.zeros((nbody, 3))
    for bi, b in enumerate(bodies):
        if bi == 0:
            continue
        use_geoms = inertiafromgeom == "true" or (inertiafromgeom == "auto" and b.inertial is None)
        if use_geoms and body_geomnum[bi] > 0:
            ga = body_geomadr[bi]; gn = body_geomnum[bi]
            ms = np.array([g["mass"][ga + k] for k in range(gn)])
            mt = ms.sum()
            if mt <= 0:
                continue
            com = sum(ms[k] * g["pos"][ga + k] for k in range(gn)) / mt
            I = np.zeros((3, 3))
            for k in range(gn):
                R = quat_to_mat(g["quat"][ga + k])
                Ik = R @ np.diag(g["inertia"][ga + k]) @ R.T
                dvec = g["pos"][ga + k] - com
                Ik += ms[k] * (np.dot(dvec, dvec) * np.eye(3) - np.outer(dvec, dvec))
                I += Ik
            body_mass[bi] = mt; body_ipos[bi] = com
            if gn == 1:
                body_iquat[bi] = g["quat"][ga]; body_inertia[bi] = g["inertia"][ga]
            else:
                w, V = np.linalg.eigh(I)
                order = np.argsort(-w)
                w = w[order]; V = V[:, order]
                if np.linalg.det(V) < 0:
                    V[:, 2] = -V[:, 2]
                body_iquat[bi] = mat_to_quat(V); body_inertia[bi] = w
        elif b.inertial is not None:
            ie = b.inertial
            body_mass[bi] = float(ie.get("mass"))
            body_ipos[bi] = _floats(ie.get("pos"), 3, [0, 0, 0])
            body_iquat[bi] = _orientation(ie, asc, eulerseq)
            if "diaginertia" in ie.attrib:
                body_inertia[bi] = _floats(ie.get("diaginertia"))
            elif "fullinertia" in ie.attrib:
                f = _floats(ie.get("fullinertia"))
                I = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
                w, V = np.linalg.eigh(I)
                order = np.argsort(-w); w = w[order]; V = V[:, order]
                if np.linalg.det(V) < 0:
                    V[:, 2] = -V[:, 2]
                body_iquat[bi] = quat_mul(body_iquat[bi], mat_to_quat(V)); body_inertia[bi] = w
    for bi in range(1, nbody):
        if body_weldid[bi] == bi and (body_mass[bi] < MJMINVAL or body_inertia[bi].min() < MJMINVAL):
            # MuJoCo: "mass and inertia of moving bodies must be larger than mjMINVAL"
            has_child_mass = any(body_mass[c] > 0 for c in range(bi + 1, nbody)
                                 if _is_descendant(body_parentid, c, bi))
            if not has_child_mass:
                raise MjcfCompileError(f"moving body {bodies[bi].name} has no mass/inertia")
    body_subtreemass = body_mass.copy()
    for bi in range(nbody - 1, 0, -1):
        body_subtreemass[body_parentid[bi]] += body_subtreemass[bi]

    # ---- actuators (joint transmission only)
    act = dict(dofid=[], gear=[], ctrllimited=[], ctrlrange=[], forcelimited=[], forcerange=[], gainprm=[],
               biasprm=[], name=[])
    an = root.find("actuator")
    if an is not None:
        for ae in an:
            if ae.tag not in ("motor", "position", "general", "velocity"):
                raise MjcfCompileError(f"actuator {ae.tag} not in the supported subset")
            jn = ae.get("joint")
            if jn is None or jn not in jnt["name"]:
                raise MjcfCompileError(f"actuator {ae.get('name')} needs a known joint")
            j = jnt["name"].index(jn)
            if jnt["type"][j] == JNT_FREE:
                raise MjcfCompileError("joint transmission on a free joint is not supported")
            act["dofid"].append(jnt["dofadr"][j])
            gear = _floats(dfl.get(ae.tag, ae, "gear"), 1, [1.0])
            act["gear"].append(float(gear[0]))
            cr = dfl.get(ae.tag, ae, "ctrlrange")
            cl = dfl.get(ae.tag, ae, "ctrllimited", "auto")
            act["ctrlrange"].append(_floats(cr, 2, [0, 0]))
            act["ctrllimited"].append(int(cl == "true" or (cl == "auto" and autolimits and cr is not None)))
            fr = dfl.get(ae.tag, ae, "forcerange")
            fl = dfl.get(ae.tag, ae, "forcelimited", "auto")
            act["forcerange"].append(_floats(fr, 2, [0, 0]))
            act["forcelimited"].append(int(fl == "true" or (fl == "auto" and autolimits and fr is not None)))
            if ae.tag == "motor":
                act["gainprm"].append(1.0); act["biasprm"].append(np.zeros(3))
            elif ae.tag == "position":
                kp = float(dfl.get(ae.tag, ae, "kp", 1.0))
                kv = float(dfl.get(ae.tag, ae, "kv", 0.0))
                act["gainprm"].append(kp); act["biasprm"].append(np.array([0.0, -kp, -kv]))
            elif ae.tag == "velocity":
                kv = float(dfl.get(ae.tag, ae, "kv", 1.0))
                act["gainprm"].append(kv); act["biasprm"].append(np.array([0.0, 0.0, -kv]))
            else:
                gp = _floats(dfl.get(ae.tag, ae, "gainprm"), 1, [1.0])
                bp = _floats(dfl.get(ae.tag, ae, "biasprm"), 3, [0, 0, 0])
                act["gainprm"].append(float(gp[0])); act["biasprm"].append(bp)
            act["name"].append(ae.get("name", ""))
    nu = len(act["dofid"])

    # ---- explicit contact pairs / excludes
    explicit = {}
    excludes = set()
    cn = root.find("contact")
    if cn is not None:
        for pe in cn:
            if pe.tag == "pair":
                g1 = g["name"].index(pe.get("geom1")); g2 = g["name"].index(pe.get("geom2"))
                fr = _floats(pe.get("friction"), 5, [1, 1, 0.005, 0.0001, 0.0001])
                explicit[(min(g1, g2), max(g1, g2))] = dict(
                    g1=g1, g2=g2, condim=int(pe.get("condim", 3)), friction=fr,
                    margin=float(pe.get("margin", 0.0)), gap=float(pe.get("gap", 0.0)),
                    solref=_floats(pe.get("solref"), 2, [0.02, 1.0]),
                    solimp=_floats(pe.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]))
            elif pe.tag == "exclude":
                names = [b.name for b in bodies]
                b1 = names.index(pe.get("body1")); b2 = names.index(pe.get("body2"))
                excludes.add((min(b1, b2), max(b1, b2)))

    # ---- candidate geom pairs, in MuJoCo's contact order (App. B.4): body pairs ascending, then geoms
    pairs = _candidate_pairs(nbody, body_parentid, body_weldid, body_geomadr, body_geomnum, g, explicit, excludes)

    arrays: Dict[str, np.ndarray] = {}
    A = arrays
    def f64(x, shape=None):
        a = np.array(x, np.float64)
        if shape:
            return np.ascontiguousarray(a.reshape(shape))
        return a if a.ndim == 0 else np.ascontiguousarray(a)
    i32 = lambda x: np.array(x, np.int32) if np.ndim(x) == 0 else np.ascontiguousarray(np.array(x, np.int32))
    A["nq"] = i32(nq); A["nv"] = i32(nv); A["nu"] = i32(nu); A["nbody"] = i32(nbody); A["njnt"] = i32(njnt)
    A["ngeom"] = i32(ngeom); A["nsite"] = i32(nsite); A["ntree"] = i32(ntree); A["npair"] = i32(len(pairs["g1"]))
    A["timestep"] = f64(opt["timestep"]); A["gravity"] = f64(opt["gravity"])
    A["iterations"] = i32(opt["iterations"]); A["tolerance"] = f64(opt["tolerance"])
    A["ls_iterations"] = i32(opt["ls_iterations"]); A["ls_tolerance"] = f64(opt["ls_tolerance"])
    A["solver"] = i32(opt["solver"]); A["integrator"] = i32(opt["integrator"]); A["impratio"] = f64(opt["impratio"])
    A["body_parentid"] = body_parentid; A["body_weldid"] = body_weldid; A["body_rootid"] = body_rootid
    A["body_depth"] = body_depth; A["body_treeid"] = body_treeid
    A["body_jntadr"] = body_jntadr; A["body_jntnum"] = body_jntnum
    A["body_dofadr"] = body_dofadr; A["body_dofnum"] = body_dofnum
    A["body_geomadr"] = body_geomadr; A["body_geomnum"] = body_geomnum
    A["body_pos"] = f64([b.pos for b in bodies], (nbody, 3)); A["body_quat"] = f64([b.quat for b in bodies], (nbody, 4))
    A["body_ipos"] = body_ipos; A["body_iquat"] = body_iquat; A["body_mass"] = body_mass
    A["body_subtreemass"] = body_subtreemass; A["body_inertia"] = body_inertia
    A["jnt_type"] = i32(jnt["type"]); A["jnt_qposadr"] = i32(jnt["qposadr"]); A["jnt_dofadr"] = i32(jnt["dofadr"])
    A["jnt_bodyid"] = i32(jnt["bodyid"]); A["jnt_pos"] = f64(jnt["pos"], (njnt, 3)); A["jnt_axis"] = f64(jnt["axis"], (njnt, 3))
    A["jnt_limited"] = i32(jnt["limited"]); A["jnt_range"] = f64(jnt["range"], (njnt, 2))
    A["jnt_stiffness"] = f64(jnt["stiffness"]); A["jnt_margin"] = f64(jnt["margin"])
    A["jnt_solref"] = f64([[0.02, 1.0]] * njnt, (njnt, 2))
    A["jnt_solimp"] = f64([[0.9, 0.95, 0.001, 0.5, 2.0]] * njnt, (njnt, 5))
    A["qpos0"] = f64(qpos0); A["qpos_spring"] = f64(qpos_spring)
    A["dof_bodyid"] = dof_bodyid; A["dof_jntid"] = dof_jntid; A["dof_parentid"] = dof_parentid
    A["dof_treeid"] = dof_treeid; A["dof_armature"] = dof_armature; A["dof_damping"] = dof_damping
    A["geom_type"] = i32(g["type"]); A["geom_bodyid"] = i32(g["bodyid"]); A["geom_size"] = f64(g["size"], (ngeom, 3))
    A["geom_pos"] = f64(g["pos"], (ngeom, 3)); A["geom_quat"] = f64(g["quat"], (ngeom, 4))
    A["geom_contype"] = i32(g["contype"]); A["geom_conaffinity"] = i32(g["conaffinity"])
    A["geom_condim"] = i32(g["condim"]); A["geom_friction"] = f64(g["friction"], (ngeom, 3))
    A["geom_margin"] = f64(g["margin"]); A["geom_gap"] = f64(g["gap"]); A["geom_rbound"] = f64(g["rbound"])
    A["site_bodyid"] = i32(s_body); A["site_pos"] = f64(s_pos, (nsite, 3)); A["site_quat"] = f64(s_quat, (nsite, 4))
    A["act_dofid"] = i32(act["dofid"]); A["act_gear"] = f64(act["gear"])
    A["act_ctrllimited"] = i32(act["ctrllimited"]); A["act_ctrlrange"] = f64(act["ctrlrange"], (nu, 2))
    A["act_forcelimited"] = i32(act["forcelimited"]); A["act_forcerange"] = f64(act["forcerange"], (nu, 2))
    A["act_gainprm"] = f64(act["gainprm"]); A["act_biasprm"] = f64(act["biasprm"], (nu, 3))
    for k in ("g1", "g2", "condim"):
        A["pair_" + k] = i32(pairs[k])
    npair = len(pairs["g1"])
    A["pair_friction"] = f64(pairs["friction"], (npair, 5)); A["pair_margin"] = f64(pairs["margin"])
    A["pair_gap"] = f64(pairs["gap"]); A["pair_solref"] = f64(pairs["solref"], (npair, 2))
    A["pair_solimp"] = f64(pairs["solimp"], (npair, 5))

    m = ModelTables(name=name, arrays=arrays,
                    names=dict(body=[b.name for b in bodies], joint=jnt["name"], geom=g["name"],
                               site=s_name, actuator=act["name"]))
    _set_const(m)
    return m


def _is_descendant(parentid, c, anc):
    while c > 0:
        c = parentid[c]
        if c == anc:
            return True
    return False


def _candidate_pairs(nbody, body_parentid, body_weldid, body_geomadr, body_geomnum, g, explicit, excludes):
    """Static collision filter: weld ids, parent filter, contype/conaffinity bitmask."""
    out = dict(g1=[], g2=[], condim=[], friction=[], margin=[], gap=[], solref=[], solimp=[])
    weldparent = np.array([body_weldid[body_parentid[body_weldid[b]]] for b in range(nbody)])

    def add(g1, g2, prm):
        # geoms are ordered by type: the lower type id is geom1 (plane first)
        if g["type"][g1] > g["type"][g2]:
            g1, g2 = g2, g1
        out["g1"].append(g1); out["g2"].append(g2)
        for k in ("condim", "friction", "margin", "gap", "solref", "solimp"):
            out[k].append(prm[k])

    used_explicit = set()
    for b1 in range(nbody):
        for b2 in range(b1 + 1, nbody):
            if body_geomnum[b1] == 0 or body_geomnum[b2] == 0:
                continue
            w1, w2 = body_weldid[b1], body_weldid[b2]
            dyn_ok = True
            if w1 == w2:
                dyn_ok = False
            elif w1 != 0 and w2 != 0 and (w1 == weldparent[b2] or w2 == weldparent[b1]):
                dyn_ok = False
            if (b1, b2) in excludes:
                dyn_ok = False
            for k1 in range(body_geomnum[b1]):
                for k2 in range(body_geomnum[b2]):
                    g1 = body_geomadr[b1] + k1; g2 = body_geomadr[b2] + k2
                    key = (min(g1, g2), max(g1, g2))
                    if key in explicit:
                        e = explicit[key]
                        add(e["g1"], e["g2"], e); used_explicit.add(key)
                        continue
                    if not dyn_ok:
                        continue
                    if not ((g["contype"][g1] & g["conaffinity"][g2]) or (g["contype"][g2] & g["conaffinity"][g1])):
                        continue
                    if g["type"][g1] == GEOM_PLANE and g["type"][g2] == GEOM_PLANE:
                        continue
                    f1, f2 = g["friction"][g1], g["friction"][g2]
                    if g["priority"][g1] != g["priority"][g2]:
                        hi = g1 if g["priority"][g1] > g["priority"][g2] else g2
                        fr = g["friction"][hi]; sref = g["solref"][hi]; simp = g["solimp"][hi]
                        cd = g["condim"][hi]
                    else:
                        fr = np.maximum(f1, f2)
                        s1, s2 = g["solmix"][g1], g["solmix"][g2]
                        mix = s1 / (s1 + s2) if (s1 + s2) > MJMINVAL else 0.5
                        r1, r2 = g["solref"][g1], g["solref"][g2]
                        sref = mix * r1 + (1 - mix) * r2 if (r1[0] > 0 and r2[0] > 0) else np.minimum(r1, r2)
                        simp = mix * g["solimp"][g1] + (1 - mix) * g["solimp"][g2]
                        cd = max(g["condim"][g1], g["condim"][g2])
                    prm = dict(condim=cd, friction=np.array([fr[0], fr[0], fr[1], fr[2], fr[2]]),
                               margin=max(g["margin"][g1], g["margin"][g2]), gap=max(g["gap"][g1], g["gap"][g2]),
                               solref=sref, solimp=simp)
                    add(g1, g2, prm)
    # explicit pairs between geoms of the same body pair are consumed above; same-body pairs append last
    for key, e in explicit.items():
        if key not in used_explicit:
            add(e["g1"], e["g2"], e)
    return out


# ----------------------------------------------------------------------------- mj_setConst restatement
def _kinematics0(m: ModelTables):
    """Forward kinematics at qpos0 (numpy); returns world frames, com frames, cdof. Used only for constants."""
    A = m.arrays
    nbody, nv = int(A["nbody"]), int(A["nv"])
    xpos = np.zeros((nbody, 3)); xquat = np.zeros((nbody, 4)); xquat[0, 0] = 1
    xanchor = np.zeros((int(A["njnt"]), 3)); xaxis = np.zeros((int(A["njnt"]), 3))
    q0 = A["qpos0"]
    for b in range(1, nbody):
        p = A["body_parentid"][b]
        ja, jn = A["body_jntadr"][b], A["body_jntnum"][b]
        if jn == 1 and A["jnt_type"][ja] == JNT_FREE:
            qa = A["jnt_qposadr"][ja]
            xpos[b] = q0[qa:qa + 3]; xquat[b] = quat_normalize(q0[qa + 3:qa + 7])
            xanchor[ja] = xpos[b]; xaxis[ja] = quat_to_mat(xquat[b])[:, 2]
        else:
            R = quat_to_mat(xquat[p])
            xpos[b] = xpos[p] + R @ A["body_pos"][b]; xquat[b] = quat_mul(xquat[p], A["body_quat"][b])
            for k in range(jn):
                j = ja + k
                Rb = quat_to_mat(xquat[b])
                xanchor[j] = Rb @ A["jnt_pos"][j] + xpos[b]; xaxis[j] = Rb @ A["jnt_axis"][j]
                # at qpos0 the joint displacement is zero: no motion applied
        xquat[b] = quat_normalize(xquat[b])
    xmat = np.stack([quat_to_mat(q) for q in xquat])
    xipos = xpos + np.einsum("bij,bj->bi", xmat, A["body_ipos"])
    ximat = np.stack([quat_to_mat(quat_mul(xquat[b], A["body_iquat"][b])) for b in range(nbody)])
    return xpos, xquat, xmat, xipos, ximat, xanchor, xaxis


def _dense_mass_matrix(m: ModelTables, xmat, xipos, ximat, xanchor, xaxis):
    """Dense joint-space inertia at qpos0 via world-frame Jacobians (independent of the CRB pass the
    engine uses; serves as a cross-check in tests)."""
    A = m.arrays
    nbody, nv = int(A["nbody"]), int(A["nv"])
    M = np.zeros((nv, nv))
    Jp_all = np.zeros((nbody, 3, nv)); Jr_all = np.zeros((nbody, 3, nv))
    for b in range(1, nbody):
        Jp = Jp_all[b]; Jr = Jr_all[b]
        c = b
        while c > 0:
            for k in range(A["body_jntnum"][c]):
                j = A["body_jntadr"][c] + k
                d = A["jnt_dofadr"][j]; t = A["jnt_type"][j]
                if t == JNT_FREE:
                    Jp[:, d:d + 3] = np.eye(3)
                    R = xmat[c]
                    for a in range(3):
                        ax = R[:, a]
                        Jr[:, d + 3 + a] = ax
                        Jp[:, d + 3 + a] = np.cross(ax, xipos[b] - xanchor[j])
                elif t == JNT_SLIDE:
                    Jp[:, d] = xaxis[j]
                else:
                    Jr[:, d] = xaxis[j]; Jp[:, d] = np.cross(xaxis[j], xipos[b] - xanchor[j])
            c = A["body_parentid"][c]
        Iw = ximat[b] @ np.diag(A["body_inertia"][b]) @ ximat[b].T
        M += A["body_mass"][b] * Jp.T @ Jp + Jr.T @ Iw @ Jr
    M += np.diag(A["dof_armature"])
    return M, Jp_all, Jr_all


def _set_const(m: ModelTables) -> None:
    """``mj_setConst`` subset: dof_invweight0, body_invweight0, meaninertia, dof_Madr (App. C.6)."""
    A = m.arrays
    nbody, nv = int(A["nbody"]), int(A["nv"])
    xpos, xquat, xmat, xipos, ximat, xanchor, xaxis = _kinematics0(m)
    M, Jp, Jr = _dense_mass_matrix(m, xmat, xipos, ximat, xanchor, xaxis)
    dof_invweight0 = np.zeros(nv); body_invweight0 = np.zeros((nbody, 2))
    if nv > 0:
        Minv = np.linalg.inv(M)
        dof_invweight0 = np.diag(Minv).copy()
        for j in range(int(A["njnt"])):
            if A["jnt_type"][j] == JNT_FREE:
                d = A["jnt_dofadr"][j]
                dof_invweight0[d:d + 3] = dof_invweight0[d:d + 3].mean()
                dof_invweight0[d + 3:d + 6] = dof_invweight0[d + 3:d + 6].mean()
        for b in range(1, nbody):
            if A["body_weldid"][b] == 0:
                continue
            At = Jp[b] @ Minv @ Jp[b].T; Ar = Jr[b] @ Minv @ Jr[b].T
            body_invweight0[b, 0] = max(MJMINVAL, np.trace(At) / 3.0)
            body_invweight0[b, 1] = max(MJMINVAL, np.trace(Ar) / 3.0)
        A["meaninertia"] = np.array(float(np.mean(np.diag(M))))
    else:
        A["meaninertia"] = np.array(1.0)
    A["dof_invweight0"] = dof_invweight0; A["body_invweight0"] = body_invweight0
    # sparse-M addressing: entry k of row i is (i, k-th ancestor of i), row i starts at dof_Madr[i]
    Madr = np.zeros(nv, np.int32); n = 0
    for i in range(nv):
        Madr[i] = n
        j = i
        while j >= 0:
            n += 1; j = A["dof_parentid"][j]
    A["dof_Madr"] = Madr; A["nM"] = np.array(n, np.int32)
    A["M0"] = M  # dense inertia at qpos0, kept for tests
