"""Per-task MJCF composition: assets directory -> one MJCF string per task.

Each function restates what the reference env's model-assembly step produces
(SURVEY.md §8 row a12), so the compiled tables describe the model the reference
actually simulates, including the filtering side effects of its merging rules
(SURVEY F7).  The asset files themselves are *read* from the reference package
(``assets_root``); they are not part of this repository.  Compiled tables are
committed under ``tables/`` by ``tools/compile_tables.py`` so the engine runs
where the reference package is not installed.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import xml.etree.ElementTree as ET


def _read(path: str) -> ET.Element:
    with open(path, "r") as f:
        return ET.fromstring(f.read())


def quadruped_parkour_mjcf(assets_root: str) -> str:
    """Composite model of ``QuadrupedParkourEnv`` (quadruped_parkour_env/parkour_env.py:105-179).

    Base document = ``quadruped.xml`` (its compiler/option/size/default/sensor survive).  Appended to its
    worldbody: every non-light child of the course worldbody; from the terrain worldbody every child except
    lights and the geom named ``floor``.  Actuators of both extra files are appended in file order; assets
    are merged by first-seen name.  The extra files' <default> blocks are *not* carried over, so appended
    geoms/joints inherit the base defaults (contype 1 / conaffinity 0: they never collide with the robot).
    """
    adir = os.path.join(assets_root, "quadruped_parkour_env", "assets")
    base = _read(os.path.join(adir, "quadruped.xml"))
    extras = [_read(os.path.join(adir, "parkour_course.xml")), _read(os.path.join(adir, "terrain_variations.xml"))]
    wb = base.find("worldbody")
    for k, doc in enumerate(extras):
        src = doc.find("worldbody")
        if src is None:
            continue
        for node in list(src):
            if node.tag == "light":
                continue
            if k == 1 and node.tag == "geom" and node.get("name") == "floor":
                continue
            wb.append(node)
    for section in ("actuator", "asset"):
        dst = base.find(section)
        if dst is None:
            dst = ET.SubElement(base, section)
        for doc in extras:
            src = doc.find(section)
            if src is None:
                continue
            for node in list(src):
                if section == "asset" and node.get("name") in {n.get("name") for n in dst}:
                    continue
                dst.append(node)
    return ET.tostring(base, encoding="unicode")


# ---------------------------------------------------------------------------------------------- inline generators
# Five reference envs build their MJCF in Python (dancing ``dancing_env.py:156-678``, soccer ``soccer_env.py:120-220``,
# martial arts ``martial_arts_env.py:150-381``, construction ``construction_env.py:177-494``, rescue
# ``rescue_env.py:121-277``).  Neither ``mujoco`` nor ``gymnasium`` is installed, so stand-in modules are registered whose
# ``MjModel.from_xml_string`` captures its argument; the *unmodified* reference module is imported from ``assets_root``
# and its constructor runs up to that call.  Only the compiled numeric tables are committed.
class _Captured(Exception):
    def __init__(self, xml):
        self.xml = xml


def _install_stubs():
    mj = types.ModuleType("mujoco")

    class MjModel:
        @staticmethod
        def from_xml_string(s, *a, **k):
            raise _Captured(s)

        @staticmethod
        def from_xml_path(p, *a, **k):
            with open(p) as f:
                raise _Captured(f.read())

    mj.MjModel = MjModel
    mj.MjData = object
    mj.viewer = types.ModuleType("mujoco.viewer")
    sys.modules["mujoco"] = mj
    sys.modules["mujoco.viewer"] = mj.viewer

    gym = types.ModuleType("gymnasium")

    class Env:
        pass

    gym.Env = Env
    gym.register = lambda *a, **k: None
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box = lambda *a, **k: None
    spaces.Discrete = lambda *a, **k: None
    spaces.Dict = lambda *a, **k: None
    utils = types.ModuleType("gymnasium.utils")
    seeding = types.ModuleType("gymnasium.utils.seeding")
    import numpy as np
    seeding.np_random = lambda seed=None: (np.random.default_rng(seed), seed)
    utils.seeding = seeding
    gym.spaces = spaces
    gym.utils = utils
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = lambda *a, **k: None
    envs.registration = reg
    gym.envs = envs
    for n, m in (("gymnasium", gym), ("gymnasium.spaces", spaces), ("gymnasium.utils", utils),
                 ("gymnasium.utils.seeding", seeding), ("gymnasium.envs", envs), ("gymnasium.envs.registration", reg)):
        sys.modules[n] = m


TARGETS = {
    "dancing": ("humanoid_dancing_env", "dancing_env", "HumanoidDancingEnv"),
    "soccer": ("humanoid_soccer_env", "soccer_env", "HumanoidSoccerEnv"),
    "martial_arts": ("humanoid_martial_arts_env", "martial_arts_env", "HumanoidMartialArtsEnv"),
    "construction": ("humanoid_construction_env", "construction_env", "HumanoidConstructionEnv"),
    "rescue": ("bipedal_rescue_env", "rescue_env", "BipedalRescueEnv"),
    "arm": ("robotic_arm_assembly_env", "assembly_env", "RoboticArmAssemblyEnv"),
}


def inline_mjcf(task: str, root: str) -> str:
    pkg, mod, cls = TARGETS[task]
    _install_stubs()
    sys.path.insert(0, os.path.join(root, pkg))
    cwd = os.getcwd()
    os.chdir("/tmp")
    # martial arts writes its scene to <module dir>/assets/martial_arts_scene.xml and reads it back before building the
    # model (martial_arts_env.py:379-381, 134-148); the reference tree is read-only, so writes under it go to memory
    import builtins, io
    real_open, real_makedirs, written = builtins.open, os.makedirs, {}

    class _MemFile(io.StringIO):
        def __init__(self, path): super().__init__(); self._path = path
        def close(self): written[self._path] = self.getvalue(); super().close()

    def mem_open(path, mode="r", *a, **k):
        ap = os.path.abspath(path) if isinstance(path, (str, os.PathLike)) else path
        if isinstance(ap, str) and ap.startswith(os.path.abspath(root)):
            if "w" in mode:
                return _MemFile(ap)
            if ap in written:
                return io.StringIO(written[ap])
        return real_open(path, mode, *a, **k)

    def mem_makedirs(path, *a, **k):
        if os.path.abspath(path).startswith(os.path.abspath(root)):
            return None
        return real_makedirs(path, *a, **k)

    builtins.open, os.makedirs = mem_open, mem_makedirs
    try:
        m = importlib.import_module(mod)
        try:
            getattr(m, cls)()
        except _Captured as c:
            return c.xml
        except OSError as e:
            raise RuntimeError(f"{task}: constructor failed before the model was built: {e}")
    finally:
        builtins.open, os.makedirs = real_open, real_makedirs
        os.chdir(cwd)
        sys.path.pop(0)
    raise RuntimeError(f"{task}: constructor finished without building a model")



COMPOSERS = {
    "quadruped_parkour": quadruped_parkour_mjcf,
    "humanoid_dancing": lambda root: inline_mjcf("dancing", root),
    "humanoid_soccer": lambda root: inline_mjcf("soccer", root),
    "bipedal_rescue": lambda root: inline_mjcf("rescue", root),
    "humanoid_construction": lambda root: inline_mjcf("construction", root),
    "humanoid_martial_arts": lambda root: inline_mjcf("martial_arts", root),
    # the arm loads a file (assembly_env.py:53-61); complete_model.xml has no <include>, so the file is the model
    "robotic_arm_assembly": lambda root: open(os.path.join(root, "robotic_arm_assembly_env", "assets", "complete_model.xml")).read(),
}

