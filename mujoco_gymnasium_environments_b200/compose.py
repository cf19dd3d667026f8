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

import os
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


COMPOSERS = {
    "quadruped_parkour": quadruped_parkour_mjcf,
}
