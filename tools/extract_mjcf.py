"""Print the MJCF string a reference env hands to ``mujoco.MjModel.from_xml_string`` (authoring container only).

Usage:  python tools/extract_mjcf.py dancing [/root/reference] > /tmp/dancing.xml
The capture technique lives in ``mujoco_gymnasium_environments_b200/compose.py::inline_mjcf``.
"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mujoco_gymnasium_environments_b200.compose import inline_mjcf

if __name__ == "__main__":
    sys.stdout.write(inline_mjcf(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "/root/reference"))
