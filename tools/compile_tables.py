"""Compile each task's composed MJCF (read from the reference assets) into tables/*.npz.

Run in the authoring container:  python tools/compile_tables.py [/root/reference]
The GPU box has no /root/reference; it uses the committed tables.
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mujoco_gymnasium_environments_b200 import compose, mjcf

root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = os.path.join(os.path.dirname(__file__), "..", "mujoco_gymnasium_environments_b200", "tables")
os.makedirs(out, exist_ok=True)
for task, fn in compose.COMPOSERS.items():
    m = mjcf.compile_mjcf(fn(root), name=task)
    m.save(os.path.join(out, task + ".npz"))
    print(f"{task}: nq={m.nq} nv={m.nv} nu={m.nu} nbody={m.nbody} njnt={m.njnt} ngeom={m.ngeom} "
          f"npair={m.npair} ntree={m.ntree} nM={m.nM} meaninertia={m.meaninertia:.4f}")
