"""Pin the oracle to the real ``mujoco.mj_step`` in one command (runs wherever ``import mujoco`` works).

    python tools/pin_with_mujoco.py [--reference /root/reference] [--write] [--tasks quadruped_parkour,...]

For every physics fixture under tests/golden/ (the per-task files and wide_states.npz) the stored *states*
(qpos, qvel, ctrl, qacc_warmstart) are replayed through MuJoCo itself on the model compiled from the reference's own MJCF
(``compose.COMPOSERS`` reads the assets / captures the inline generators from the reference checkout):
mj_forward -> ncon, contact geom pairs, dist, nefc, qacc;  mj_step -> qpos1, qvel1 (and the stored multi-step horizons).
The script prints a diff table (oracle-made expectation vs MuJoCo) -- the places DESIGN.md section 2 lists as
[EXT-unverified] (pair order, capsule tangent hint, box-box contact sets, cylinder pairs, the per-forward warm start) are
where differences would show -- and with --write replaces the expected values, so that tests/test_oracle_*.py and the GPU
parity tests are pinned to MuJoCo from then on.  Model dimensions (nq nv nu nbody ngeom) are checked against the compiled
tables first.  Without mujoco it says so and exits 0: nothing is pinned in this container or on the GPU box (neither has
the wheel).

Reference call sites replayed: quadruped_parkour_env/parkour_env.py:348,368; humanoid_dancing_env/dancing_env.py:849;
robotic_arm_assembly_env/assembly_env.py:229 and the siblings listed in SURVEY.md section 8(c).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
TASKS = ["quadruped_parkour", "humanoid_dancing", "humanoid_soccer", "bipedal_rescue", "humanoid_construction",
         "humanoid_martial_arts", "robotic_arm_assembly"]
HORIZONS = {"qpos5": 5, "qpos10": 10}           # multi-step expectations some fixtures carry


def fixtures(task):
    """(path, key prefix) of every physics fixture holding states of ``task``."""
    out = []
    p = os.path.join(GOLD, task + ".npz")
    if os.path.exists(p):
        out.append((p, ""))
    w = os.path.join(GOLD, "wide_states.npz")
    if os.path.exists(w) and f"{task}__qpos" in np.load(w).files:
        out.append((w, task + "__"))
    return out


def replay(mujoco, model, z, pre):
    """MuJoCo's values for every expectation the fixture holds, keyed like the fixture."""
    n = z[pre + "qpos"].shape[0]
    new = {}
    cap = z[pre + "pairs"].shape[1] if pre + "pairs" in z else 0
    for k in range(n):
        d = mujoco.MjData(model)

        def load():
            d.qpos[:] = z[pre + "qpos"][k]; d.qvel[:] = z[pre + "qvel"][k]; d.ctrl[:] = z[pre + "ctrl"][k]
            d.qacc_warmstart[:] = z[pre + "warm"][k]; d.time = 0.0
        load(); mujoco.mj_forward(model, d)
        pairs = np.full((cap, 2), -1, np.int32); dist = np.zeros(cap)
        for i in range(min(d.ncon, cap)):
            pairs[i] = (d.contact[i].geom1, d.contact[i].geom2); dist[i] = d.contact[i].dist
        row = dict(ncon=d.ncon, nefc=d.nefc, pairs=pairs, dist=dist, qacc=d.qacc.copy(), iters=int(d.solver_niter[0]) if hasattr(d, "solver_niter") else 0)
        load(); mujoco.mj_step(model, d)
        row.update(qpos1=d.qpos.copy(), qvel1=d.qvel.copy(), warm1=d.qacc_warmstart.copy())
        for key, h in HORIZONS.items():
            if pre + key in z:
                load()
                for _ in range(h):
                    mujoco.mj_step(model, d)
                row[key] = d.qpos.copy()
        for key, v in row.items():
            if pre + key in z:
                new.setdefault(pre + key, []).append(v)
    return {k: np.array(v) for k, v in new.items()}


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--reference", default=os.environ.get("B2_REFERENCE_ROOT", "/root/reference"))
    ap.add_argument("--write", action="store_true", help="replace the expected values in tests/golden/ by MuJoCo's")
    ap.add_argument("--tasks", default=",".join(TASKS))
    ap.add_argument("--native-ccd", action="store_true",
                    help="leave MuJoCo's native GJK/EPA convex path on (default from 3.2.3); by default it is switched off so that cylinder pairs "
                         "take libccd's MPR, the algorithm oracle/mjstep_ref.c::mpr_convex and csrc/b2_mpr.cuh restate")
    a = ap.parse_args()
    try:
        import mujoco
    except ImportError as e:
        print(f"pin_with_mujoco: `import mujoco` failed ({e}); nothing replayed, parity stays unpinned.")
        return 0
    from mujoco_gymnasium_environments_b200 import compose
    from mujoco_gymnasium_environments_b200.tasks import load_tables
    if not os.path.isdir(a.reference):
        print(f"pin_with_mujoco: reference checkout {a.reference} not found (needed for the MJCF); nothing replayed.")
        return 0
    print(f"mujoco {mujoco.__version__}; reference at {a.reference}")
    print(f"{'task':24s} {'fixture':22s} {'quantity':8s} {'states':>6s} {'max |oracle - mujoco|':>22s} {'rel':>10s}  note")
    for task in a.tasks.split(","):
        xml = compose.COMPOSERS[task](a.reference)
        model = mujoco.MjModel.from_xml_string(xml)
        native = getattr(getattr(mujoco, "mjtDisableBit", None), "mjDSBL_NATIVECCD", None)
        if native is not None and not a.native_ccd:
            model.opt.disableflags |= int(native)        # MuJoCo >= 3.2.3: route convex pairs through libccd MPR again
        t = load_tables(task)
        dims = {k: (int(getattr(t, k)), int(getattr(model, k))) for k in ("nq", "nv", "nu", "nbody", "ngeom")}
        bad = {k: v for k, v in dims.items() if v[0] != v[1]}
        if bad:
            print(f"{task:24s} MODEL DIMENSIONS DIFFER (tables, MjModel): {bad} -- compiler parity first (SURVEY 8(f)1)")
            continue
        for path, pre in fixtures(task):
            z = dict(np.load(path))
            new = replay(mujoco, model, z, pre)
            for key, val in sorted(new.items()):
                old = z[key]
                if key.endswith(("pairs", "ncon", "nefc", "iters")):
                    same = int(np.sum(np.all(np.reshape(old == val, (old.shape[0], -1)), axis=1)))
                    print(f"{task:24s} {os.path.basename(path):22s} {key[len(pre):]:8s} {old.shape[0]:6d} {'identical in ' + str(same):>22s} {'':>10s}  bit-exact quantity")
                else:
                    diff = float(np.max(np.abs(old.astype(np.float64) - val))); ref = float(np.max(np.abs(val))) + 1e-300
                    print(f"{task:24s} {os.path.basename(path):22s} {key[len(pre):]:8s} {old.shape[0]:6d} {diff:22.3e} {diff / ref:10.2e}")
            if a.write:
                z.update({k: v.astype(z[k].dtype) for k, v in new.items()})
                z["pinned_by"] = np.array(f"mujoco {mujoco.__version__}")
                np.savez_compressed(path, **z)
                print(f"{task:24s} {os.path.basename(path):22s} rewritten with MuJoCo's values")
    print("task-level fixtures (task_obs / task_rew / task_term) are made by oracle/tasks_ref.py; re-run tools/make_golden*.py after "
          "pinning, or compare against the reference classes with bench.py --impl reference (kind = \"reference\").")
    return 0


if __name__ == "__main__":
    sys.exit(main())
