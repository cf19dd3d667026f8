#!/bin/bash
# usage: bench_all.sh <tag> [extra bench args]
tag=$1; shift
for t in quadruped_parkour humanoid_dancing humanoid_soccer bipedal_rescue humanoid_construction humanoid_martial_arts robotic_arm_assembly; do
  timeout 300 python bench.py --task $t --no-cpu-baseline "$@" > gpurun_out/bench_${tag}_$t.json 2> gpurun_out/bench_${tag}_$t.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}_$t.json").read().strip().splitlines()[-1])
    es = d["episode_stats"]
    print("$t", "value %.0f e2e %.0f ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "epb", d["config"]["envs_per_cta"], {k: es[k] for k in ("nan_resets","contacts_dropped","rows_dropped","arena_overflows","wide_passes","substeps")})
except Exception as e:
    print("$t FAILED", e); print(open("gpurun_out/bench_${tag}_$t.err").read()[-800:])
PY
done
