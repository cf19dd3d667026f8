"""Print the per-task table of a bench.py JSON line (file argument or stdin)."""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads([l for l in txt.strip().splitlines() if l.startswith("{")][-1])
print("HEAD value %.0f e2e %.0f graph %.0f" % (d["value"], d["e2e"]["value"], (d.get("graph_rollout") or {}).get("value", 0)), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["clocks"])
for r in d.get("per_task", []):
    es = r["episode_stats"]
    print("%-22s x%.1f value %8.0f e2e %8.0f graph %8.0f kern %6.2f ms [%5.2f..%5.2f] epb %d drops %d/%d/%d wide %d (rows %d) nan %d" % (
        r["task"], r["action_scale"], r["value"], r["e2e"]["value"], r["graph_rollout"]["value"], r["kernel_ms"], r["kernel_ms_min"], r["kernel_ms_max"],
        r["envs_per_cta"], es["contacts_dropped"], es["rows_dropped"], es["arena_overflows"], es["wide_passes"], es.get("wide_passes_rows", 0), es["nan_resets"]))
