"""Write a text summary of an .ncu-rep (key metrics + top source lines) for profiles/."""
import subprocess, sys, csv, io
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ["Duration", "Executed Ipc Active", "Issue Slots Busy", "Warp Cycles Per Issued Instruction", "Registers Per Thread",
        "Dynamic Shared Memory Per Block", "Block Limit Registers", "Block Limit Shared Mem", "Block Limit Warps",
        "Theoretical Occupancy", "Achieved Occupancy", "DRAM Throughput", "Memory Throughput", "Grid Size", "Block Size",
        "L1/TEX Hit Rate", "L2 Hit Rate", "Shared Memory Configuration Size", "Local"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "launch__registers_per_thread"]
hot = subprocess.run([sys.executable, "tools/ncu_hot.py", rep, "30"], capture_output=True, text=True).stdout
with open(out, "w") as f:
    f.write(f"# {title}\n# source: {rep} (ncu --set full --clock-control none --import-source on; one launch)\n\n## details page (selected)\n")
    for l in det.splitlines():
        if any(k in l for k in keys):
            f.write(l.rstrip() + "\n")
    f.write("\n## raw metrics (selected)\n")
    if len(rows) >= 3:
        for h, u, v in zip(rows[0], rows[1], rows[2]):
            if h in want or "issue_stalled" in h and "per_issue_active" in h:
                f.write(f"{h} [{u}] = {v}\n")
    f.write("\n## hottest CUDA source lines (warp-stall samples / executed warp-instructions)\n" + hot)
    ic = subprocess.run([sys.executable, "tools/ncu_icache.py", rep], capture_output=True, text=True).stdout
    f.write("\n## instruction-fetch view (tools/ncu_icache.py)\n" + ic)
print("wrote", out)
