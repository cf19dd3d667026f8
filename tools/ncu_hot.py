"""Summarise an ncu report's source page: top CUDA source lines by stall samples and by executed instructions.
usage: python tools/ncu_hot.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != "-": continue          # keep only per-source-line summary rows (Address == '-')
    try:
        s = int(r[idx["# Samples"]]); ie = int(r[idx["Instructions Executed"]])
    except ValueError:
        continue
    key = (cur, r[0])
    a = agg.setdefault(key, [0, 0, r[1]])
    a[0] += s; a[1] += ie
tot_s = sum(a[0] for a in agg.values()) or 1; tot_i = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot_s}  total warp-instructions {tot_i}")
print("---- by stall samples")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot_s:5.1f}% smp {100*a[1]/tot_i:5.1f}% ins  {f}:{ln}: {a[2].strip()[:120]}")
print("---- by instructions")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top // 2]:
    print(f"{100*a[0]/tot_s:5.1f}% smp {100*a[1]/tot_i:5.1f}% ins  {f}:{ln}: {a[2].strip()[:120]}")
