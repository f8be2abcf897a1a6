#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --print-source cuda,sass --csv` dump: stall samples and executed warp
instructions by CUDA source line of one kernel launch.  Usage: ncu_src_summary.py dump.csv <kernel substring> [launch index] [top N]"""
import csv
import sys
import collections

csv.field_size_limit(10 ** 9)
path, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(path)))
secs, cur = [], None
for r in rows:
    if r and r[0] in ("Function Name", "Kernel Name"):
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
        continue
    if cur is not None:
        cur["rows"].append(r)
# consecutive sections with the same name belong to one launch (one per source file); a launch starts at the largest section
launches, last = [], None
for s in secs:
    if last is not None and s["name"] == last["name"] and len(s["rows"]) < len(launches[-1][0]["rows"]):
        launches[-1].append(s)
    else:
        launches.append([s])
    last = s
sel = [l for l in launches if pat in l[0]["name"]]
L = sel[which]
print("kernel:", L[0]["name"][:100], "| launch", which, "of", len(sel))
tot_s = tot_i = 0
lines = []
stall_tot = collections.Counter()
for fi, s in enumerate(L):
    hdr = s["rows"][0]
    ix = {h: i for i, h in enumerate(hdr)}
    # columns: "Line No","Source",(Address,Source for sass rows)...
    i_samp = hdr.index("Warp Stall Sampling (All Samples)")
    i_inst = hdr.index("Instructions Executed")
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    for r in s["rows"][1:]:
        if len(r) <= i_inst or not r[0].strip().isdigit():
            continue            # sass rows have an empty line number
        try:
            sm, ins = float(r[i_samp] or 0), float(r[i_inst] or 0)
        except ValueError:
            continue
        st = {h: float(r[i] or 0) for h, i in stall_cols if len(r) > i and r[i] not in ("", "-")}
        for h, v in st.items():
            stall_tot[h] += v
        tot_s += sm; tot_i += ins
        lines.append((sm, ins, fi, int(r[0]), r[1].strip()[:110], st))
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
print("stall totals:", ", ".join(f"{h[6:]} {v / max(tot_s, 1) * 100:.1f}%" for h, v in stall_tot.most_common(9)))
for key, title in ((0, "by stall samples"), (1, "by executed instructions")):
    print("----", title)
    for sm, ins, fi, ln, src, st in sorted(lines, key=lambda t: -t[key])[:top]:
        main = max(st.items(), key=lambda kv: kv[1])[0][6:] if st and sm else ""
        print(f"{sm / max(tot_s, 1) * 100:5.1f}% smp {ins / max(tot_i, 1) * 100:5.1f}% ins  f{fi}:{ln:<5d} {main:12s} {src}")
