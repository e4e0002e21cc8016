"""Top CUDA source lines of an ncu report by warp-stall samples (ncu -i REP --page source --csv --print-source cuda,sass).
usage: python tools/ncu_by_line.py report.ncu-rep [top_n]"""
import collections, csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = hdr = None
cols = ["# Samples", "stall_long_sb", "stall_barrier", "stall_no_inst", "stall_wait", "stall_short_sb", "Instructions Executed"]
agg = collections.defaultdict(lambda: [0] * len(cols))
src = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "" or r[2] != "-":
        continue
    key = (cur, int(r[0])); src[key] = r[1]
    for j, c in enumerate(cols):
        try: agg[key][j] += int(r[ix[c]])
        except ValueError: pass
tot = [sum(v[j] for v in agg.values()) for j in range(len(cols))]
print("totals", dict(zip(cols, tot)))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]}:{key[1]:4d} samp {v[0]:4d} lsb {v[1]:4d} bar {v[2]:4d} noi {v[3]:4d} wait {v[4]:4d} ssb {v[5]:4d} inst {v[6]:8d} | {src[key].strip()[:100]}")
