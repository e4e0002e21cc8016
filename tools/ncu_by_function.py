"""Aggregate an `ncu --page source --csv` export of k_step by device function (ELF symbol ranges of the cubin).
usage: python tools/ncu_by_function.py <source.csv> <cubin> [kernel_prefix]"""
import csv, subprocess, sys, re, collections
src, cubin = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "_Z6k_step"
syms = []
for line in subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout.splitlines():
    p = line.split()
    if len(p) >= 8 and p[3] == "FUNC" and kern in p[7]:
        size = int(p[2], 0)
        name = p[7].split("$")[-1] if "$" in p[7] else p[7]
        syms.append((int(p[1], 16), size, name))
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ithr = hdr.index("Thread Instructions Executed")
base = int(rows[2][ia], 16)
main = [s for s in syms if "$" not in s[2] and s[0] == 0]
agg = collections.defaultdict(lambda: [0, 0, 0])
stall_cols = [(i, n[6:]) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
stalls = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    if len(r) <= ithr: continue
    off = int(r[ia], 16) - base
    name = "k_step(main)"
    for a, sz, n in syms:
        if a != 0 and a <= off < a + sz:
            name = n; break
    agg[name][0] += int(r[ii]); agg[name][1] += int(r[isamp]); agg[name][2] += int(r[ithr])
    for i, n in stall_cols:
        if r[i] not in ("", "0"): stalls[name][n] += int(r[i])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"{'function':45s} {'inst %':>7s} {'samples %':>9s} {'lanes':>6s}  top stall reasons (% of the function's samples)")
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    nm = re.sub(r'^_ZN\d+_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+', '', k)
    nm = re.sub(r'^_Z[0-9]+', '', nm)
    m2 = re.match(r'^(\d+)(.*)$', nm)
    if m2:
        nm = m2.group(2)[:int(m2.group(1))]
    print(f"{nm[:45]:45s} {100*v[0]/ti:7.2f} {100*v[1]/max(ts,1):9.2f} {v[2]/max(v[0],1):6.1f}  " +
          " ".join(f"{n}:{100*c/max(sum(stalls[k].values()),1):.0f}" for n, c in stalls[k].most_common(5)))
print("total inst", ti, "samples", ts)
