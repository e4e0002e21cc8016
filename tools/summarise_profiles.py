"""Turn the captures of tools/profile_r02.sh (gpurun_out/r02_final_*) into the tracked summaries under profiles/:
launch-list summary, DRAM traffic per hoic_step (profiles/r02_dram_traffic.json, read by bench.py), raw ncu pages and
per-function tables of the pipeline kernels.   python tools/summarise_profiles.py"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = os.environ.get("TAG", "r02_final")      # capture set (tools/profile_r02.sh)


def rows_of(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    return rows[hi], rows[hi + 2:]


def launch_summary(src, dst, title):
    h, rows = rows_of(src)
    ik, iv, im = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    agg = collections.defaultdict(list)
    for r in rows:
        if len(r) > iv and r[im] == "gpu__time_duration.sum":
            agg[r[ik].split("(")[0][:60]].append(float(r[iv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(title + "\n")
        for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
            f.write(f"{k:62s} n={len(v):5d} mean={sum(v)/len(v)/1e3:9.1f} us  total={sum(v)/1e6:9.3f} ms {100*sum(v)/tot:5.1f}%\n")
    return agg


if os.path.exists(os.path.join(G, TAG + "_bench_launches.csv")):
    shutil.copy(os.path.join(G, TAG + "_bench_launches.csv"), os.path.join(P, TAG + "_bench_launches.csv"))
    launch_summary(os.path.join(G, TAG + "_bench_launches.csv"), os.path.join(P, TAG + "_bench_launch_summary.txt"),
                   "ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 1 --preroll 2 --no-extras --no-cpu "
                   "(box, 4096 envs; kernels of the replayed CUDA graphs are profiled one by one: cold, serialised)")

def traffic(tp, tag, note):
    h, rows = rows_of(tp)
    ik, iv, im, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("Metric Unit")
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    units, n_tail = {}, 0
    for r in rows:
        if len(r) <= iv:
            continue
        k = r[ik].split("(")[0]
        per[k][r[im]] += float(r[iv].replace(",", ""))
        units[r[im]] = r[iu]
        if k == "k_tail" and r[im] == "gpu__time_duration.sum":
            n_tail += 1
    steps = max(n_tail / 2.0, 1.0)             # two batch parts per hoic_step
    sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rs, ws = sc[units["dram__bytes_read.sum"]], sc[units["dram__bytes_write.sum"]]
    tot = sum(v["dram__bytes_read.sum"] * rs + v["dram__bytes_write.sum"] * ws for k, v in per.items() if k.startswith("k_")) / steps
    with open(os.path.join(P, f"{TAG}_step_traffic_{tag}.txt"), "w") as f:
        f.write(f"ncu dram__bytes_read.sum + dram__bytes_write.sum and durations of the launches of {steps:.1f} hoic_step calls (box, 4096 envs, "
                f"2 parts), per step; {note}:\n")
        for k, v in sorted(per.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
            if k.startswith("k_"):
                f.write(f"{k:10s} read {v['dram__bytes_read.sum']*rs/steps/1e6:9.2f} MB  write {v['dram__bytes_write.sum']*ws/steps/1e6:9.2f} MB  "
                        f"time {v['gpu__time_duration.sum']/steps/1e6:8.3f} ms (serialised)\n")
        f.write(f"total {tot/1e6:.1f} MB per hoic_step = {tot/4096/1e3:.1f} KB per env-step\n")
    return tot


cold, warm = os.path.join(G, TAG + "_step_traffic.csv"), os.path.join(G, TAG + "_step_traffic_warm.csv")
res = {}
if os.path.exists(cold):
    res["box:4096:cold_caches"] = traffic(cold, "cold", "ncu's default cache control: L2 flushed before every kernel, so every kernel re-reads its part "
                                                        "of the working set from DRAM -- an upper bound, not what a run does")
if os.path.exists(warm):
    res["box:4096"] = traffic(warm, "warm", "--cache-control none: caches as the preceding kernels left them, i.e. what a run does; this is the figure "
                                            "bench.py reports as roofline.traffic")
if res:
    json.dump(res, open(os.path.join(P, "r02_dram_traffic.json"), "w"))
    print(res)

cub = "/tmp/cub_final"
os.makedirs(cub, exist_ok=True)
for kn, obj in (("k_p1", "k_p1"), ("k_p2", "k_p2"), ("k_p3s", "k_p3s"), ("k_tail", "k_tail"), ("banana_k_p2", "k_p2")):
    rep = os.path.join(G, f"{TAG}_{kn}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = os.path.join(P, f"{TAG}_{kn}_raw.csv")
    subprocess.run(f"ncu -i {rep} --page raw --csv > {raw} 2>/dev/null", shell=True)
    src = f"/tmp/{TAG}_{kn}_src.csv"
    subprocess.run(f"ncu -i {rep} --page source --csv > {src} 2>/dev/null", shell=True)
    subprocess.run(f"cd {cub} && cuobjdump -xelf all {ROOT}/hoic_b200/build/{obj}.o > /dev/null 2>&1", shell=True)
    kernel = "_Z5k_p3s" if obj == "k_p3s" else ("_Z6k_tail" if obj == "k_tail" else "_Z4" + obj)
    subprocess.run(f"python {ROOT}/tools/ncu_by_function.py {src} {cub}/{obj}.sm_100a.cubin {kernel} > {P}/{TAG}_{kn}_by_function.txt 2>/dev/null", shell=True)
    print(kn, "done")
