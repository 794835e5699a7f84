"""Turns the artefacts of scripts/gpu_round_end.sh (gpurun_out/) into the committed summaries in
profiles/ (round tag as argv[1], default r01)."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)

def launches(src, dst):
    rows = list(csv.reader(open(src)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    tot, per = 0.0, {}
    out = ["kernel,grid,block,duration_ms"]
    for r in rows[h + 1:]:
        name = r[4].split("(")[0].replace("void ", "")
        ms = float(r[-1].replace(",", "")) / 1e6
        tot += ms
        per[name] = per.get(name, 0.0) + ms
        out.append(f'"{name}","{r[8]}","{r[7]}",{ms:.4f}')
    out.append("")
    out.append("# share of the summed kernel time (ncu serialises launches and runs them cold: compare shares)")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]):
        out.append(f"# {k}: {v:.3f} ms = {100 * v / tot:.1f} %")
    open(dst, "w").write("\n".join(out) + "\n")

if os.path.exists(os.path.join(G, "bench_launches.csv")):
    launches(os.path.join(G, "bench_launches.csv"), os.path.join(P, f"{tag}_bench_launches.csv"))
rep = os.path.join(G, "prof_agg.ncu-rep")
if os.path.exists(rep):
    raw = os.path.join(G, "prof_agg_raw.csv")
    subprocess.run(f"ncu -i {rep} --page raw --csv > {raw}", shell=True, check=False, stderr=subprocess.DEVNULL)
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), raw], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_ncu_full_summary.txt"), "w").write(
        "# ncu --set full --clock-control none, one cfg3 frame (1800x1500x256, r=2): V/H aggregation kernels, FIRST and steady-state launches\n" + txt)
    # DRAM traffic per launch of the steady-state aggregation kernels (bench.py reports it as roofline.traffic)
    rows = list(csv.reader(open(raw)))
    hd = {n: i for i, n in enumerate(rows[0])}
    traffic = {}
    for r in rows[2:]:
        name = r[hd["Kernel Name"]]
        key = "k_vagg_v2" if "k_vagg_v2<8, 0>" in name or "k_vagg_v2<(int)8, (bool)0>" in name else \
              "k_hagg_split" if "k_hagg_split<0" in name or "k_hagg_split<(bool)0" in name else None
        if key:
            def gb(col):
                v, u = float(r[hd[col]].replace(",", "")), rows[1][hd[col]]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
            rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
            traffic[key] = {"kernel": name.split("(")[0], "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                            "duration_ms_under_ncu": float(r[hd["gpu__time_duration.sum"]].replace(",", "")) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(rows[1][hd["gpu__time_duration.sum"]], 1.0),
                            "source": f"ncu --set full --clock-control none on one cfg3 frame; summary in profiles/{tag}_ncu_full_summary.txt"}
    if traffic:
        json.dump(traffic, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
for f, d in (("ubench.json", f"{tag}_ubench.json"), ("bench_cfg3.json", f"{tag}_bench_cfg3.json"), ("bench_reference.json", f"{tag}_bench_reference.json"),
             ("pytest_gpu.log", f"{tag}_pytest_gpu.log"), ("smoke.log", f"{tag}_smoke.log")):
    if os.path.exists(os.path.join(G, f)):
        shutil.copyfile(os.path.join(G, f), os.path.join(P, d))
stages = {}
for wl in ("cfg2", "cfg3", "cfg4", "cfg5"):
    f = os.path.join(G, f"stage_{wl}.json")
    if os.path.exists(f):
        try:
            stages[wl] = json.loads(open(f).read().strip().splitlines()[-1])
        except Exception:
            pass
if stages:
    json.dump(stages, open(os.path.join(P, f"{tag}_stage_times.json"), "w"), indent=1)
print("profiles written:", sorted(os.listdir(P)))
