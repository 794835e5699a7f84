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
