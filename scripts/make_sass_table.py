"""CPU-side evidence for profiles/: per hot kernel of the in-tree libasw_b200.so, the SASS instruction classes
(cuobjdump -sass) that show the Blackwell-native pieces (UTMALDG / UBLKCP = TMA, SYNCS = mbarrier, FFMA2 / FMUL2 = packed
FP32, USETMAXREG = setmaxnreg) and the instruction mix of the tap loops; plus the `nvcc -Xptxas -v` register / spill report.
usage: make_sass_table.py <round tag>"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "stereo_matchin_b200", "libasw_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
dem = subprocess.run(["c++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()
hot = ("k_vagg_v2<8", "k_hagg_split<", "k_wta_v2", "k_raw_v2", "k_support_v2", "k_vpad_v2", "k_unpack_v2", "k_support_ref", "k_pack_support",
       "k_ref_to_volume_v2", "k_volume_to_ref_v2", "k_vden_to_ref", "k_wta_merge")
cols = ["UTMALDG", "UBLKCP", "SYNCS", "USETMAXREG", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "LDS", "LDG", "STG", "MUFU", "IMAD", "BRA", "STL", "LDL"]
out = ["# SASS instruction classes per kernel of stereo_matchin_b200/libasw_b200.so (static counts, cuobjdump -sass; sm_100a)",
       "# UTMALDG = cp.async.bulk.tensor (tiled TMA), UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg,",
       "# FFMA2/FMUL2/FADD2 = packed FP32 (fma/mul/add.rn.f32x2), STL/LDL = local-memory spills", "",
       "%-62s %6s " % ("kernel", "total") + " ".join("%7s" % c for c in cols)]
for name, d in zip(funcs, dem):
    short = d.split("(")[0].replace("void ", "").replace("asw::", "")
    if not any(h in short for h in hot):
        continue
    c = funcs[name]
    out.append("%-62s %6d " % (short[:62], sum(c.values())) + " ".join("%7d" % c.get(k, 0) for k in cols))
open(os.path.join(ROOT, "profiles", f"{tag}_sass_classes.txt"), "w").write("\n".join(out) + "\n")
# ptxas -v
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas=-v", "-ccbin", "/usr/bin/g++",
       "-I" + os.path.join(ROOT, "include"), "-c", "-o", "/dev/null", os.path.join(ROOT, "stereo_matchin_b200", "csrc", "asw_api.cu")]
r = subprocess.run(cmd, capture_output=True, text=True)
lines, fn = ["# nvcc -Xptxas -v (sm_100a) for stereo_matchin_b200/csrc/asw_api.cu: registers, spills, shared memory per kernel",
              "# (k_vagg_v2: the reported count is the launch allocation for 384 threads; the math warps raise theirs to 232 with setmaxnreg,",
              "#  the 56-96 spilled bytes belong to the helper warps that run with 40 registers)", ""], None
for line in r.stderr.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
    elif "bytes stack frame" in line and fn:
        spill = line.strip()
    elif "Used" in line and fn:
        lines.append("%-60s %s | %s" % (fn[:60], line.split(":", 1)[1].strip(), spill))
        fn = None
open(os.path.join(ROOT, "profiles", f"{tag}_ptxas.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(out[:4]))
print(len(out) - 5, "kernels;", len(lines) - 4, "ptxas entries")
