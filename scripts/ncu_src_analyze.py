"""Summarise an `ncu --page source --csv --print-source sass` dump: per kernel, instruction mix,
stall samples per opcode class and the hottest instructions."""
import csv, sys, collections
path, pick = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
secs, cur = [], None
for row in csv.reader(open(path)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; secs.append(cur)
    elif cur is not None and row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
seen = set()
for s in secs:
    if pick not in s["name"] or s["name"] in seen: continue
    seen.add(s["name"])
    h = {n: i for i, n in enumerate(s["hdr"])}
    rows = s["rows"]
    tot_s = sum(int(r[h["# Samples"]]) for r in rows); tot_i = sum(int(r[h["Instructions Executed"]]) for r in rows)
    print("==", s["name"][:70], "instr", len(rows), "executed", tot_i, "samples", tot_s)
    mix = collections.defaultdict(lambda: [0, 0])
    for r in rows:
        op = r[h["Source"]].split()
        op = [o for o in op if not o.startswith("@")][0].split(".")[0]
        mix[op][0] += int(r[h["Instructions Executed"]]); mix[op][1] += int(r[h["# Samples"]])
    for op, (n, sm) in sorted(mix.items(), key=lambda kv: -kv[1][0])[:22]:
        print("  %-10s exec %6.2f%%  samples %6.2f%%" % (op, 100 * n / tot_i, 100 * sm / max(tot_s, 1)))
    stalls = [n for n in s["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(int(r[h[n]]) for r in rows) for n in stalls}
    print("  stalls:", ", ".join("%s %.1f%%" % (n[6:], 100 * v / max(tot_s, 1)) for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][h["# Samples"]]))[:top]
    for i in sorted(order):
        r = rows[i]
        st = sorted(((int(r[h[n]]), n[6:]) for n in stalls), reverse=True)[:2]
        print("  #%5d %5.2f%% x%-9s %-70s %s" % (i, 100 * int(r[h["# Samples"]]) / max(tot_s, 1), r[h["Instructions Executed"]], r[h["Source"]].strip()[:70], st))
