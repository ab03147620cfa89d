"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export per CUDA source line."""
import csv, sys, os, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = csv.reader(open(path))
fname = None; H = None; lines = []
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": H = r; continue
    if H is None or r[0] in ("", "..."): continue
    if len(r) < 10 or r[2] != "-": continue          # per-line summary rows have '-' in the Address column
    d = dict(zip(H[4:], r[4:]))
    lines.append((fname, r[0], r[1], d))
ts = sum(num(d["# Samples"]) for *_, d in lines); ti = sum(num(d["Instructions Executed"]) for *_, d in lines)
print("lines", len(lines), "samples", ts, "warp-instructions", ti)
stall_keys = [k for k in lines[0][3] if k.startswith("stall_") and "Not Issued" not in k]
tot_st = collections.Counter()
for *_, d in lines:
    for k in stall_keys: tot_st[k] += num(d[k])
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(ts, 1)) for k, v in tot_st.most_common(8)))
byfile = collections.Counter()
for f, ln, src, d in lines: byfile[f] += num(d["# Samples"])
print("by file:", dict((k, round(100 * v / ts, 1)) for k, v in byfile.items()))
for f, ln, src, d in sorted(lines, key=lambda x: -num(x[3]["# Samples"]))[:topn]:
    top = max(stall_keys, key=lambda k: num(d[k]))
    print("%-14s L%-4s samp %5.2f%% inst %5.2f%% thr %5s %-14s| %s" % (f, ln, 100 * num(d["# Samples"]) / ts, 100 * num(d["Instructions Executed"]) / ti, d["Avg. Threads Executed"], top[6:], src.strip()[:105]))
