"""Per-kernel time and DRAM bytes of an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) with the
iteration log of the same run: python tools/launch_summary.py <launches.csv> <iterlog.json>"""
import csv, collections, re, json, sys
rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
per = collections.defaultdict(dict)
for r in rows:
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"): v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if r["Metric Name"].startswith("gpu__time"): v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "second": 1e3}.get(u, 1)
    per[int(r["ID"])][r["Metric Name"]] = v; per[int(r["ID"])]["name"] = re.sub(r"\(.*", "", r["Kernel Name"])
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, k in per.items():
    a = agg[k["name"]]; a[0] += 1; a[1] += k.get("gpu__time_duration.sum", 0); a[2] += k.get("dram__bytes_read.sum", 0); a[3] += k.get("dram__bytes_write.sum", 0)
T = sum(a[1] for a in agg.values())
log = json.load(open(sys.argv[2]))
nit = agg["k_queue_reset"][0]; its = log["iterations"][:nit]; closest = sum(a for a, b in its); shadow = sum(b for a, b in its)
print("# %s: first %d wave iterations of `tools/prof_run.py %s %d`: %d closest-hit rays (= bounces), %d shadow rays; cold-cache, serialised launches: compare shares" % (sys.argv[1].split("/")[-1], nit, log["workload"], log["spp"], closest, shadow))
print("# total %.1f ms" % T)
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-54s n=%4d %8.2f ms %5.1f%%  DRAM read %8.1f MB  write %8.1f MB   per bounce: read %6.0f B  write %6.0f B" % (n[:54], a[0], a[1], 100 * a[1] / T, a[2] / 1e6, a[3] / 1e6, a[2] / max(closest, 1), a[3] / max(closest, 1)))
