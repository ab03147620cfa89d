"""DRAM traffic per unit of work from an ncu launch list with dram__bytes_{read,write}.sum (per launch, cold cache, serialised) and
the iteration log of the same run (tools/prof_run.py with PROF_ITERLOG): sums the kernels of each class over the captured wave
iterations and divides by the rays / bounces of those iterations.
usage: python tools/traffic_from_ncu.py <launches.csv> <iterlog.json> <workload> [profiles/traffic.json]"""
import csv, json, re, sys, collections
rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
log = json.load(open(sys.argv[2])); workload = sys.argv[3]
CLASS = [("trace", r"k_closest_|k_wave_classify|k_wave_trace"), ("occlude", r"k_occl_|k_wave_occlude|k_shadow_apply"), ("shade", r"k_terminal|k_scatter|k_nee_|k_terms_reset"), ("regen", r"k_retire|k_compact")]
per_id = collections.defaultdict(dict)
for r in rows:
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    per_id[int(r["ID"])][r["Metric Name"]] = v
    per_id[int(r["ID"])]["name"] = r["Kernel Name"]
# wave iterations in launch order: every k_queue_reset ends one
it = 0; dram = collections.defaultdict(float); n_it = 0
for i in sorted(per_id):
    k = per_id[i]
    name = k["name"]
    for cls, rx in CLASS:
        if re.search(rx, name): dram[cls] += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
    if "k_queue_reset" in name: n_it += 1
iters = log["iterations"][:n_it]
closest = sum(a for a, b in iters); shadow = sum(b for a, b in iters)
out = {"trace": {"dram_bytes_per_unit": dram["trace"] / max(closest, 1), "unit": "ray"},
       "occlude": {"dram_bytes_per_unit": dram["occlude"] / max(shadow, 1), "unit": "ray"},
       "shade": {"dram_bytes_per_unit": dram["shade"] / max(closest, 1), "unit": "bounce"}}
for v in out.values():
    v["source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum over the first %d wave iterations of `tools/prof_run.py %s %d` (%s), per launch, cold cache" % (n_it, workload, log["spp"], sys.argv[1].split("/")[-1])
    v["iterations"] = n_it; v["closest_rays"] = closest; v["shadow_rays"] = shadow
print(json.dumps(out, indent=1))
if len(sys.argv) > 4:
    try: db = json.load(open(sys.argv[4]))
    except Exception: db = {}
    db[workload] = out
    json.dump(db, open(sys.argv[4], "w"), indent=1)
