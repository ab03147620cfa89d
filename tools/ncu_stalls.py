"""Summarises an `ncu --page source --csv` export of one kernel: stall mix, lanes per instruction, opcode mix, hottest regions.
usage: python tools/ncu_stalls.py <report.ncu-rep> <kernel regex> [launch index]"""
import csv, subprocess, sys, collections, re
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:120])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
f = lambda r, h: float(r[ix[h]] or 0)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for h in stalls: tot[h] += f(r, h)
S = sum(tot.values())
print("stall mix:", ", ".join("%s %.1f%%" % (h[6:], 100 * v / S) for h, v in tot.most_common(8)))
wi = sum(f(r, "Instructions Executed") for r in data); ti = sum(f(r, "Thread Instructions Executed") for r in data)
print("SASS instructions %d, warp instructions %.3g, thread instructions %.3g, lanes per instruction %.2f" % (len(data), wi, ti, ti / max(wi, 1)))
ops = collections.Counter(); ops_s = collections.Counter()
for r in data:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    op = m.group(1) if m else "?"
    ops[op] += f(r, "Instructions Executed"); ops_s[op] += f(r, "# Samples")
print("opcode mix (warp instr %, samples %):", ", ".join("%s %.1f/%.1f" % (o, 100 * v / wi, 100 * ops_s[o] / max(sum(ops_s.values()), 1)) for o, v in ops.most_common(14)))
# hottest 256-instruction regions by samples
R = 256
reg = [(sum(f(r, "# Samples") for r in data[k:k + R]), k) for k in range(0, len(data), R)]
reg.sort(reverse=True)
ts = sum(s for s, _ in reg)
print("hottest regions of %d instructions (samples %%, first address, lanes):" % R)
for s, k in reg[:8]:
    w = sum(f(r, "Instructions Executed") for r in data[k:k + R]); t = sum(f(r, "Thread Instructions Executed") for r in data[k:k + R])
    print("  %5.1f%%  idx %6d  %s  lanes %.1f" % (100 * s / ts, k, data[k][0], t / max(w, 1)))
