"""Extracts the reference's only known-answer vectors for this path — 32 random RGB -> (c0,c1,c2) triples
plus white (src/tracer/color/spectrum/spectrum_tests.rs:37-111) — into tests/golden/spectrum_rgb_coeffs.json.
Run in the build container (reads /root/reference); the JSON is committed."""
import json, re, os
src = open("/root/reference/src/tracer/color/spectrum/spectrum_tests.rs").read()
body = src[src.index("const TEST_DATA"):]
rows = [[float(x) for x in m.groups()] for m in re.finditer(r"\[\s*(-?[\d.]+),\s*(-?[\d.]+),\s*(-?[\d.]+)\s*\]", body)]
assert len(rows) == 64
pairs = [{"rgb": rows[2 * i], "coeffs": rows[2 * i + 1]} for i in range(32)]
pairs.append({"rgb": [1.0, 1.0, 1.0], "coeffs": [0.001685, -2.276728, 807.041931]})
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "spectrum_rgb_coeffs.json")
json.dump({"source": "ekarpp/lumo src/tracer/color/spectrum/spectrum_tests.rs:37-111", "tolerance_reference": 1e-10 ** (1.0 / 3.0), "vectors": pairs}, open(out, "w"), indent=1)
print(out, len(pairs))
