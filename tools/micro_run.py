"""micro closest-hit benchmark only (bench.micro_trace) for a workload: python tools/micro_run.py bunny"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lumo_b200 import native
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "bunny"
prog, blob, integrator, _ = bench.build_workload(name)
ctx = native.GpuContext(0)
G = native.GpuScene(ctx, blob)
dev = torch.device("cuda", 0)
print(json.dumps({name: {k: round(v["mrays_per_s"], 1) for k, v in bench.micro_trace(G, dev, torch, np).items()}}))
