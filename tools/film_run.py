"""k_film_encode alone on a synthetic 3840x2160 film (for ncu): python tools/film_run.py [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lumo_b200 import native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
H, W = 2160, 3840
torch.manual_seed(1)
w = torch.rand(H, W, 1, dtype=torch.float64, device="cuda") * 40 + 0.5
px = torch.cat([torch.exp(torch.randn(H, W, 3, dtype=torch.float64, device="cuda") * 2 - 2) * w, w], -1).contiguous()
sp = torch.rand(H, W, 3, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ctx = native.GpuContext(0)
for i in range(n):
    flush.zero_(); torch.cuda.synchronize()
    rgb, ms = ctx.film_encode_dev(px.data_ptr(), sp.data_ptr(), (H, W), 1.0 / 64, 1.0, 0)
    print("launch %d: %.4f ms, %.0f GB/s" % (i, ms, H * W * 59 / ms / 1e6))
ctx.close()
