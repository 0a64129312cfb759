import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
from microbench import timeit
E = d.engine
gl, gr = torch.randn(1, 320, 96, 312, device="cuda"), torch.randn(1, 320, 96, 312, device="cuda")
cl, cr = torch.randn(1, 12, 96, 312, device="cuda"), torch.randn(1, 12, 96, 312, device="cuda")
for mode in (1, 3, 5, 7, 0):
    d._lib.call("dca_volume_set_v2", mode)
    for D in (24, 48):
        ms = timeit(lambda: E.fused_volume(gl, gr, cl, cr, D, 40, 2), 20)
        print(json.dumps({"mode": mode, "D": D, "ms": round(ms, 4)}), flush=True)
d._lib.call("dca_volume_set_v2", 1)
