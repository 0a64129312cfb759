#!/usr/bin/env python
"""A tiny forward (hot path, front end kernels, baseline model) for compute-sanitizer:
    compute-sanitizer --tool memcheck python benchmarks/sanitize_forward.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
import workloads

net = workloads.init_bench_weights_(d.GwcNet(48), 0).cuda().eval()
f = [t.cuda() for t in workloads.feature_maps(1, 1, 16, 40)]
left = torch.randn(1, 3, 64, 160, device="cuda")
with torch.no_grad():
    p4, pv = net.hot_path(*f)
    q4, _ = net(left, torch.roll(left, -3, 3))
    base = workloads.init_bench_weights_(d.gwcnet.GwcNet_GC(48), 0).cuda().eval()
    vis = base.hot_path(f[0], f[1], f[2], f[3])
torch.cuda.synchronize()
print("ok", float(p4.mean()), float(q4.mean()), float(vis.mean()))
