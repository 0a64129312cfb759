#!/usr/bin/env python
"""Run N hot-path forwards (KITTI shape, synthetic) -- the command profiled by ncu for profiles/*.
    python profiles/run_forward.py [--n 3] [--precision parity|fast]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dcanet_b200 as d  # noqa: E402
import workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=3)
ap.add_argument("--precision", default="parity")
ap.add_argument("--config", default="kitti_384x1248")
a = ap.parse_args()
H, W, maxdisp, B = workloads.CONFIGS[a.config]
net = workloads.init_bench_weights_(d.GwcNet(maxdisp, precision=a.precision), 0).cuda().eval()
feats = workloads.feature_maps(0, B, H // 4, W // 4, device="cuda")
with torch.no_grad():
    for _ in range(a.n):
        out = net.hot_path(*feats)
torch.cuda.synchronize()
print("ok", float(out[0].mean()))
