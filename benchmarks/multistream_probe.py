#!/usr/bin/env python
"""Throughput with independent pairs in flight on several streams (the kernels are single-wave persistent grids: a second
pair's kernels can take the SMs that the tail of the first pair's kernel leaves idle)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, workloads
import dcanet_b200 as d
H, W, maxdisp, B = workloads.CONFIGS["kitti_384x1248"]
dev = torch.device("cuda", 0)
net = workloads.init_bench_weights_(d.GwcNet(maxdisp), 0).to(dev).eval()
sets = [[t.to(dev) for t in workloads.feature_maps(s, B, H // 4, W // 4)] for s in range(4)]
K = 120
with torch.no_grad():
    for ns in (1, 2, 3, 1, 2):
        streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]
        for i in range(2 * ns):
            with torch.cuda.stream(streams[i % ns]):
                net.hot_path(*sets[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(K):
            with torch.cuda.stream(streams[i % ns]):
                net.hot_path(*sets[i % 4])
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        print(f"{ns} stream(s): {K / (e0.elapsed_time(e1) * 1e-3):.1f} pairs/s", flush=True)
