#!/usr/bin/env python
"""pairs/s of one forward over B pairs (the grids scale with B: better fill of the last wave of every persistent kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, workloads
import dcanet_b200 as d
cfg = sys.argv[1] if len(sys.argv) > 1 else "kitti_384x1248"
H, W, maxdisp, _ = workloads.CONFIGS[cfg]
dev = torch.device("cuda", 0)
net = workloads.init_bench_weights_(d.GwcNet(maxdisp), 0).to(dev).eval()
with torch.no_grad():
    ref = None
    for B in (1, 2, 4, 1, 2, 4):
        sets = [[t.to(dev) for t in workloads.feature_maps(10 * s, B, H // 4, W // 4)] for s in range(2)]
        for i in range(3):
            out = net.hot_path(*sets[i % 2])
        if B == 1 and ref is None:
            ref = net.hot_path(*sets[0])[0].clone()
        if B == 2:          # batch element 0 of a B = 2 forward = the B = 1 forward of the same pair?
            one = [t[:1].contiguous() for t in sets[0]]
            a = net.hot_path(*sets[0])[0][:1]
            b = net.hot_path(*one)[0]
            print("   batch element 0 equals the single-pair forward:", bool(torch.equal(a, b)), float((a - b).abs().max()))
        K = max(8, 96 // B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            net.hot_path(*sets[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"B={B}: {ms:.3f} ms per forward, {B / (ms * 1e-3):.1f} pairs/s", flush=True)
        del sets
        torch.cuda.empty_cache()
