#!/usr/bin/env python
"""Signed error of the tcgen05 convs vs the fp32 CUDA-core kernel on the same planes: is the tensor-core accumulation
error a systematic shrink (truncation toward zero) or random?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
E = d.engine
torch.manual_seed(0)
for cin, dims, relu_in in ((32, (1, 16, 32, 64), True), (64, (1, 16, 32, 64), True), (32, (1, 16, 32, 64), False)):
    x = torch.randn(1, cin, *dims[1:], device="cuda")
    if relu_in:
        x = torch.relu(x)
    w = torch.randn(32, cin, 3, 3, 3, device="cuda") * (2.0 / (cin * 27)) ** 0.5
    xp = E.Planes.from_ncdhw(x, 2)
    pc = E.PackedConv(w, None); pc.pack_tc(2)
    ref = torch.nn.functional.conv3d(xp.to_ncdhw().double(), w.double(), padding=1)
    outs = {}
    for tag, tc, march in (("march", True, True), ("halo", True, False), ("fp32 cuda-core", False, False)):
        E.Options.use_tc, E.Options.use_march = tc, march
        E.Options.march_min_items = 1
        y = E.conv(xp, pc, E.K3S1, E.ACT_NONE).to_ncdhw().double()
        e = y - ref
        scale = ref.abs().mean()
        big = ref.abs() > ref.abs().median()
        shrink = float((e[big] * torch.sign(ref[big])).mean() / ref[big].abs().mean())
        print(f"cin {cin} relu_in {relu_in} {tag:15s} rms rel {float(e.pow(2).mean().sqrt() / scale):.2e}  mean signed shrink {shrink:+.2e}  max rel {float(e.abs().max() / scale):.2e}", flush=True)
E.Options.use_tc = E.Options.use_march = True
