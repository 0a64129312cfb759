#!/usr/bin/env python
"""Relative error of every hot-path boundary tensor vs the live oracle at KITTI full size: tensor-core (split-bf16)
path and the all-fp32 CUDA-core path side by side."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
from oracle import dcanet_oracle as O
E = d.engine
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3
feats = O.synth_features(seed, 1, 96, 312, shift=3)
sd = O.calibrate_state_dict(O.synth_state_dict(seed), feats, 192)
col = {}
with torch.no_grad():
    ref4, refpv = O.hot_path(sd, *feats, maxdisp=192, collect=col)
net = d.GwcNet(192); own = net.state_dict(); own.update(sd); net.load_state_dict(own); net = net.cuda().eval()
gf = [f.cuda() for f in feats]

def rel(got, ref):
    got = got.to_ncdhw().cpu() if hasattr(got, "to_ncdhw") else got.cpu().float()
    got = got[:, :ref.shape[1]] if got.dim() == ref.dim() and got.shape[1] != ref.shape[1] else got
    e = (got.reshape(ref.shape) - ref).abs()
    return float(e.max() / ref.abs().max()), float(e.mean() / ref.abs().mean())

for tag, tc in (("tensor-core", True), ("fp32 cuda-core", False)):
    E.Options.use_tc = tc
    keep = {}
    with torch.no_grad():
        p4, pv = net.hot_path(*gf, keep=keep)
    print("==", tag)
    pairs = [("volume", keep["volume"], col["volume"]), ("dres0", keep["dres0"], col["dres0"]), ("cost0", keep["cost0"], col["cost0"]),
             ("out1", keep["out1"], col["out1"]),
             ("cva2.out", keep["cva2"]["out"], col["cva2.out"]), ("cva3.out", keep["cva3"]["out"], col["cva3.out"]),
             ("cva3.logits", keep["cva3"]["logits"], col["cva3.logits"]),
             ("classif3_logits", keep["classif3_logits"], col["classif3_logits"]),
             ("pred_quarter", keep["pred_quarter"], col["pred_quarter"]), ("pred4", p4, col["pred4"])]
    for name, g, r in pairs:
        mx, mn = rel(g, r)
        print(f"  {name:18s} rel max {mx:.2e} rel mean {mn:.2e}", flush=True)
E.Options.use_tc = True
