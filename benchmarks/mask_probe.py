#!/usr/bin/env python
"""Full-size parity (max/mean |d disp|, class-mask mismatches) vs the live oracle, per seed and kernel selector.
    python benchmarks/mask_probe.py [H4xW4] seed...      (default 96x312 = KITTI 384x1248)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
from oracle import dcanet_oracle as O
E = d.engine
args = [a for a in sys.argv[1:]]
H4, W4 = 96, 312
if args and "x" in args[0]:                      # e.g. 136x240 = SceneFlow 544x960 at 1/4 resolution
    H4, W4 = (int(v) for v in args.pop(0).split("x"))
seeds = [int(a) for a in args] or [3]
for seed in seeds:
    feats = O.synth_features(seed, 1, H4, W4, shift=3)
    sd = O.calibrate_state_dict(O.synth_state_dict(seed), feats, 192)
    col = {}
    with torch.no_grad():
        ref4, refpv = O.hot_path(sd, *feats, maxdisp=192, collect=col)
    net = d.GwcNet(192); own = net.state_dict(); own.update(sd); net.load_state_dict(own); net = net.cuda().eval()
    gf = [f.cuda() for f in feats]

    def run(tag):
        keep = {}
        with torch.no_grad():
            p4, pv = net.hot_path(*gf, keep=keep)
        dd = (p4.cpu() - ref4).abs()
        dq = (keep["pred_quarter"].cpu() - col["pred_quarter"]).abs() if "pred_quarter" in col else None
        mm = [int((keep[f"cva{s}"]["class_map"].cpu().long() != col[f"cva{s}.class_map"]).sum()) for s in (1, 2, 3)]
        n_bad = int((dd > 0.05).sum())
        print(f"seed {seed} {tag:22s} max {float(dd.max()):.4f} mean {float(dd.mean()):.5f} >0.05: {n_bad} px"
              + (f" | quarter-res max {float(dq.max()):.4f}" if dq is not None else "") + f" masks {mm}", flush=True)

    run("default")
    d._lib.call("dca_tc_set_trunc_comp", 0.0); run("no trunc. compensation"); d._lib.call("dca_tc_set_trunc_comp", 1.56e-8)
    E.Options.prop_on_tc = False; run("prop cuda-core"); E.Options.prop_on_tc = True
    E.Options.cout1_on_tc = False; run("cout1 cuda-core"); E.Options.cout1_on_tc = True
    E.Options.use_tc = False; run("all convs cuda-core"); E.Options.use_tc = True
    for st in ("dres", "cva", "cls3"):
        E.Options.fp32_stages = frozenset([st]); run(f"fp32 stage {st}")
    E.Options.fp32_stages = frozenset()
