#!/usr/bin/env python
"""Intermittent-difference hunt: the same forward N times, every kept tensor compared (bit for bit) with the first run's.
Prints the first tensor of the stage order that differs in each bad iteration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
import workloads

H4, W4, maxdisp, N = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (96, 312, 192, 300)))
if os.environ.get("POOL_SIMPLE"):
    d._lib.call("dca_pool_set_march", 0)
net = workloads.init_bench_weights_(d.GwcNet(maxdisp), 0).cuda().eval()
gf = [t.cuda() for t in workloads.feature_maps(3, 1, H4, W4)]
E = d.engine
order = ["volume", "dres0", "cost0"]
for s in (1, 2, 3):
    order += [f"cva{s}.{k}" for k in ("pooled", "cost_down", "logits", "class_map", "e", "S", "t", "fused", "c1", "c2", "out")]
order += ["classif3_logits", "pred_quarter", "mask", "pred4", "pv"]


def flat(keep, p4, pv):
    out = {"pred4": p4, "pv": pv}
    for k, v in keep.items():
        if isinstance(v, dict):
            for k2, v2 in v.items():
                out[f"{k}.{k2}"] = v2
        else:
            out[k] = v
    return {k: (v.t if isinstance(v, E.Planes) else v) for k, v in out.items() if v is not None}


def same(a, b):
    if a.is_floating_point():
        return bool(((a == b) | (a.isnan() & b.isnan())).all())
    return torch.equal(a, b)


def forward():
    keep = {}
    with torch.no_grad():
        p4, pv = net.hot_path(*gf, keep=keep)
    return flat(keep, p4, pv)


ref = {k: v.clone() for k, v in forward().items()}
nbad = 0
for it in range(N):
    got = forward()
    bad = [k for k in order if k in ref and not same(ref[k], got[k])]
    if bad:
        nbad += 1
        k = bad[0]
        a, b = ref[k], got[k]
        idx = (a != b).nonzero()
        print(f"iter {it}: first differing tensor {k} ({len(idx)} of {a.numel()} elements; shape {tuple(a.shape)}); "
              f"index range {idx.min(0).values.tolist()} .. {idx.max(0).values.tolist()}; all bad: {bad}", flush=True)
    del got
print(f"{nbad} bad iterations of {N}")
