#!/usr/bin/env python
"""images -> disparity: how far does the disparity move when only the front end's arithmetic changes?
kernel front end vs cuDNN fp32 front end vs CPU (ATen) front end, same hot path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import dcanet_b200 as d
from test_gpu_frontend import _net, _images

net = _net(maxdisp=96)
left, right = _images(1, 128, 256, seed=2)


def q(tag, a, b):
    dd = (a - b).abs().flatten()
    n = dd.numel()
    print(f"{tag:42s} max {float(dd.max()):.4f} mean {float(dd.mean()):.5f} p99 {float(dd.kthvalue(int(n * .99)).values):.4f} "
          f"p99.9 {float(dd.kthvalue(int(n * .999)).values):.4f} frac>0.05 {float((dd > 0.05).float().mean()):.5f}")


with torch.no_grad():
    d.frontend.Options.enabled = False
    ref4, _ = net(left, right)
    # CPU front end (different summation order, true fp32), same hot path
    cpu = net.feature_extraction.cpu()
    f = cpu(torch.cat((left, right)).cpu())
    net.feature_extraction.cuda()
    g = net.guidance.cpu()(left.cpu())["g"]
    net.guidance.cuda()
    gwc, cat = f["gwc_feature"].cuda(), f["concat_feature"].cuda()
    cpu4, _ = net.hot_path(gwc[:1].contiguous(), gwc[1:].contiguous(), cat[:1].contiguous(), cat[1:].contiguous(), g.cuda())
    d.frontend.Options.enabled = True
    k4, _ = net(left, right)
q("kernel front end vs cuDNN fp32 front end", k4, ref4)
q("CPU ATen front end vs cuDNN fp32 front end", cpu4, ref4)
q("kernel front end vs CPU ATen front end", k4, cpu4)

# ---- per-block error growth against an fp64 CPU run of the same blocks
import copy
E = d.engine
fe = net.feature_extraction
with torch.no_grad(), d.frontend._no_tf32():
    x = torch.cat((left, right))
    s = fe.layer2[0](fe.layer1(fe.firstconv(x)))
    fe64 = copy.deepcopy(fe).double().cpu()
    t64 = s.double().cpu()
    t32 = s.clone()
    pk = d.frontend.PackedFeatureExtraction(fe, 2)
    p = E.Planes.from_ncdhw(s, planes=2)
    blocks = [("layer2", i + 1, b) for i, b in enumerate(pk.layer2)] + [("layer3", i, b) for i, b in enumerate(pk.layer3)] + \
             [("layer4", i, b) for i, b in enumerate(pk.layer4)]
    for name, i, kb in blocks:
        t64 = getattr(fe64, name)[i](t64)
        t32 = getattr(fe, name)[i](t32)
        p = kb(p)
        out = torch.empty((p.B, p.C, p.H, p.W), dtype=torch.float32, device="cuda")
        E.planes_to_nchw_slice(p, out, 0)
        sc = float(t64.abs().max())
        print(f"{name}[{i}]: |x|max {sc:9.3f}  kernel err {float((out.double().cpu() - t64).abs().max()) / sc:.2e}  "
              f"cuDNN fp32 err {float((t32.double().cpu() - t64).abs().max()) / sc:.2e}")

with torch.no_grad(), d.frontend._no_tf32():
    xin = torch.cat((left, right))
    f64 = fe64(xin.double().cpu())
    g64 = copy.deepcopy(net.guidance).double().cpu()(left.double().cpu())["g"]
    d.frontend.Options.enabled = False
    ft = net.feature_extraction(xin); gt = net.guidance(left)["g"]
    d.frontend.Options.enabled = True
    fk = net.feature_extraction(xin); gk = net.guidance(left)["g"]
    for name, t, a, b in (("gwc_feature", f64["gwc_feature"], fk["gwc_feature"], ft["gwc_feature"]),
                          ("concat_feature", f64["concat_feature"], fk["concat_feature"], ft["concat_feature"]),
                          ("g", g64, gk, gt)):
        sc = float(t.abs().max())
        ea, eb = (a.double().cpu() - t).abs(), (b.double().cpu() - t).abs()
        print(f"{name:16s} |x|max {sc:9.3f} kernel err max {float(ea.max()) / sc:.2e} rms {float(ea.pow(2).mean().sqrt()) / sc:.2e}   "
              f"cuDNN err max {float(eb.max()) / sc:.2e} rms {float(eb.pow(2).mean().sqrt()) / sc:.2e}")

with torch.no_grad():
    a4, _ = net(left, right); b4, _ = net(left, right)
    q("kernel FE net() twice", a4, b4)
    d.frontend.Options.enabled = False
    c4, _ = net(left, right); e4, _ = net(left, right)
    d.frontend.Options.enabled = True
    q("cuDNN FE net() twice", c4, e4)
    hk, _ = net.hot_path(fk["gwc_feature"][:1], fk["gwc_feature"][1:], fk["concat_feature"][:1], fk["concat_feature"][1:], gk)
    ht, _ = net.hot_path(ft["gwc_feature"][:1], ft["gwc_feature"][1:], ft["concat_feature"][:1], ft["concat_feature"][1:], gt)
    q("hot_path(kernel feats) vs hot_path(cuDNN feats)", hk, ht)
    q("hot_path(kernel feats) vs net() kernel", hk, a4)
    q("hot_path(cuDNN feats) vs net() cuDNN", ht, c4)
    hm, _ = net.hot_path(ft["gwc_feature"][:1], ft["gwc_feature"][1:], ft["concat_feature"][:1], ft["concat_feature"][1:], gk)
    q("cuDNN feats + kernel g vs all cuDNN", hm, ht)
    hm, _ = net.hot_path(fk["gwc_feature"][:1], fk["gwc_feature"][1:], ft["concat_feature"][:1], ft["concat_feature"][1:], gt)
    q("kernel gwc + rest cuDNN vs all cuDNN", hm, ht)
    hm, _ = net.hot_path(ft["gwc_feature"][:1], ft["gwc_feature"][1:], fk["concat_feature"][:1], fk["concat_feature"][1:], gt)
    q("kernel concat + rest cuDNN vs all cuDNN", hm, ht)
    f64c = [f64["gwc_feature"][:1].float().cuda(), f64["gwc_feature"][1:].float().cuda(), f64["concat_feature"][:1].float().cuda(),
            f64["concat_feature"][1:].float().cuda(), g64.float().cuda()]
    h64, _ = net.hot_path(*[t.contiguous() for t in f64c])
    q("hot_path(fp64 feats) vs hot_path(kernel feats)", h64, hk)
    q("hot_path(fp64 feats) vs hot_path(cuDNN feats)", h64, ht)
