#!/usr/bin/env python
"""Where the front end's time goes at KITTI size (left + right as one batch of 2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
import workloads
E, F = d.engine, d.frontend
H, W = 384, 1248
net = workloads.init_bench_weights_(d.GwcNet(192), 0).cuda().eval()
x = torch.randn(2, 3, H, W, device="cuda")
fe, g = net.feature_extraction, net.guidance


def t(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


with torch.no_grad(), F._no_tf32():
    print("fe stem (firstconv+layer1+layer2[0], cuDNN fp32, B=2) %.3f ms" % t(lambda: fe.layer2[0](fe.layer1(fe.firstconv(x)))))
    print("  firstconv %.3f  layer1 %.3f" % (t(lambda: fe.firstconv(x)), t(lambda: fe.layer1(fe.firstconv(x))) ))
    print("guidance stem (conv_start+layer1+layer2[0], B=1)        %.3f ms" % t(lambda: g.layer2[0](g.layer1(g.conv_start(x[:1])))))
    print("feature_extraction total (kernels)                      %.3f ms" % t(lambda: fe(x)))
    print("guidance total (kernels)                                %.3f ms" % t(lambda: g(x[:1])))
    F.Options.enabled = False
    print("feature_extraction total (torch)                        %.3f ms" % t(lambda: fe(x)))
    print("guidance total (torch)                                  %.3f ms" % t(lambda: g(x[:1])))
    F.Options.enabled = True
    fe(x); torch.cuda.synchronize()
    d._lib.PROFILE = []
    for _ in range(5):
        fe(x)
    agg = d._lib.profile_summary()
    d._lib.PROFILE = None
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms / 5:8.3f} ms {n // 5:4d}x  {k}")
    # per-shape timing of the 2-D conv kernel
    pk = E.cached_pack(fe, ("frontend", 2), lambda: None)
    s = fe.layer2[0](fe.layer1(fe.firstconv(x)))
    p = E.Planes.from_ncdhw(s, planes=2)
    b2, b3, b4 = pk.layer2[0], pk.layer3[1], pk.layer4[1]
    print("64->64 conv      %.1f us" % (1e3 * t(lambda: E.conv2d_tc(p, b2.c1, E.ACT_RELU))))
    print("64->64 conv +res %.1f us" % (1e3 * t(lambda: E.conv2d_tc(p, b2.c2, E.ACT_NONE, res=p))))
    p128 = E.Planes(2, 1, p.H, p.W, 128, 2, "cuda"); p128.t.normal_()
    print("128->128 conv    %.1f us" % (1e3 * t(lambda: E.conv2d_tc(p128, b3.c1, E.ACT_RELU))))
    print("128->128 dil 2   %.1f us" % (1e3 * t(lambda: E.conv2d_tc(p128, b4.c1, E.ACT_RELU, dil=2))))
    print("320->128 cat     %.1f us" % (1e3 * t(lambda: E.conv2d_tc_cat((p, p128, p128), pk.last0, E.ACT_RELU))))
    print("from_ncdhw       %.1f us" % (1e3 * t(lambda: E.Planes.from_ncdhw(s, planes=2))))
    out = torch.empty((2, 320, p.H, p.W), device="cuda")
    print("to nchw slice128 %.1f us" % (1e3 * t(lambda: E.planes_to_nchw_slice(p128, out, 64))))
