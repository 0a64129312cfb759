#!/usr/bin/env python
"""Kernel microbenchmarks (BASELINE.json configs[3]): volume kernel swept over groups {8,20,40} x D/4 {24,48,96}
at H4xW4 = 96x312, and the conv family at the KITTI shapes.  CUDA-event timing, prints one JSON per line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dcanet_b200 as d  # noqa: E402

E = d.engine


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def conv_case(mode, cin, cout, dims, planes, tc=True, halo=True, res=False, check=False):
    B, D, H, W = dims
    x = E.Planes(B, D, H, W, cin, planes, "cuda")
    x.t.normal_()
    transposed = mode == E.T3S2
    k = 1 if mode == E.K1 else 3
    w = torch.randn((cin, cout, k, k, k) if transposed else (cout, cin, k, k, k), device="cuda") * 0.05
    bn = torch.nn.BatchNorm3d(cout).cuda().eval()
    pc = E.PackedConv(w, bn, transposed)
    pc.pack_tc(planes, transposed)
    E.Options.use_tc = tc
    d._lib.call("dca_tc_set_halo", int(halo))
    y = E.conv(x, pc, mode, E.ACT_RELU)
    maxdiff = None
    if check:
        E.Options.use_tc = False
        yd = E.conv(x, pc, mode, E.ACT_RELU)
        E.Options.use_tc = tc
        maxdiff = float((y.to_ncdhw() - yd.to_ncdhw()).abs().max())
    r = E.Planes(y.B, y.D, y.H, y.W, y.C, planes, "cuda") if res else None
    if r is not None:
        r.t.normal_()
    ms = timeit(lambda: E.conv(x, pc, mode, E.ACT_RELU, res_pre=r))
    n_mac_vox = (x.D * x.H * x.W) if transposed else (y.D * y.H * y.W)
    fl = 2.0 * cin * cout * (k ** 3) * n_mac_vox * B
    E.Options.use_tc = True
    d._lib.call("dca_tc_set_halo", 1)
    return {"op": "conv", "mode": mode, "cin": cin, "cout": cout, "dims": dims, "planes": planes, "tc": tc, "halo": halo,
            "res": res, "maxdiff_vs_direct": maxdiff, "ms": round(ms, 4), "tflops_algorithmic": round(fl / ms / 1e9, 1)}


def volume_case(G, D4, planes, H4=96, W4=312):
    gl, gr = torch.randn(1, 320, H4, W4, device="cuda"), torch.randn(1, 320, H4, W4, device="cuda")
    cl, cr = torch.randn(1, 12, H4, W4, device="cuda"), torch.randn(1, 12, H4, W4, device="cuda")
    ms = timeit(lambda: E.fused_volume(gl, gr, cl, cr, D4, G, planes))
    Cv = (G + 24 + 7) // 8 * 8
    by = 2 * 332 * H4 * W4 * 4 + planes * Cv * D4 * H4 * W4 * 2
    return {"op": "volume", "groups": G, "D4": D4, "planes": planes, "Cv": Cv, "ms": round(ms, 4),
            "GBps_algorithmic": round(by / ms / 1e6, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="conv,volume")
    args = ap.parse_args()
    q4, q8 = (1, 48, 96, 312), (1, 24, 48, 156)
    if "conv" in args.what:
        for planes in (2, 1):
            for halo in (True, False):
                print(json.dumps(conv_case(E.K3S1, 32, 32, q4, planes, halo=halo)), flush=True)
            print(json.dumps(conv_case(E.K3S1, 32, 32, q4, planes, res=True)), flush=True)
            print(json.dumps(conv_case(E.K3S1, 64, 32, q4, planes)), flush=True)
            print(json.dumps(conv_case(E.K3S1, 32, 32, q8, planes)), flush=True)
            print(json.dumps(conv_case(E.K3S1, 64, 64, q8, planes)), flush=True)
            print(json.dumps(conv_case(E.K3S2, 32, 64, q4, planes)), flush=True)
            print(json.dumps(conv_case(E.T3S2, 64, 32, q8, planes)), flush=True)
            print(json.dumps(conv_case(E.K1, 32, 32, q4, planes)), flush=True)
        print(json.dumps(conv_case(E.K3S1, 32, 32, q4, 2, tc=False)), flush=True)
    if "dbg" in args.what:
        for planes in (2, 1):
            for dbg in (0, 1, 2, 3):
                d._lib.call("dca_tc_set_tuning", 1, dbg << 4)
                for (ci, co, dims) in ((32, 32, q4), (64, 32, q4)):
                    r = conv_case(E.K3S1, ci, co, dims, planes)
                    r.update(dbg=dbg)
                    print(json.dumps(r), flush=True)
        d._lib.call("dca_tc_set_tuning", 1, 0)
    if "volume" in args.what:
        for G in (8, 20, 40):
            for D4 in (24, 48, 96):
                print(json.dumps(volume_case(G, D4, 2)), flush=True)


if __name__ == "__main__":
    main()
