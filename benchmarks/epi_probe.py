#!/usr/bin/env python
"""Timing probe: how much of the small-K tcgen05 kernels is epilogue (dbg bit0 = skip the epilogue math + stores)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
E = d.engine
from microbench import timeit

def main():
    P = 2
    q4, q8 = (1, 48, 96, 312), (1, 24, 48, 156)
    x4 = E.Planes(*q4, 32, P, "cuda"); x4.t.normal_()
    r4 = E.Planes(*q4, 32, P, "cuda"); r4.t.normal_()
    x8 = E.Planes(*q8, 64, P, "cuda"); x8.t.normal_()
    t8 = E.Planes(1, 48, 50, 158, 32, P, "cuda"); t8.t.normal_()
    bn = torch.nn.BatchNorm3d(32).cuda().eval()
    pc1 = E.PackedConv(torch.randn(32, 32, 1, 1, 1, device="cuda") * .1, bn); pc1.pack_tc(P)
    pcs2 = E.PackedConv(torch.randn(64, 32, 3, 3, 3, device="cuda") * .05, torch.nn.BatchNorm3d(64).cuda().eval()); pcs2.pack_tc(P)
    lib = d._lib.load()
    per_tap = lib.dca_pack_weights_tc_bytes(32, 64, 1, P)
    w28 = torch.zeros(28 * per_tap, dtype=torch.uint8, device="cuda")
    nb4 = lib.dca_pack_weights_tc_bytes(32, 32, 4, P)
    w4 = torch.zeros(nb4, dtype=torch.uint8, device="cuda")
    sc = torch.ones(32, device="cuda"); sh = torch.zeros(32, device="cuda")
    pc27 = E.PackedCout1(torch.randn(1, 32, 3, 3, 3, device="cuda") * .1, P)
    Pbuf = torch.empty((27, 48 * 96 * 312), dtype=torch.float32, device="cuda")
    cases = {
        "conv1_taps@1/4": lambda: d._lib.call("dca_conv1_taps_tc", x4.ptr, P, pc27.w_tc.data_ptr(), Pbuf.data_ptr(), 27, 1, 48, 96, 312, E._stream()),
        "k1_linear_32_32@1/4": lambda: E.conv(x4, pc1, E.K1, E.ACT_NONE),
        "s2_32_64@1/4": lambda: E.conv(x4, pcs2, E.K3S2, E.ACT_RELU),
        "up2_deconv64+side": lambda: E.up2(0, x8, x4, w28, sc, sh, E.ACT_RELU, 64, 24, 48, 156),
        "up2_deconv64+side+res": lambda: E.up2(0, x8, x4, w28, sc, sh, E.ACT_RELU, 64, 24, 48, 156, res_post=r4),
        "up2_bilinear_fuse": lambda: E.up2(2, t8, x4, w4, sc, sh, E.ACT_NONE, 32, 48, 48, 156),
    }
    for dbg in (0, 1, 2, 3):
        d._lib.call("dca_tc_set_tuning", 1, dbg << 4)
        for k, fn in cases.items():
            print(json.dumps({"case": k, "dbg": dbg, "ms": round(timeit(fn, 20), 4)}), flush=True)
    d._lib.call("dca_tc_set_tuning", 1, 0)

if __name__ == "__main__":
    main()
