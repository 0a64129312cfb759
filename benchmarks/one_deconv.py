"""Time dca_up2_tc kind 0 (transposed conv 64->32 + redir) at the KITTI shape with tuning flags.
    python benchmarks/one_deconv.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dcanet_b200 as d
import workloads

E, L = d.engine, d._lib
net = workloads.init_bench_weights_(d.GwcNet(192), 0).cuda().eval()
pk = net.packed().cva[0].agg
B, D8, H8, W8 = 1, 24, 48, 156
c2 = E.Planes(B, D8, H8, W8, 64, 2, "cuda"); c2.t.normal_()
fused = E.Planes(B, 2 * D8, 2 * H8, 2 * W8, 32, 2, "cuda"); fused.t.normal_()
res = E.Planes(B, 2 * D8, 2 * H8, 2 * W8, 32, 2, "cuda"); res.t.normal_()
fd = pk.conv3_fused


def run(n=20, with_res=True):
    for _ in range(3):
        E.up2(0, c2, fused, fd.w_tc, fd.scale, fd.shift, E.ACT_RELU, 64, D8, H8, W8, res_post=res if with_res else None)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        E.up2(0, c2, fused, fd.w_tc, fd.scale, fd.shift, E.ACT_RELU, 64, D8, H8, W8, res_post=res if with_res else None)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for name, pair, flags in (("up2 kernel (one tile per fetch)", 0, 0), ("pair kernel", 1, 0), ("pair, no L2 prefetch", 1, 4 << 4),
                          ("pair, no epilogue stores", 1, 1 << 4), ("pair, no MMAs", 1, 2 << 4),
                          ("pair, no MMAs no stores", 1, 3 << 4)):
    L.call("dca_tc_set_deconv_pair", pair)
    L.call("dca_tc_set_tuning", 1, flags)
    print(f"{name:36s} {run():7.1f} us   (no res_post: {run(with_res=False):7.1f} us)")
L.call("dca_tc_set_tuning", 1, 0)
L.call("dca_tc_set_deconv_pair", 1)
