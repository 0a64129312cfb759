import sys, os
sys.path.insert(0, '/root/repo')
import torch, dcanet_b200 as d, workloads
net = workloads.init_bench_weights_(d.GwcNet(192), 0).cuda().eval()
feats = workloads.feature_maps(0, 1, 96, 312, device="cuda")
for dbg in (0, 1, 2, 3):
    d._lib.call("dca_tc_set_tuning", 1, dbg << 4)
    with torch.no_grad():
        for _ in range(2): net.hot_path(*feats)
        torch.cuda.synchronize()
        d._lib.PROFILE = []
        for _ in range(3): net.hot_path(*feats)
        agg = d._lib.profile_summary(); d._lib.PROFILE = None
    print("dbg", dbg, {k.split(' dims')[0][:40]: round(v[1]/3, 3) for k, v in agg.items() if 'up2' in k or 'mode=3' in k or 'mode=1' in k})
