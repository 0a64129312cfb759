import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, workloads
import dcanet_b200 as d
H, W, maxdisp, B = workloads.CONFIGS["kitti_384x1248"]
dev = torch.device("cuda", 0)
net = workloads.init_bench_weights_(d.GwcNet(maxdisp), 0).to(dev).eval()
host_sets = [workloads.feature_maps(s, B, H // 4, W // 4, pin=True) for s in range(4)]
for graph in (False, True, False, True):
    for depth in (2, 3):
        pipe = d.HotPathPipeline(net, depth=depth, graph=graph)
        for i in range(6):
            pipe.wait(pipe.submit(i, host_sets[i % 4]))
        K = 60
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(pipe.copy_stream)
        for i in range(K):
            last = pipe.submit(i, host_sets[i % 4])
        pipe.wait(last)
        e1.record(pipe.compute_stream)
        torch.cuda.synchronize()
        print(f"graph={graph} depth={depth}: {K / (e0.elapsed_time(e1) * 1e-3):.1f} pairs/s", flush=True)
        del pipe
