import sys
sys.path.insert(0, '/root/repo')
import torch, dcanet_b200 as d
E = d.engine
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = E.Planes(1, 48, 96, 312, 32, planes, "cuda"); x.t.normal_()
w = torch.randn(32, 32, 3, 3, 3, device="cuda") * 0.05
pc = E.PackedConv(w, torch.nn.BatchNorm3d(32).cuda().eval()); pc.pack_tc(planes)
for _ in range(6):
    y = E.conv(x, pc, E.K3S1, E.ACT_RELU)
torch.cuda.synchronize(); print("ok")
