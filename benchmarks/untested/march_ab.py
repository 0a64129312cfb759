import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/benchmarks')
import torch, json
import dcanet_b200 as d
from microbench import timeit
E=d.engine
x=E.Planes(1,48,96,312,32,2,"cuda"); x.t.normal_()
pc=E.PackedConv(torch.randn(32,32,3,3,3,device="cuda")*0.05, torch.nn.BatchNorm3d(32).cuda().eval()); pc.pack_tc(2)
x6=E.Planes(1,48,96,312,64,2,"cuda"); x6.t.normal_()
pc6=E.PackedConv(torch.randn(32,64,3,3,3,device="cuda")*0.05, torch.nn.BatchNorm3d(32).cuda().eval()); pc6.pack_tc(2)
for split in (1,0):
    d._lib.call("dca_tc_set_march_split", split)
    print(split, round(timeit(lambda: E.conv(x,pc,E.K3S1,E.ACT_RELU),20),4), round(timeit(lambda: E.conv(x6,pc6,E.K3S1,E.ACT_RELU),20),4), flush=True)
