import sys
sys.path.insert(0, '/root/repo')
import torch, torch.nn.functional as F
import dcanet_b200 as d
E = d.engine
torch.manual_seed(0)
for (ci, co, bnflag, f32) in ((64, 64, False, False), (64, 128, True, False), (128, 64, False, True), (128, 144, False, True)):
    g = torch.randn(1, ci, 16, 8)
    c = torch.nn.Conv2d(ci, co, 3, padding=1, bias=False)
    with torch.no_grad():
        ref = c(g)
    gp = E.Planes.from_ncdhw(g.cuda(), 2)
    pc = E.PackedConv2dTc(c.weight.cuda(), None, 2)
    y = E.conv2d_tc(gp, pc, E.ACT_NONE, out_fp32=f32)
    got = (y[:, 0].permute(0, 3, 1, 2) if f32 else y.to_ncdhw()[:, :, 0]).cpu()
    err = (got - ref).abs()
    print(ci, co, 'max err', float(err.max()), 'ref max', float(ref.abs().max()))
    # per-channel error summary
    pe = err.amax(dim=(0, 2, 3))
    print('  bad channels:', [i for i in range(co) if pe[i] > 1e-3][:20], ' per-row err h:', [round(float(v), 3) for v in err.amax(dim=(0, 1, 3))][:18])
