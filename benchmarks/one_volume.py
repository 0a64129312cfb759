#!/usr/bin/env python
"""A few launches of the fused volume kernel at the KITTI shape (for `ncu --set full`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
E = d.engine
gl, gr = torch.randn(1, 320, 96, 312, device="cuda"), torch.randn(1, 320, 96, 312, device="cuda")
cl, cr = torch.randn(1, 12, 96, 312, device="cuda"), torch.randn(1, 12, 96, 312, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    v = E.fused_volume(gl, gr, cl, cr, 48, 40, 2)
torch.cuda.synchronize(); print("ok")
