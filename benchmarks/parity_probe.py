#!/usr/bin/env python
"""config-1 parity (256x512, maxdisp 192) of the hot path vs the live oracle under the kernel selectors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dcanet_b200 as d
from oracle import dcanet_oracle as O

feats = O.synth_features(0, 1, 64, 128, shift=3)
sd = O.calibrate_state_dict(O.synth_state_dict(0), feats, 192)
with torch.no_grad():
    ref4, refpv = O.hot_path(sd, *feats, maxdisp=192)
net = d.GwcNet(192)
own = net.state_dict(); own.update(sd); net.load_state_dict(own)
net = net.cuda().eval()
gf = [f.cuda() for f in feats]

def run(tag):
    with torch.no_grad():
        p4, pv = net.hot_path(*gf)
    dd = (p4.cpu() - ref4).abs()
    print(f"{tag:28s} max {float(dd.max()):.4f} mean {float(dd.mean()):.5f} p99.9 {float(dd.flatten().kthvalue(int(dd.numel()*0.999)).values):.4f}", flush=True)

E = d.engine
run("default")
run("default again")
with torch.no_grad():
    a = net.hot_path(*gf)[0]; b = net.hot_path(*gf)[0]
print("bit-reproducible:", bool(torch.equal(a, b)), "max diff", float((a - b).abs().max()), flush=True)
d._lib.call("dca_set_pdl", 0); run("no PDL")
with torch.no_grad():
    a = net.hot_path(*gf)[0]; b = net.hot_path(*gf)[0]
print("no PDL bit-reproducible:", bool(torch.equal(a, b)), flush=True)
d._lib.call("dca_set_pdl", 1)
d._lib.call("dca_tc_set_deconv_pair", 0); run("deconv single-tile"); d._lib.call("dca_tc_set_deconv_pair", 1)
d._lib.call("dca_pool_set_march", 0); run("avgpool simple"); d._lib.call("dca_pool_set_march", 1)
E.Options.fuse_tail = False; run("no fused 32->1 tail"); E.Options.fuse_tail = True
E.Options.prop_side_stream = False; run("no side stream"); E.Options.prop_side_stream = True
d._lib.call("dca_volume_set_v2", 0); run("volume generic"); d._lib.call("dca_volume_set_v2", 1)
d._lib.call("dca_attention_set_team", 0); run("attention 1 warp"); d._lib.call("dca_attention_set_team", 1)
d._lib.call("dca_tc_set_halo", 3); run("s2 per-tap"); d._lib.call("dca_tc_set_halo", 1)
E.Options.use_march = False; run("no march"); E.Options.use_march = True
E.Options.use_up2 = False; run("no up2"); E.Options.use_up2 = True
E.Options.cout1_on_tc = False; run("cout1 cuda-core"); E.Options.cout1_on_tc = True
E.Options.prop_on_tc = False; run("prop cuda-core"); E.Options.prop_on_tc = True
E.Options.use_tc = False; run("all cuda-core convs"); E.Options.use_tc = True
