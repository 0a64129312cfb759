"""The plain-GwcNet baseline models/gwcnet.py (three full hourglass blocks, eval returns the `vis_tsne1`
tensor, gwcnet.py:186-190,236-244).  CPU: the oracle's restatement against a fixture made by the reference's own module
(tests/golden/make_golden_baseline.py), and the state_dict layout of this repo's module.  GPU: the module on the kernels."""
import importlib
import os

import numpy as np
import pytest
import torch

from _util import GOLD, _t
from oracle import dcanet_oracle as O


def load_baseline():
    z = np.load(os.path.join(GOLD, "baseline_gwcnet_gc_32x64_d48.npz"))
    H, W, maxdisp, seed = [int(v) for v in z["meta"]]
    sd = O.synth_state_dict_from_keys(os.path.join(GOLD, "state_dict_keys_gwcnet_gc.txt"), seed)
    for k in z.files:
        if k.startswith("bn:"):
            sd[k[3:]] = _t(z[k])
    feats = [_t(z[k]) for k in ("gwc_l", "gwc_r", "cat_l", "cat_r")]
    return z, sd, feats, maxdisp


def test_oracle_baseline_matches_the_reference_module():
    z, sd, feats, maxdisp = load_baseline()
    with torch.no_grad():
        vis = O.baseline_hot_path(sd, *feats, maxdisp=maxdisp)
    ref = _t(z["vis"])
    assert vis.shape == ref.shape == (1, 24, 67, 120)
    assert float((vis - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("tag,concat", [("g", False), ("gc", True)])
def test_baseline_state_dict_layout_matches_reference(tag, concat):
    mod = importlib.import_module("cost-volume-aggregation-in-stereo-matching-revisited_b200.gwcnet")
    ref = [l.strip().split(" ", 1) for l in open(os.path.join(GOLD, f"state_dict_keys_gwcnet_{tag}.txt"))]
    sd = mod.GwcNet(48, concat).state_dict()
    assert [k for k, _ in ref] == list(sd)
    for k, shp in ref:
        assert str(tuple(sd[k].shape)) == shp, k
    assert mod.GwcNet(48).use_concat_volume is False                 # the reference's default (gwcnet.py:108)


@pytest.mark.gpu
def test_baseline_module_on_the_kernels_matches_the_reference_fixture():
    import dcanet_b200 as d
    z, sd, feats, maxdisp = load_baseline()
    net = d.gwcnet.GwcNet(maxdisp, True)
    own = net.state_dict()
    own.update(sd)
    net.load_state_dict(own)
    net = net.cuda().eval()
    with torch.no_grad():
        vis = net.hot_path(*[f.cuda() for f in feats])
    ref = _t(z["vis"])
    err = float((vis.cpu() - ref).abs().max()) / max(1.0, float(ref.abs().max()))
    print("baseline vis_tsne1: max rel err %.2e" % err)
    assert vis.shape == ref.shape and err < 2e-4


@pytest.mark.gpu
def test_baseline_without_concat_volume_matches_the_oracle():
    """GwcNet_G (the reference's default: 40-channel volume, dres0.0 with 40 input channels, zero-padded to the 64-channel
    plane row here) against the oracle run live, BN statistics calibrated by the oracle."""
    import dcanet_b200 as d
    sd = O.synth_state_dict_from_keys(os.path.join(GOLD, "state_dict_keys_gwcnet_g.txt"), 3)
    feats = O.synth_features(3, 1, 12, 24, shift=2)[:2]
    with torch.no_grad():
        O.baseline_hot_path(sd, feats[0], feats[1], None, None, maxdisp=64, calibrate=True, vis_size=(8, 5, 12))
        ref = O.baseline_hot_path(sd, feats[0], feats[1], None, None, maxdisp=64)
    net = d.gwcnet.GwcNet_G(64)
    own = net.state_dict()
    own.update(sd)
    net.load_state_dict(own)
    net = net.cuda().eval()
    with torch.no_grad():
        vis = net.hot_path(feats[0].cuda(), feats[1].cuda())
    err = float((vis.cpu() - ref).abs().max()) / max(1.0, float(ref.abs().max()))
    print("baseline (no concat volume): max rel err %.2e" % err)
    assert vis.shape == ref.shape == (1, 24, 67, 120) and err < 2e-4
