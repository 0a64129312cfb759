"""Stage-count variants of the reference (models/gwcnet_dca{0,1,2,4}_g.py: 0/1/2/4 cva stages, head classif<N>).

CPU: the oracle's N-stage forward against fixtures produced by the reference's own variant modules
(tests/golden/make_golden_variants.py); the state_dict layout of this repo's variant modules against the reference's key
lists; the kernel sequence their forward would launch (dry run).
"""
import collections
import importlib
import os

import numpy as np
import pytest
import torch

import _dryrun
from _util import GOLD, _t
from oracle import dcanet_oracle as O

VARIANTS = (0, 1, 2, 4)


def load_variant(n):
    z = np.load(os.path.join(GOLD, f"variant_dca{n}_32x64_d48.npz"))
    H, W, maxdisp, seed, nn_ = [int(v) for v in z["meta"]]
    assert nn_ == n
    sd = O.synth_state_dict(seed, num_cva=max(n, 3))
    for k in z.files:
        if k.startswith("bn:"):
            sd[k[3:]] = _t(z[k])
    feats = [_t(z[k]) for k in ("gwc_l", "gwc_r", "cat_l", "cat_r", "g")]
    return z, sd, feats, maxdisp


@pytest.mark.parametrize("n", VARIANTS)
def test_oracle_n_stage_forward_matches_the_reference_variant(n):
    z, sd, feats, maxdisp = load_variant(n)
    with torch.no_grad():
        pred4, pv = O.hot_path(sd, *feats, maxdisp=maxdisp, num_cva=n, pv_stage=n)
    ref = _t(z["pred"]).reshape(pred4.shape)        # dca0/1/2 return pred.squeeze(1), dca4 keeps the channel
    assert float((pred4 - ref).abs().max()) < 2e-4
    assert float((pv - _t(z["pv"])).abs().max()) < 2e-5


@pytest.mark.parametrize("n", VARIANTS)
def test_variant_state_dict_layout_matches_reference(n):
    mod = importlib.import_module(f"cost-volume-aggregation-in-stereo-matching-revisited_b200.gwcnet_dca{n}_g")
    ref = [l.strip().split(" ", 1) for l in open(os.path.join(GOLD, f"state_dict_keys_dca{n}.txt"))]
    sd = mod.GwcNet(48).state_dict()
    assert [k for k, _ in ref] == list(sd)                      # same keys in the same order
    for k, shp in ref:
        assert str(tuple(sd[k].shape)) == shp, k
    assert mod.GwcNet_G(48).use_concat_volume is False and mod.GwcNet_GC(48).use_concat_volume is True


@pytest.mark.parametrize("n", VARIANTS)
def test_variant_kernel_sequence(n):
    """N stages launch the 3-stage sequence with the cva block repeated N times (10 launches per stage)."""
    import dcanet_b200 as d
    mod = importlib.import_module(f"cost-volume-aggregation-in-stereo-matching-revisited_b200.gwcnet_dca{n}_g")
    tr, (pred, pv) = _dryrun.forward_trace(mod.GwcNet(48).eval(), 16, 32)
    base, _ = _dryrun.forward_trace(d.GwcNet(48).eval(), 16, 32)
    assert len(base) == 41 and len(tr) == 41 + 10 * (n - 3)
    per_stage = collections.Counter(t[0] for t in base[5:15])
    got = collections.Counter(t[0] for t in tr)
    want = collections.Counter(t[0] for t in base)
    for k, v in per_stage.items():
        want[k] += v * (n - 3)
    assert got == +want
    assert pred.shape == (1, 1, 64, 128)                        # hot_path(); forward() squeezes per variant:
    assert mod.GwcNet.SQUEEZE_PRED == (n != 4)                  # gwcnet_dca{0,1,2}_g.py return pred.squeeze(1), dca4 does not
    assert pv.shape == ((1, 12, 16, 32) if n == 0 else (1, 6, 8, 16))
