"""Pin the CPU oracle against vectors produced by the reference itself (tests/golden/make_golden.py).
CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import dcanet_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


from _util import _t, load_e2e  # noqa: E402


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_volume_ops_match_reference(tag):
    z = np.load(os.path.join(GOLD, "ops_small.npz"))
    L, R = _t(z[f"gwc_{tag}_L"]), _t(z[f"gwc_{tag}_R"])
    D, G = [int(v) for v in z[f"gwc_{tag}_meta"]]
    assert torch.equal(O.build_gwc_volume(L, R, D, G), _t(z[f"gwc_{tag}_out"]))
    assert torch.equal(O.build_concat_volume(L[:, :12], R[:, :12], D), _t(z[f"cat_{tag}_out"]))


def test_regression_matches_reference():
    z = np.load(os.path.join(GOLD, "ops_small.npz"))
    x = _t(z["regress_in"])
    got = O.disparity_regression(torch.softmax(x, 1), 12)
    assert torch.allclose(got, _t(z["regress_out"]), atol=1e-6)


def test_hot_path_matches_reference_every_boundary():
    z, sd, maxdisp = load_e2e()
    col = {}
    with torch.no_grad():
        pred4, pv2 = O.hot_path(sd, _t(z["gwc_l"]), _t(z["gwc_r"]), _t(z["cat_l"]), _t(z["cat_r"]),
                                _t(z["g"]), maxdisp, collect=col)
    tol = dict(atol=2e-5, rtol=1e-4)
    assert torch.allclose(col["dres0"], _t(z["dres0"]), **tol)
    for s in (1, 2, 3):
        p = f"cva{s}"
        assert torch.allclose(col[p + ".cost_down"], _t(z[p + ".cost_down"]), **tol)
        assert torch.allclose(col[p + ".logits"], _t(z[p + ".logits"]), **tol)
        assert np.array_equal(col[p + ".class_map"].numpy().astype(np.int32), z[p + ".class_map"])
        assert torch.allclose(col[p + ".aug_down"], _t(z[p + ".aug_down"]), **tol)
        assert torch.allclose(col[p + ".fused"], _t(z[p + ".fused"]), **tol)
        assert torch.allclose(col[p + ".out"], _t(z[p + ".out"]), **tol)
    assert torch.allclose(col["classif3_logits"], _t(z["classif3_logits"]), **tol)
    assert torch.allclose(pv2, _t(z["prob_volume2"]), **tol)
    d = (pred4 - _t(z["pred4"])).abs()
    assert float(d.max()) < 1e-3 and float(d.mean()) < 1e-4, (float(d.max()), float(d.mean()))


def test_split_planes_carry_the_advertised_bits():
    torch.manual_seed(0)
    x = torch.randn(4096)
    hi, lo = O.split_bf16(x)
    rel = ((hi.float() + lo.float()) - x).abs() / x.abs().clamp_min(1e-20)
    assert float(rel.max()) < 2.0 ** -15
    hi, lo = O.split_h16(x, torch.float16)          # default plane format: 11 + 11 bits
    err = ((hi.float() + lo.float()) - x).abs()
    # relative 2^-21 where lo stays in the fp16 normal range (|x| >= 0.25), absolute 2^-23 below it
    assert float((err / x.abs().clamp_min(0.25)).max()) < 2.0 ** -21