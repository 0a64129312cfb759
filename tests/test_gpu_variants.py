"""Stage-count variants on the GPU against fixtures made by the reference's own variant modules
(tests/golden/make_golden_variants.py).  Same kernels as the 3-stage model, N-stage host graph.  The unchanged
launch sequence of the default model is checked on CPU (tests/test_host_sequence.py)."""
import importlib

import pytest
import torch

from test_variants import VARIANTS, load_variant
from _util import _t

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("n", VARIANTS)
def test_variant_hot_path_matches_reference_fixture(n):
    mod = importlib.import_module(f"cost-volume-aggregation-in-stereo-matching-revisited_b200.gwcnet_dca{n}_g")
    z, sd, feats, maxdisp = load_variant(n)
    net = mod.GwcNet(maxdisp)
    own = net.state_dict()
    own.update({k: v for k, v in sd.items() if k in own})
    net.load_state_dict(own)
    net = net.cuda().eval()
    with torch.no_grad():
        pred4, pv = net.hot_path(*[f.cuda() for f in feats])
    torch.cuda.synchronize()
    err = (pred4.cpu() - _t(z["pred"]).reshape(pred4.shape)).abs()
    assert float(err.max()) <= 0.05 and float(err.mean()) <= 0.01          # north_star tolerance
    assert float((pv.cpu() - _t(z["pv"])).abs().max()) <= 1e-3
