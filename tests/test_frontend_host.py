"""CPU-only host logic of the front end on the kernels (frontend.py), by dry run (tests/_dryrun.py: nothing is computed):
every Conv2d module of feature_extraction / Guidance maps to exactly one conv entry-point call, in module order, with the
shapes the layer has; the outputs leave through the NCHW slice writer."""
import collections

import torch

import _dryrun


def _trace(fn):
    import dcanet_b200 as d
    F = d.frontend
    saved = F.use_kernels
    F.use_kernels = lambda module, x: not module.training
    try:
        with _dryrun.recording() as trace:
            with torch.no_grad():
                out = fn()
            return list(trace), out
    finally:
        F.use_kernels = saved


CONV_CALLS = ("dca_conv2d_stem", "dca_conv2d_tc", "dca_conv2d_tc_ex", "dca_conv2d_tc_cat", "dca_conv3d_tc")


def test_every_conv2d_of_the_front_end_is_one_kernel_call():
    import dcanet_b200 as d
    net = d.GwcNet(48).eval()
    x = torch.zeros(2, 3, 64, 128)
    tr, out = _trace(lambda: net.feature_extraction(x))
    convs = [t for t in tr if t[0] in CONV_CALLS]
    n_mod = sum(isinstance(m, torch.nn.Conv2d) for m in net.feature_extraction.modules())
    assert len(convs) == n_mod == 57
    names = collections.Counter(t[0] for t in convs)
    assert names["dca_conv2d_stem"] == 1 and names["dca_conv3d_tc"] == 2 and names["dca_conv2d_tc_cat"] == 1
    assert out["gwc_feature"].shape == (2, 320, 16, 32) and out["concat_feature"].shape == (2, 12, 16, 32)
    slices = [t for t in tr if t[0] == "dca_planes_to_nchw_slice"]
    assert len(slices) == 4                          # l2, l3, l4 into gwc_feature; the concat feature
    # dilation 2 exactly on layer4's six convs (last integer argument before the stream placeholder)
    dil2 = [t for t in convs if t[0] == "dca_conv2d_tc_ex" and t[-2] == 2]
    assert len(dil2) == 6

    tr, out = _trace(lambda: net.guidance(x[:1]))
    convs = [t for t in tr if t[0] in CONV_CALLS]
    n_mod = sum(isinstance(m, torch.nn.Conv2d) for m in net.guidance.modules())
    assert len(convs) == n_mod == 13
    assert collections.Counter(t[0] for t in convs)["dca_conv2d_stem"] == 1
    assert out["g"].shape == (1, 64, 16, 32)


def test_front_end_keeps_torch_modules_for_cpu_inputs_and_training():
    """The kernel route is for CUDA inputs in eval mode; anything else takes the plain torch modules (the reference's own
    arithmetic), so CPU callers -- e.g. the golden-fixture generators -- keep working."""
    import dcanet_b200 as d
    net = d.GwcNet(48).eval()
    x = torch.randn(1, 3, 32, 64)
    with torch.no_grad():
        f = net.feature_extraction(x)
        g = net.guidance(x)["g"]
    assert f["gwc_feature"].shape == (1, 320, 8, 16) and g.shape == (1, 64, 8, 16)
    assert not d.frontend.use_kernels(net.feature_extraction, x)
    net.train()
    assert not d.frontend.use_kernels(net.feature_extraction, x)
