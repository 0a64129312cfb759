"""The 2-D front end (feature_extraction, Guidance) on this repo's kernels (frontend.py) against the same
torch modules on cuDNN in true fp32 (TF32 off), against an fp64 run, and against torch on the CPU (op level)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mods():
    import dcanet_b200 as d
    return d, d.engine


def _planes_to_nchw(E, y):
    out = torch.empty((y.B, y.C, y.H, y.W), dtype=torch.float32, device="cuda")
    E.planes_to_nchw_slice(y, out, 0)
    return out


@pytest.mark.parametrize("B,Cin,Cout,H,W,dil,bias,bn,res,act_post", [
    (1, 64, 64, 20, 28, 1, False, True, True, 0),      # BasicBlock.conv2 + residual
    (2, 64, 128, 17, 23, 1, False, True, False, 0),    # layer3[0].conv1, odd sizes, two output chunks
    (1, 128, 128, 24, 40, 2, False, True, True, 0),    # layer4: dilation 2 through the four parity sub-images
    (2, 128, 128, 10, 52, 2, False, True, False, 0),
    (1, 64, 64, 19, 33, 1, True, True, True, 1),       # ResidualBlock.conv2: bias + BN + ReLU, relu(x + y)
    (1, 64, 64, 16, 24, 1, True, False, False, 0),     # conv with bias, no BN
    (2, 32, 32, 30, 44, 1, False, True, True, 0),      # the 1/2-res stem: 32-channel tile (firstconv, layer1)
    (1, 32, 32, 17, 25, 1, True, True, True, 1),       # Guidance.layer1 ResidualBlock at 32 channels
    (1, 32, 64, 12, 20, 1, False, True, False, 0),     # 32 -> 64: two 32-channel output chunks
])
def test_conv2d_tc_ex_matches_torch(B, Cin, Cout, H, W, dil, bias, bn, res, act_post):
    d, E = _mods()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5
    bvec = torch.randn(Cout, generator=g) * 0.1 if bias else None
    norm = None
    if bn:
        norm = torch.nn.BatchNorm2d(Cout).eval()
        norm.weight.data.uniform_(0.8, 1.2, generator=g); norm.bias.data.normal_(0, 0.1, generator=g)
        norm.running_mean.normal_(0, 0.1, generator=g); norm.running_var.uniform_(0.5, 1.5, generator=g)
    r = torch.randn(B, Cout, H, W, generator=g) if res else None
    with torch.no_grad():
        ref = F.conv2d(x, w, bvec, padding=dil, dilation=dil)
        if norm is not None:
            ref = norm(ref)
        ref = torch.relu(ref)
        if r is not None:
            ref = ref + r
            if act_post:
                ref = torch.relu(ref)
    pc = E.PackedConv2dTc(w.cuda(), norm.cuda() if norm is not None else None, 2, bias=None if bvec is None else bvec.cuda())
    xp = E.Planes.from_ncdhw(x.cuda(), 2)
    rp = E.Planes.from_ncdhw(r.cuda(), 2) if r is not None else None
    y = E.conv2d_tc(xp, pc, E.ACT_RELU, res=rp, act_post=act_post, dil=dil)
    got = _planes_to_nchw(E, y).cpu()
    err = float((got - ref).abs().max())
    assert err <= 3e-5 * max(1.0, float(ref.abs().max())), err


def test_conv2d_tc_cat_and_1x1_match_torch():
    """lastconv: 3x3 over cat(64, 128, 128 channels) without materialising the concat, then the 1x1 128 -> 12 embedded as a
    centre tap with zero-padded output channels."""
    d, E = _mods()
    g = torch.Generator().manual_seed(4)
    B, H, W = 2, 21, 30
    xs = [torch.randn(B, c, H, W, generator=g) for c in (64, 128, 128)]
    w = torch.randn(128, 320, 3, 3, generator=g) * (2.0 / (9 * 320)) ** 0.5
    w1 = torch.randn(12, 128, 1, 1, generator=g) * (1.0 / 128) ** 0.5
    with torch.no_grad():
        h = torch.relu(F.conv2d(torch.cat(xs, 1), w, padding=1))
        ref = F.conv2d(h, w1)
    pc = E.PackedConv2dTc(w.cuda(), None, 2)
    pc1 = E.PackedConv2dTc(w1.cuda(), None, 2, pad_cout=True)
    hp = E.conv2d_tc_cat([E.Planes.from_ncdhw(x.cuda(), 2) for x in xs], pc, E.ACT_RELU)
    assert float((_planes_to_nchw(E, hp).cpu() - h).abs().max()) <= 3e-5 * float(h.abs().max())
    c = E.conv2d_tc(hp, pc1, E.ACT_NONE)
    out = torch.zeros((B, 20, H, W), dtype=torch.float32, device="cuda")
    E.planes_to_nchw_slice(c, out, 5, channels=12)                      # a channel slice of a wider NCHW tensor
    assert float((out[:, 5:17].cpu() - ref).abs().max()) <= 3e-5 * float(ref.abs().max())
    assert float(out[:, :5].abs().max()) == 0.0 and float(out[:, 17:].abs().max()) == 0.0


@pytest.mark.parametrize("K,bias,H,W", [(3, False, 36, 52), (7, True, 40, 64), (7, True, 33, 47), (3, False, 31, 45)])
def test_conv2d_stem_matches_torch(K, bias, H, W):
    """Conv2d(3 -> 32, K x K, stride 2, pad K/2) (+bias) + BN + ReLU: firstconv[0] (3x3) and Guidance.conv_start (7x7)."""
    d, E = _mods()
    g = torch.Generator().manual_seed(K)
    conv = torch.nn.Conv2d(3, 32, K, 2, K // 2, bias=bias)
    conv.weight.data.normal_(0, (2.0 / (3 * K * K)) ** 0.5, generator=g)
    bn = torch.nn.BatchNorm2d(32).eval()
    bn.weight.data.uniform_(0.8, 1.2, generator=g); bn.bias.data.normal_(0, 0.1, generator=g)
    bn.running_mean.normal_(0, 0.1, generator=g); bn.running_var.uniform_(0.5, 1.5, generator=g)
    x = torch.randn(2, 3, H, W, generator=g)
    with torch.no_grad():
        ref = torch.relu(bn(conv(x)))
    y = E.conv2d_stem(x.cuda(), E.PackedStem(conv.cuda(), bn.cuda()), 2, E.ACT_RELU)
    got = _planes_to_nchw(E, y).cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("k,bias,H,W", [(3, False, 24, 40), (1, False, 24, 40), (3, True, 18, 36), (1, True, 30, 28)])
def test_stride2_conv2d_on_the_3d_slab_kernel_matches_torch(k, bias, H, W):
    """Conv2d(32 -> 64, 3x3 pad 1 or 1x1, stride 2) + BN as the middle depth slice of a K3S2 Conv3d on a depth-1 volume."""
    d, E = _mods()
    g = torch.Generator().manual_seed(10 + k)
    conv = torch.nn.Conv2d(32, 64, k, 2, k // 2, bias=bias)
    conv.weight.data.normal_(0, (2.0 / (32 * k * k)) ** 0.5, generator=g)
    bn = torch.nn.BatchNorm2d(64).eval()
    bn.weight.data.uniform_(0.8, 1.2, generator=g); bn.bias.data.normal_(0, 0.1, generator=g)
    bn.running_mean.normal_(0, 0.1, generator=g); bn.running_var.uniform_(0.5, 1.5, generator=g)
    x = torch.randn(2, 32, H, W, generator=g)
    with torch.no_grad():
        ref = torch.relu(bn(conv(x)))
    pc = E.pack_conv2d_s2(conv.cuda(), bn.cuda(), 2)
    y = E.conv(E.Planes.from_ncdhw(x.cuda(), 2), pc, E.K3S2, E.ACT_RELU)
    got = _planes_to_nchw(E, y).cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 3e-5 * max(1.0, float(ref.abs().max()))


def _net(maxdisp=48, seed=0):
    import dcanet_b200 as d
    import workloads
    net = workloads.init_bench_weights_(d.GwcNet(maxdisp), seed)
    g = torch.Generator().manual_seed(seed + 1)
    for m in net.modules():                      # non-trivial BN statistics and conv biases in the front end
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1, generator=g)
            m.running_var.uniform_(0.7, 1.3, generator=g)
        if isinstance(m, torch.nn.Conv2d) and m.bias is not None:
            m.bias.data.normal_(0, 0.05, generator=g)
    return net.cuda().eval()


def _images(B, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    lo = torch.randn(B, 3, H // 8, W // 8, generator=g)
    left = F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False) + 0.1 * torch.randn(B, 3, H, W, generator=g)
    right = torch.roll(left, -6, dims=3) + 0.05 * torch.randn(B, 3, H, W, generator=g)
    return left.cuda(), right.cuda()


@pytest.mark.parametrize("B,H,W", [(1, 64, 128), (2, 96, 160)])
def test_front_end_kernels_match_the_torch_modules(B, H, W):
    d, E = _mods()
    net = _net()
    left, _ = _images(B, H, W)
    with torch.no_grad():
        d.frontend.Options.enabled = False
        try:
            with d.frontend._no_tf32():
                ref = net.feature_extraction(left)
                refg = net.guidance(left)["g"]
        finally:
            d.frontend.Options.enabled = True
        n0 = d._lib.LAUNCHES
        real = torch.nn.Conv2d.forward
        try:                                      # CUDA + eval: NO layer of the front end may run on torch / cuDNN
            torch.nn.Conv2d.forward = lambda self, x: (_ for _ in ()).throw(AssertionError("a torch Conv2d ran"))
            got = net.feature_extraction(left)
            gotg = net.guidance(left)["g"]
        finally:
            torch.nn.Conv2d.forward = real
    assert d._lib.LAUNCHES - n0 > 60, "the kernel path did not run"
    for name, a, b in (("gwc_feature", got["gwc_feature"], ref["gwc_feature"]),
                       ("concat_feature", got["concat_feature"], ref["concat_feature"]), ("g", gotg, refg)):
        assert a.shape == b.shape and a.is_contiguous()
        err = float((a - b).abs().max()) / float(b.abs().max())
        print(f"{name}: max rel err {err:.2e}")
        assert err <= 2e-5, (name, err)


def test_images_to_disparity_with_kernel_front_end():
    """GwcNet(left, right) with the front end's 1/4-res layers on the kernels.  The un-normalised random checkpoint makes
    the disparity hypersensitive to the feature maps (fp32 re-association alone moves single pixels by ~1 px), so the
    yardstick is the cuDNN fp32 front end: both are compared with the SAME hot path fed by an fp64 CPU run of the front
    end, and the kernel front end must sit as close to it as cuDNN fp32 does (benchmarks/frontend_probe.py)."""
    import copy
    d, E = _mods()
    net = _net(maxdisp=96)
    left, right = _images(1, 128, 256, seed=2)
    with torch.no_grad():
        x = torch.cat((left, right))
        f64 = copy.deepcopy(net.feature_extraction).double().cpu()(x.double().cpu())
        g64 = copy.deepcopy(net.guidance).double().cpu()(left.double().cpu())["g"]
        gw, cc = f64["gwc_feature"].float().cuda(), f64["concat_feature"].float().cuda()
        truth, _ = net.hot_path(gw[:1].contiguous(), gw[1:].contiguous(), cc[:1].contiguous(), cc[1:].contiguous(),
                                g64.float().cuda())
        d.frontend.Options.enabled = False
        try:
            ref4, _ = net(left, right)
        finally:
            d.frontend.Options.enabled = True
        n0 = d._lib.LAUNCHES
        pred4, pv = net(left, right)
        assert d._lib.LAUNCHES - n0 > 100
    dk, dc = (pred4 - truth).abs(), (ref4 - truth).abs()
    fk, fc = float((dk > 0.05).float().mean()), float((dc > 0.05).float().mean())
    print("vs fp64 front end: kernel FE mean %.5f frac>0.05 %.4f | cuDNN fp32 FE mean %.5f frac>0.05 %.4f"
          % (float(dk.mean()), fk, float(dc.mean()), fc))
    assert pred4.shape == (1, 1, 128, 256) and pv.shape == (1, 12, 16, 32)
    assert float(dk.mean()) <= 1.5 * float(dc.mean()) + 1e-3 and fk <= 1.5 * fc + 2e-3
