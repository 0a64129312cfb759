"""The 1/4-resolution part of the 2-D front end on the halo-slab tcgen05 kernel.

`feature_extraction` (reference models/gwcnet_dca_g.py:13-66) and `Guidance` (models/submodule.py:395-460) produce the
hot path's inputs.  Their 1/4-resolution layers are 3x3 stride-1 convs with 64 / 128 / 320 input channels -- exactly the
shape family of `dca_conv2d_tc*` -- and hold 80 % of the front end's FLOPs:

  feature_extraction   layer2[1:]   15 BasicBlocks, 64 -> 64                      30 convs
                       layer3       3 BasicBlocks, 64 -> 128 -> 128 (+ 1x1 downsample)   6 + 1
                       layer4       3 BasicBlocks, 128 -> 128, dilation 2         6 convs (four parity sub-images each)
                       lastconv     320 -> 128 (+BN+ReLU) on cat(l2, l3, l4), 1x1 128 -> 12
  feature_extraction   firstconv[2,4], layer1 (3 BasicBlocks), 32 -> 32 at 1/2 resolution        8 convs
  Guidance             layer1       2 ResidualBlocks 32 -> 32 at 1/2 resolution              4 convs
                       layer2[1]    ResidualBlock 64 -> 64 (conv bias + BN + ReLU, relu(x + y))
                       conv_g0      2 x BasicConv 64 -> 64, guidance 64 -> 64

BN (eval) and conv biases are folded into the epilogue's fp32 scale/shift; residual adds (`out += x`, BasicBlock
submodule.py:272; `relu(x + y)`, ResidualBlock :347) are epilogue terms; the channel concat is never materialised in
the plane layout (the 320-channel conv reads its five 64-channel slabs through three tensor maps) and is written once,
as fp32 NCHW, for the volume kernel.  The 32-channel stride-1 layers of the 1/2-resolution stem (firstconv[2], firstconv[4],
feature_extraction.layer1, Guidance.layer1) run on the same kernel with a 32-channel tile; the stride-2 convs of the two
down-sampling blocks (3x3 32 -> 64 and the 1x1 shortcut) run on the 3-D stride-2 slab kernel at depth 1
(`engine.pack_conv2d_s2`); the two 3-channel stems (3x3 s2, 7x7 s2) on a small CUDA-core kernel (`dca_conv2d_stem`) that
reads the image and writes planes.  With CUDA inputs in eval mode no layer of the front end runs on torch / cuDNN.

torch here: device memory and load-time parameter folding.  (CPU inputs, training mode, odd sizes and
`Options.enabled = False` take the plain torch modules, in true fp32.)
"""
import torch

from . import _lib, engine as E


class Options:
    enabled = True       # CUDA + eval: run the 1/4-res layers on the tcgen05 kernels (False: plain torch modules)


def _conv_bn(seq):
    """convbn Sequential(Conv2d, BatchNorm2d) -> (conv, bn)."""
    return seq[0], seq[1]


class _Block:
    """BasicBlock (submodule.py:251-273): conv1+BN+ReLU, conv2+BN, + (downsample(x) | x).  No ReLU after the add."""

    def __init__(self, blk, planes):
        c1, b1 = _conv_bn(blk.conv1[0])
        c2, b2 = _conv_bn(blk.conv2)
        if c1.stride != (1, 1):
            raise _lib.DcaError("stride-2 BasicBlocks stay on torch")
        self.dil = int(c1.dilation[0])
        self.c1 = E.PackedConv2dTc(c1.weight, b1, planes)
        self.c2 = E.PackedConv2dTc(c2.weight, b2, planes)
        self.ds = None
        if blk.downsample is not None:
            self.ds = E.PackedConv2dTc(blk.downsample[0].weight, blk.downsample[1], planes)

    def __call__(self, x):
        h = E.conv2d_tc(x, self.c1, E.ACT_RELU, dil=self.dil)
        r = x if self.ds is None else E.conv2d_tc(x, self.ds, E.ACT_NONE)
        return E.conv2d_tc(h, self.c2, E.ACT_NONE, res=r, dil=self.dil)


class _ResBlock:
    """ResidualBlock with stride 1 (submodule.py:305-347): relu(x + relu(bn2(conv2(relu(bn1(conv1 x)))))), convs with bias."""

    def __init__(self, rb, planes):
        for n in (rb.norm1, rb.norm2):
            if not isinstance(n, torch.nn.BatchNorm2d):
                raise _lib.DcaError("Guidance on the tcgen05 kernels needs norm_fn='batch' (eval-mode BN folds into the epilogue)")
        if rb.downsample is not None:
            raise _lib.DcaError("stride-2 ResidualBlocks stay on torch")
        self.c1 = E.PackedConv2dTc(rb.conv1.weight, rb.norm1, planes, bias=rb.conv1.bias)
        self.c2 = E.PackedConv2dTc(rb.conv2.weight, rb.norm2, planes, bias=rb.conv2.bias)

    def __call__(self, x):
        y = E.conv2d_tc(x, self.c1, E.ACT_RELU)
        return E.conv2d_tc(y, self.c2, E.ACT_RELU, res=x, act_post=E.ACT_RELU)


def _planes_to_nchw(p, channels=None):
    C = channels or p.C
    out = torch.empty((p.B, C, p.H, p.W), dtype=torch.float32, device=p.t.device)
    E.planes_to_nchw_slice(p, out, 0, channels=C)
    return out


class _BlockS2:
    """The stride-2 BasicBlock feature_extraction.layer2[0] (32 -> 64; 1x1 stride-2 downsample on the shortcut): both strided
    convs on the 3-D stride-2 slab kernel at depth 1 (engine.pack_conv2d_s2), conv2 + the add on the 2-D kernel."""

    def __init__(self, blk, planes):
        c1, b1 = _conv_bn(blk.conv1[0])
        c2, b2 = _conv_bn(blk.conv2)
        self.c1 = E.pack_conv2d_s2(c1, b1, planes)
        self.ds = E.pack_conv2d_s2(blk.downsample[0], blk.downsample[1], planes)
        self.c2 = E.PackedConv2dTc(c2.weight, b2, planes)

    def __call__(self, x):
        h = E.conv(x, self.c1, E.K3S2, E.ACT_RELU)
        r = E.conv(x, self.ds, E.K3S2, E.ACT_NONE)
        return E.conv2d_tc(h, self.c2, E.ACT_NONE, res=r)


class _ResBlockS2:
    """The stride-2 ResidualBlock Guidance.layer2[0] (submodule.py:305-347): relu(norm3(down(x)) + relu(bn2(conv2(relu(bn1(conv1 x))))))."""

    def __init__(self, rb, planes):
        for n in (rb.norm1, rb.norm2, rb.norm3):
            if not isinstance(n, torch.nn.BatchNorm2d):
                raise _lib.DcaError("Guidance on the tcgen05 kernels needs norm_fn='batch'")
        self.c1 = E.pack_conv2d_s2(rb.conv1, rb.norm1, planes)
        self.ds = E.pack_conv2d_s2(rb.downsample[0], rb.norm3, planes)
        self.c2 = E.PackedConv2dTc(rb.conv2.weight, rb.norm2, planes, bias=rb.conv2.bias)

    def __call__(self, x):
        y = E.conv(x, self.c1, E.K3S2, E.ACT_RELU)
        r = E.conv(x, self.ds, E.K3S2, E.ACT_NONE)
        return E.conv2d_tc(y, self.c2, E.ACT_RELU, res=r, act_post=E.ACT_RELU)


class PackedFeatureExtraction:
    def __init__(self, fe, planes):
        self.planes = planes
        self.stem = E.PackedStem(fe.firstconv[0][0], fe.firstconv[0][1])       # 3 -> 32, 3x3 stride 2 (+BN+ReLU)
        self.down = _BlockS2(fe.layer2[0], planes)
        # 1/2-resolution stem: firstconv's two 32 -> 32 convs and layer1 (3 BasicBlocks) on the 32-channel tile
        self.first = [E.PackedConv2dTc(fe.firstconv[i][0].weight, fe.firstconv[i][1], planes) for i in (2, 4)]
        self.layer1 = [_Block(b, planes) for b in fe.layer1]
        self.layer2 = [_Block(b, planes) for b in list(fe.layer2)[1:]]
        self.layer3 = [_Block(b, planes) for b in fe.layer3]
        self.layer4 = [_Block(b, planes) for b in fe.layer4]
        self.last0 = self.last2 = None
        if fe.concat_feature:
            self.last0 = E.PackedConv2dTc(fe.lastconv[0][0].weight, fe.lastconv[0][1], planes)
            self.last2 = E.PackedConv2dTc(fe.lastconv[2].weight, None, planes, pad_cout=True)
            self.cc = fe.lastconv[2].weight.shape[0]


class _no_tf32:
    """The stem runs on cuDNN in true fp32 (TF32's 10-bit mantissa moves the disparity by more than the 0.05 px the hot
    path is held to)."""

    def __enter__(self):
        self.a, self.b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.a, self.b


def feature_extraction_forward(fe, x, planes=2):
    """feature_extraction.forward (gwcnet_dca_g.py:53-66) for a CUDA batch in eval mode."""
    pk = E.cached_pack(fe, ("frontend", planes), lambda: PackedFeatureExtraction(fe, planes))
    p = E.conv2d_stem(x, pk.stem, planes, E.ACT_RELU)         # image -> 32-channel planes at 1/2 resolution
    for pc in pk.first:
        p = E.conv2d_tc(p, pc, E.ACT_RELU)
    for blk in pk.layer1:
        p = blk(p)
    p = pk.down(p)                                            # the stride-2 block: 64 channels at 1/4 resolution
    for blk in pk.layer2:
        p = blk(p)
    l2 = p
    for blk in pk.layer3:
        p = blk(p)
    l3 = p
    for blk in pk.layer4:
        p = blk(p)
    l4 = p
    B, H, W = l2.B, l2.H, l2.W
    gwc = torch.empty((B, 320, H, W), dtype=torch.float32, device=x.device)
    E.planes_to_nchw_slice(l2, gwc, 0)
    E.planes_to_nchw_slice(l3, gwc, 64)
    E.planes_to_nchw_slice(l4, gwc, 192)
    if pk.last0 is None:
        return {"gwc_feature": gwc}
    h = E.conv2d_tc_cat((l2, l3, l4), pk.last0, E.ACT_RELU)
    c = E.conv2d_tc(h, pk.last2, E.ACT_NONE)                  # 128 -> 12, stored in a 64-channel plane row
    cat = torch.empty((B, pk.cc, H, W), dtype=torch.float32, device=x.device)
    E.planes_to_nchw_slice(c, cat, 0, channels=pk.cc)
    return {"gwc_feature": gwc, "concat_feature": cat}


class PackedGuidance:
    def __init__(self, g, planes):
        if not isinstance(g.conv_start[1], torch.nn.BatchNorm2d):
            raise _lib.DcaError("Guidance on the tcgen05 kernels needs norm_fn='batch'")
        self.stem = E.PackedStem(g.conv_start[0], g.conv_start[1])        # 3 -> 32, 7x7 stride 2 (+bias+BN+ReLU)
        self.layer1 = [_ResBlock(rb, planes) for rb in g.layer1]          # 32 channels at 1/2 resolution
        self.down = _ResBlockS2(g.layer2[0], planes)
        self.rb = _ResBlock(g.layer2[1], planes)
        self.g0 = E.PackedConv2dTc(g.conv_g0[0].conv.weight, g.conv_g0[0].bn, planes)
        self.g1 = E.PackedConv2dTc(g.conv_g0[1].conv.weight, g.conv_g0[1].bn, planes)
        self.out = E.PackedConv2dTc(g.guidance.weight, None, planes, bias=g.guidance.bias)
        self.cout = g.guidance.weight.shape[0]


def guidance_forward(g, x, planes=2):
    """Guidance.forward (submodule.py:452-460) for a CUDA batch in eval mode -> {'g': [B,64,H/4,W/4]}."""
    pk = E.cached_pack(g, ("frontend", planes), lambda: PackedGuidance(g, planes))
    p = E.conv2d_stem(x, pk.stem, planes, E.ACT_RELU)
    for rb in pk.layer1:
        p = rb(p)
    p = pk.rb(pk.down(p))
    p = E.conv2d_tc(p, pk.g0, E.ACT_RELU)
    p = E.conv2d_tc(p, pk.g1, E.ACT_RELU)
    p = E.conv2d_tc(p, pk.out, E.ACT_NONE)
    return {"g": _planes_to_nchw(p, pk.cout)}


def use_kernels(module, x):
    return Options.enabled and x.is_cuda and not module.training and x.dtype == torch.float32
