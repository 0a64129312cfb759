"""Mirror of the reference's models/submodule.py API for the hot path (same names and argument
meaning), executed by the sm_100a kernels.  The 2-D front-end blocks (convbn, BasicBlock, BasicConv,
ResidualBlock, Guidance) are OUT OF SCOPE of the hot path: they stay
ordinary torch modules and exist so that reference checkpoints load with identical keys.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, engine


# ------------------------------------------------------------------ hot-path functions
def build_gwc_volume(refimg_fea, targetimg_fea, maxdisp, num_groups):
    """-> fp32 [B, num_groups, maxdisp, H, W]; reference models/submodule.py:157-167."""
    engine._require_cuda(refimg_fea, targetimg_fea)
    B, C, H, W = refimg_fea.shape
    assert C % num_groups == 0
    l, r = refimg_fea.contiguous().float(), targetimg_fea.contiguous().float()
    vol = torch.empty((B, num_groups, maxdisp, H, W), dtype=torch.float32, device=l.device)
    _lib.call("dca_build_gwc_volume_f32", l.data_ptr(), r.data_ptr(), vol.data_ptr(), B, C, num_groups, maxdisp, H, W,
              engine._stream())
    return vol


def build_concat_volume(refimg_fea, targetimg_fea, maxdisp):
    """-> fp32 [B, 2C, maxdisp, H, W]; reference models/submodule.py:134-145."""
    engine._require_cuda(refimg_fea, targetimg_fea)
    B, C, H, W = refimg_fea.shape
    l, r = refimg_fea.contiguous().float(), targetimg_fea.contiguous().float()
    vol = torch.empty((B, 2 * C, maxdisp, H, W), dtype=torch.float32, device=l.device)
    _lib.call("dca_build_concat_volume_f32", l.data_ptr(), r.data_ptr(), vol.data_ptr(), B, C, maxdisp, H, W,
              engine._stream())
    return vol


def build_cost_planes(gwc_l, gwc_r, cat_l, cat_r, maxdisp, num_groups, planes=2):
    """Fused gwc+concat volume in the kernels' own layout (what GwcNet.forward uses)."""
    return engine.fused_volume(gwc_l.contiguous().float(), gwc_r.contiguous().float(),
                               cat_l.contiguous().float() if cat_l is not None else None,
                               cat_r.contiguous().float() if cat_r is not None else None, maxdisp, num_groups, planes)


def disparity_regression(x, maxdisp):
    """x [B,D,H,W] -> sum_d d * x[:, d] as [B,1,H,W]; reference models/submodule.py:127-131 (x is NOT renormalised:
    un-normalised input gives the same un-normalised result as the reference).  Callers holding LOGITS should use
    softmax_disparity_regression (softmax fused, the probability volume never exists)."""
    assert len(x.shape) == 4
    assert x.shape[1] == maxdisp
    engine._require_cuda(x)
    return engine.regress(x.contiguous().float())


def softmax_disparity_regression(logits, maxdisp):
    """F.softmax(logits, 1) + disparity_regression fused (gwcnet_dca_g.py:238-239)."""
    assert len(logits.shape) == 4 and logits.shape[1] == maxdisp
    engine._require_cuda(logits)
    return engine.softmax_regress(logits.contiguous().float())


def convbn(in_channels, out_channels, kernel_size, stride, pad, dilation):
    return nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                   padding=dilation if dilation > 1 else pad, dilation=dilation, bias=False),
                         nn.BatchNorm2d(out_channels))


def convbn_3d(in_channels, out_channels, kernel_size, stride, pad):
    """Parameter container with the reference's key layout (`0.weight`, `1.{weight,bias,running_*}`);
    the arithmetic is done by engine.conv with BN folded into the epilogue."""
    return nn.Sequential(nn.Conv3d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=pad,
                                   bias=False),
                         nn.BatchNorm3d(out_channels))


def run_convbn_3d(seq, x, act=engine.ACT_NONE, planes=2):
    """Execute a convbn_3d container on an fp32 NCDHW tensor through the kernels (API-level helper)."""
    conv = seq[0]
    k, s = conv.kernel_size[0], conv.stride[0]
    mode = {(3, 1): engine.K3S1, (3, 2): engine.K3S2, (1, 1): engine.K1}[(k, s)]

    def build():
        pc = engine.pack_convbn(seq)
        pc.pack_tc(planes)
        return pc
    y = engine.conv(engine.Planes.from_ncdhw(x, planes), engine.cached_pack(seq, ("convbn", planes), build), mode, act)
    return y.to_ncdhw()


# ------------------------------------------------------------------ 2-D front end (torch, out of scope)
class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride, downsample, pad, dilation):
        super().__init__()
        self.conv1 = nn.Sequential(convbn(inplanes, planes, 3, stride, pad, dilation), nn.ReLU(inplace=True))
        self.conv2 = convbn(planes, planes, 3, 1, pad, dilation)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        if self.downsample is not None:
            x = self.downsample(x)
        return out + x


class BasicConv(nn.Module):
    def __init__(self, in_channels, out_channels, deconv=False, is_3d=False, bn=True, relu=True, **kwargs):
        super().__init__()
        self.relu, self.use_bn = relu, bn
        if is_3d:
            cls = nn.ConvTranspose3d if deconv else nn.Conv3d
            self.conv = cls(in_channels, out_channels, bias=False, **kwargs)
            self.bn = nn.BatchNorm3d(out_channels)
        else:
            cls = nn.ConvTranspose2d if deconv else nn.Conv2d
            self.conv = cls(in_channels, out_channels, bias=False, **kwargs)
            self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, x):
        x = self.conv(x)
        if self.use_bn:
            x = self.bn(x)
        return F.relu(x, inplace=True) if self.relu else x


class ResidualBlock(nn.Module):
    def __init__(self, in_planes, planes, norm_fn="group", stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1)
        self.relu = nn.ReLU(inplace=True)

        def norm():
            if norm_fn == "group":
                return nn.GroupNorm(num_groups=planes // 8, num_channels=planes)
            if norm_fn == "batch":
                return nn.BatchNorm2d(planes)
            if norm_fn == "instance":
                return nn.InstanceNorm2d(planes)
            return nn.Sequential()

        self.norm1, self.norm2 = norm(), norm()
        if stride != 1:
            self.norm3 = norm()
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride), self.norm3)
        else:
            self.downsample = None

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


class Guidance(nn.Module):
    """Left-image guidance net -> {'g': [B,64,H/4,W/4]} (reference models/submodule.py:395-460)."""

    def __init__(self, output_dim=64, norm_fn="batch"):
        super().__init__()
        self.norm_fn = norm_fn
        if norm_fn == "group":
            self.norm1 = nn.GroupNorm(num_groups=8, num_channels=32)
        elif norm_fn == "batch":
            self.norm1 = nn.BatchNorm2d(32)
        elif norm_fn == "instance":
            self.norm1 = nn.InstanceNorm2d(32)
        else:
            self.norm1 = nn.Sequential()
        self.conv_start = nn.Sequential(nn.Conv2d(3, 32, kernel_size=7, stride=2, padding=3), self.norm1,
                                        nn.ReLU(inplace=True))
        self.in_planes = 32
        self.layer1 = self._make_layer(32, stride=1)
        self.layer2 = self._make_layer(64, stride=2)
        self.conv_g0 = nn.Sequential(BasicConv(64, 64, kernel_size=3, padding=1),
                                     BasicConv(64, 64, kernel_size=3, padding=1))
        self.guidance = nn.Conv2d(64, output_dim, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)
        self.frontend_planes = 2

    def _make_layer(self, dim, stride=1):
        layers = (ResidualBlock(self.in_planes, dim, self.norm_fn, stride=stride),
                  ResidualBlock(dim, dim, self.norm_fn, stride=1))
        self.in_planes = dim
        return nn.Sequential(*layers)

    def forward(self, x):
        from . import frontend
        if frontend.use_kernels(self, x) and self.norm_fn == "batch":
            # CUDA + eval: layer2[1], conv_g0 and the output conv (all 64 -> 64 at 1/4 res) on the tcgen05 2-D conv kernel
            return frontend.guidance_forward(self, x, self.frontend_planes)
        x = self.layer2(self.layer1(self.conv_start(x)))
        return {"g": self.guidance(self.conv_g0(x))}


# ------------------------------------------------------------------ convex upsampling (hot-path tail)
class PropgationNet_4x(nn.Module):
    """RAFT-style convex 4x upsampling; reference models/gwcnet_dca_g.py:108-124."""

    def __init__(self, base_channels):
        super().__init__()
        self.base_channels = base_channels
        self.conv = nn.Sequential(convbn(base_channels, base_channels * 2, 3, 1, 1, 1),
                                  nn.ReLU(inplace=True),
                                  nn.Conv2d(base_channels * 2, 9 * 16, kernel_size=(3, 3), stride=(1, 1), padding=1,
                                            dilation=(1, 1), bias=False))
        self.precision_planes = 2

    def forward(self, guidance, disp):
        engine._require_cuda(guidance, disp)
        P = self.precision_planes
        gp = engine.Planes.from_ncdhw(guidance.contiguous().float(), planes=P)
        if engine.Options.use_tc and engine.Options.prop_on_tc and self.base_channels in (64, 128):
            # the two 3x3 Conv2d on the halo-slab tcgen05 kernel (the route GwcNet.forward takes)
            p0, p2 = engine.cached_pack(self, ("prop_tc", P), lambda: (
                engine.PackedConv2dTc(self.conv[0][0].weight, self.conv[0][1], P),
                engine.PackedConv2dTc(self.conv[2].weight, None, P)))
            m1 = engine.conv2d_tc(gp, p0, engine.ACT_RELU)
            mask = engine.conv2d_tc(m1, p2, engine.ACT_NONE, out_fp32=True)
        else:
            p0, p2 = engine.cached_pack(self, ("prop_direct", P), lambda: (
                engine.pack_convbn(self.conv[0]), engine.PackedConv(self.conv[2].weight)))
            m1 = engine.conv(gp, p0, engine.C2D3, engine.ACT_RELU)
            mask = engine.conv(m1, p2, engine.C2D3, engine.ACT_NONE, out_fp32=True)
        return engine.convex_upsample(mask, disp.contiguous().float())
