"""Synthetic workloads for bench.py / smoke(): shapes of BASELINE.json's configs, seeded feature maps and
a well-conditioned random checkpoint.  No arithmetic of the hot path lives here."""
import math

import torch

CONFIGS = {
    # name: (H, W, maxdisp, batch)           BASELINE.json configs[...]
    "parity_256x512": (256, 512, 192, 1),     # configs[0]
    "kitti_384x1248": (384, 1248, 192, 1),    # configs[1]  <- the metric's config
    "sceneflow_544x960": (544, 960, 192, 1),  # configs[2] (batch 64 = 64 steps of this)
    "middlebury_1536x2048": (1536, 2048, 384, 1),   # configs[4], runs un-sharded on one 180 GB B200
    "tiny_64x128": (64, 128, 48, 1),
}


def init_bench_weights_(model, seed=0):
    """Variance-preserving random init: conv ~ N(0, 2/fan_in) (ReLU gain), BN = identity with small affine
    noise, so activations stay O(1) through the 60-odd layers without a calibration pass."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (torch.nn.Conv3d, torch.nn.Conv2d, torch.nn.ConvTranspose3d)):
            w = m.weight
            if isinstance(m, torch.nn.ConvTranspose3d):
                fan_in = w.shape[0] * w[0, 0].numel() / 8.0
            else:
                fan_in = w.shape[1] * w[0, 0].numel()
            w.data.copy_(torch.randn(w.shape, generator=g) * math.sqrt(1.0 / fan_in))
        elif isinstance(m, (torch.nn.BatchNorm3d, torch.nn.BatchNorm2d)):
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) * 0.2 + 0.9)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
            m.running_mean.zero_()
            m.running_var.fill_(1.0)
    return model


def feature_maps(seed, B, H4, W4, C=320, Cc=12, Cg=64, shift=3, device="cpu", pin=False):
    """Smooth random 1/4-res feature maps (left, right = left shifted + noise), concat features, guidance."""
    g = torch.Generator().manual_seed(seed)

    def smooth(c):
        lo = torch.randn(B, c, max(H4 // 4, 1), max(W4 // 4, 1), generator=g)
        x = torch.nn.functional.interpolate(lo, size=(H4, W4), mode="bilinear", align_corners=False)
        return x + 0.1 * torch.randn(B, c, H4, W4, generator=g)

    gl = torch.relu(smooth(C))
    gr = torch.roll(gl, -shift, dims=3) + 0.05 * torch.randn(B, C, H4, W4, generator=g)
    cl = 0.5 * smooth(Cc)
    cr = torch.roll(cl, -shift, dims=3) + 0.02 * torch.randn(B, Cc, H4, W4, generator=g)
    gd = smooth(Cg)
    out = [t.contiguous() for t in (gl, gr, cl, cr, gd)]
    if pin:
        out = [t.pin_memory() for t in out]
    if device != "cpu":
        out = [t.to(device) for t in out]
    return out


def hot_path_flops(H, W, maxdisp):
    """Algorithmic FLOPs of the hot path per pair: 2*Cin*Cout*taps*N_out per conv."""
    n4 = (maxdisp // 4) * (H // 4) * (W // 4)
    n8 = n4 // 8
    k = 27
    f = 2 * 64 * 32 * k * n4 + 3 * 2 * 32 * 32 * k * n4          # dres0, dres1
    per_cva = (2 * 32 * 32 * k * n8) * 2 + 2 * 32 * k * n8        # downsample, classify.0, classify.2
    per_cva += 6 * 2 * 32 * 32 * n8 + 4 * (H // 8) * (W // 8) * (maxdisp // 8) ** 2 * 8 * 2 * 2   # projections + attention
    per_cva += 2 * 64 * 32 * n4                                    # fuse
    per_cva += 2 * 32 * 64 * k * n8 + 2 * 64 * 64 * k * n8 + 2 * 64 * 32 * k * n8 + 2 * 32 * 32 * n4
    f += 3 * per_cva
    f += 2 * 32 * 32 * k * n4 + 2 * 32 * k * n4                    # classif3
    return f


def conv_k3s1_flops(H, W, maxdisp, cin=32, cout=32):
    return 2 * cin * cout * 27 * (maxdisp // 4) * (H // 4) * (W // 4)


def volume_bytes(H, W, maxdisp, planes=2, C=320, Cc=12, Cv=64):
    """Algorithmic HBM bytes of the fused volume kernel: fp32 features in, bf16 planes out."""
    h4, w4, d4 = H // 4, W // 4, maxdisp // 4
    return 2 * (C + Cc) * h4 * w4 * 4 + planes * Cv * d4 * h4 * w4 * 2
