"""CPU oracle for the DCANet cost-volume hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 torch-on-CPU functional calls and closed forms, what the
reference computes between the 1/4-resolution feature maps and the final disparity.  It is the
checker for the CUDA path; nothing under the product package imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / ``--impl reference``) may use it.

Pinning: the reference has NO tests / golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, produced in the build container by
``tests/golden/make_golden.py`` (imports /root/reference, dumps every boundary tensor to
``tests/golden/*.npz``).  ``tests/test_oracle_golden.py`` replays the fixtures.

The arithmetic of the reference is PyTorch ATen (un-vendored third party; README pins "Pytorch
1.6.0", this image has 2.11.0): conv3d, conv_transpose3d, batch_norm (eps 1e-5), avg_pool3d
(count_include_pad=True), upsample_trilinear3d (align_corners=False), softmax, matmul.  The dense
ops below call the same ATen entry points with the weights taken from a reference-layout
state_dict; everything that the reference expresses as Python loops / permutes / boolean indexing
is restated as a closed form.

Reference citations (relative to /root/reference):
  build_gwc_volume         models/submodule.py:148-167
  build_concat_volume      models/submodule.py:134-145
  convbn_3d                models/submodule.py:121-124
  disparity_regression     models/submodule.py:127-131
  cva / Multi_Aggregation  models/augment/cva.py:13-72
  SemanticLevelContext     models/augment/semantic_level.py:96-128
  SelfAttentionBlock       models/augment/SelfAttention_bn.py:62-98, 136-160
  PropgationNet_4x         models/gwcnet_dca_g.py:108-124
  GwcNet.forward (eval)    models/gwcnet_dca_g.py:209-240, 282
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
HEAD_DIM = 8  # SelfAttention_bn.py:64


# --------------------------------------------------------------------------------------------
# operand rounding (emulates reduced-mantissa conv operands; SURVEY section 7 hard part 2)
# --------------------------------------------------------------------------------------------
def round_mantissa(x: torch.Tensor, bits: Optional[int]) -> torch.Tensor:
    """Round fp32 to `bits` explicit mantissa bits (round-to-nearest-even). 7 == bf16."""
    if bits is None or bits >= 23:
        return x
    if bits == 7:
        return x.to(torch.bfloat16).to(torch.float32)
    if bits == 10:      # IEEE fp16 (the default 16-bit plane format of the CUDA path), including its exponent range
        return x.to(torch.float16).to(torch.float32)
    drop = 23 - bits
    i = x.contiguous().view(torch.int32)
    bias = ((i >> drop) & 1) + ((1 << (drop - 1)) - 1)
    return (((i + bias) >> drop) << drop).view(torch.float32)


def split_h16(x: torch.Tensor, dtype=torch.float16):
    """hi = h16(x), lo = h16(x - hi): what the CUDA path stores for every activation/weight (fp16 planes by default,
    bf16 when the library is built with DCA_F16_PLANES=0)."""
    hi = x.to(dtype)
    lo = (x - hi.to(torch.float32)).to(dtype)
    return hi, lo


def split_bf16(x: torch.Tensor):
    return split_h16(x, torch.bfloat16)


# --------------------------------------------------------------------------------------------
# volume construction
# --------------------------------------------------------------------------------------------
def build_gwc_volume(ref_fea, tgt_fea, maxdisp, num_groups):
    """vol[b,g,d,h,w] = mean_c L[b,g*cpg+c,h,w] * R[b,g*cpg+c,h,w-d] for w>=d else 0.
    submodule.py:157-167 (loop over d with slice-assign) and :148-154 (group mean)."""
    B, C, H, W = ref_fea.shape
    assert C % num_groups == 0
    cpg = C // num_groups
    vol = ref_fea.new_zeros(B, num_groups, maxdisp, H, W)
    for d in range(min(maxdisp, W)):
        prod = ref_fea[:, :, :, d:] * tgt_fea[:, :, :, : W - d]
        vol[:, :, d, :, d:] = prod.view(B, num_groups, cpg, H, W - d).mean(dim=2)
    return vol


def build_concat_volume(ref_fea, tgt_fea, maxdisp):
    """vol[b,c,d,h,w]=L[b,c,h,w], vol[b,C+c,d,h,w]=R[b,c,h,w-d] for w>=d else 0 (BOTH halves are
    zero for w<d: the reference slice-assigns only columns d: of a zeroed tensor).
    submodule.py:134-145."""
    B, C, H, W = ref_fea.shape
    vol = ref_fea.new_zeros(B, 2 * C, maxdisp, H, W)
    for d in range(min(maxdisp, W)):
        vol[:, :C, d, :, d:] = ref_fea[:, :, :, d:]
        vol[:, C:, d, :, d:] = tgt_fea[:, :, :, : W - d]
    return vol


def disparity_regression(prob, maxdisp):
    """sum_d d * prob[:, d]  -> [B,1,H,W].  submodule.py:127-131."""
    assert prob.dim() == 4
    d = torch.arange(0, maxdisp, dtype=prob.dtype, device=prob.device).view(1, maxdisp, 1, 1)
    return torch.sum(prob * d, 1, keepdim=True)


# --------------------------------------------------------------------------------------------
# dense building blocks on a reference-layout state_dict
# --------------------------------------------------------------------------------------------
class _Ctx:
    """state_dict accessor + optional operand rounding + optional BN calibration."""

    def __init__(self, sd: Dict[str, torch.Tensor], operand_bits=None, calibrate=False):
        self.sd = sd
        self.bits = operand_bits
        self.calibrate = calibrate

    def w(self, key):
        return self.sd[key]

    def bn(self, x, prefix):
        """eval-mode BatchNorm: (x-mean)/sqrt(var+eps)*gamma+beta.  In calibrate mode the batch
        statistics of x are first written into running_mean/var (momentum=1.0 semantics of
        SURVEY section 8c; torch stores the UNBIASED variance in running_var)."""
        dims = [0] + list(range(2, x.dim()))
        if self.calibrate:
            mean = x.mean(dim=dims)
            n = x.numel() // x.shape[1]
            var_b = x.var(dim=dims, unbiased=False)
            self.sd[prefix + ".running_mean"] = mean.clone()
            self.sd[prefix + ".running_var"] = (var_b * (n / max(n - 1, 1))).clone()
            # train-mode BN normalises with the biased batch variance
            shape = [1, -1] + [1] * (x.dim() - 2)
            return ((x - mean.view(shape)) / torch.sqrt(var_b.view(shape) + BN_EPS)
                    * self.sd[prefix + ".weight"].view(shape) + self.sd[prefix + ".bias"].view(shape))
        return F.batch_norm(x, self.sd[prefix + ".running_mean"], self.sd[prefix + ".running_var"],
                            self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], False, 0.0, BN_EPS)

    def conv3d(self, x, key, stride=1, pad=1):
        return F.conv3d(round_mantissa(x, self.bits), round_mantissa(self.sd[key], self.bits),
                        None, stride, pad)

    def deconv3d(self, x, key):
        return F.conv_transpose3d(round_mantissa(x, self.bits), round_mantissa(self.sd[key], self.bits),
                                  None, stride=2, padding=1, output_padding=1)

    def conv2d(self, x, key, pad=1):
        return F.conv2d(round_mantissa(x, self.bits), round_mantissa(self.sd[key], self.bits),
                        None, 1, pad)

    def convbn3d(self, x, prefix, stride=1, pad=1, act=None):
        """convbn_3d = Sequential(Conv3d(bias=False), BatchNorm3d)  submodule.py:121-124."""
        y = self.bn(self.conv3d(x, prefix + ".0.weight", stride, pad), prefix + ".1")
        if act == "relu":
            y = F.relu(y)
        elif act == "leaky":
            y = F.leaky_relu(y, 0.1)
        return y


def _project(ctx: _Ctx, x, prefix, num_convs):
    """SelfAttentionBlock.buildproject: num_convs x (Conv3d 1x1x1 + BN + LeakyReLU(0.1)).
    With num_convs == 1 the Sequential(conv,bn,act) is returned bare, so keys are `prefix.0/.1`;
    with 2 they are `prefix.{0,1}.{0,1}`.  SelfAttention_bn.py:136-160."""
    if num_convs == 1:
        return ctx.convbn3d(x, prefix, 1, 0, "leaky")
    for i in range(num_convs):
        x = ctx.convbn3d(x, f"{prefix}.{i}", 1, 0, "leaky")
    return x


def class_stats(logits):
    """Closed form of the SemanticLevelContext class statistics (semantic_level.py:98-119).
    logits [B,D,H,W] -> P=softmax_d, k_p=argmax_d P (first index on ties), e_p=exp(P[p,k_p]),
    S[b,k]=sum_{p:k_p=k} e_p, w_p=e_p/S[b,k_p]."""
    B, D, H, W = logits.shape
    P = F.softmax(logits, dim=1)
    k = P.argmax(dim=1)                                   # [B,H,W] int64
    pk = P.gather(1, k.unsqueeze(1)).squeeze(1)           # [B,H,W]
    e = torch.exp(pk)
    S = torch.zeros(B, D, dtype=logits.dtype)
    S.scatter_add_(1, k.view(B, -1), e.view(B, -1))
    w = e / S.gather(1, k.view(B, -1)).view(B, H, W)
    return P, k, e, S, w


def semantic_level_key(x, logits):
    """key_feats = x + feats_sl, with feats_sl non-zero only at plane d == k_p where it is
    x * w_p (semantic_level.py:111-126: per-class softmax over the PIXELS of that class, not
    summed)."""
    _, k, _, _, w = class_stats(logits)
    B, C, D, H, W = x.shape
    onehot = F.one_hot(k, D).permute(0, 3, 1, 2).to(x.dtype)          # [B,D,H,W]
    scale = 1.0 + onehot * w.unsqueeze(1)
    return x * scale.unsqueeze(1), k


def disparity_attention(ctx: _Ctx, prefix, query_feats, key_feats):
    """SelfAttentionBlock.forward (SelfAttention_bn.py:62-98): per pixel, per head (4 x 8 ch),
    softmax over the key-disparity axis of q k^T / sqrt(8), times v; then out_project."""
    B, C, D, H, W = query_feats.shape
    nh = C // HEAD_DIM
    q = _project(ctx, query_feats, prefix + ".query_project", 2)
    k = _project(ctx, key_feats, prefix + ".key_project", 2)
    v = _project(ctx, key_feats, prefix + ".value_project", 1)
    q = q.reshape(B, nh, HEAD_DIM, D, H * W)
    k = k.reshape(B, nh, HEAD_DIM, D, H * W)
    v = v.reshape(B, nh, HEAD_DIM, D, H * W)
    sim = torch.einsum("bncip,bncjp->bnpij", q, k) * (HEAD_DIM ** -0.5)
    sim = F.softmax(sim, dim=-1)
    out = torch.einsum("bnpij,bncjp->bncip", sim, v).reshape(B, C, D, H, W)
    return _project(ctx, out, prefix + ".out_project", 1)


def multi_aggregation(ctx: _Ctx, prefix, x):
    """Multi_Aggregation.forward  cva.py:13-31."""
    c1 = ctx.convbn3d(x, prefix + ".conv1.0", 2, 1, "relu")
    c2 = ctx.convbn3d(c1, prefix + ".conv2.0", 1, 1, "relu")
    c3 = ctx.bn(ctx.deconv3d(c2, prefix + ".conv3.0.weight"), prefix + ".conv3.1")
    redir = ctx.convbn3d(x, prefix + ".redir", 1, 0, None)
    return F.relu(c3 + redir)


def cva_forward(ctx: _Ctx, prefix, cost, collect=None):
    """cva.forward(downsample=True)  cva.py:59-72.  Returns (logits [B,1,D8,H8,W8], aug)."""
    pooled = F.avg_pool3d(cost, 3, stride=2, padding=1)                      # count_include_pad
    cost_down = ctx.convbn3d(pooled, prefix + ".downsample.1", 1, 1, "relu")
    h = ctx.convbn3d(cost_down, prefix + ".classify.0", 1, 1, "relu")
    logits = ctx.conv3d(h, prefix + ".classify.2.weight", 1, 1).squeeze(1)
    key, cls = semantic_level_key(cost_down, logits)
    aug_down = disparity_attention(ctx, prefix + ".slc_net.cross_attention", cost_down, key)
    aug = F.interpolate(aug_down, scale_factor=(2, 2, 2), mode="trilinear")  # align_corners=False
    fused = ctx.convbn3d(torch.cat([aug, cost], dim=1), prefix + ".fuse.0", 1, 0, None)
    out = multi_aggregation(ctx, prefix + ".cost_agg", fused)
    if collect is not None:
        collect[prefix + ".cost_down"] = cost_down
        collect[prefix + ".logits"] = logits
        collect[prefix + ".class_map"] = cls
        collect[prefix + ".key"] = key
        collect[prefix + ".aug_down"] = aug_down
        collect[prefix + ".fused"] = fused
        collect[prefix + ".out"] = out
    return logits.unsqueeze(1), out


def convex_upsample(ctx: _Ctx, prefix, g, disp):
    """PropgationNet_4x.forward  gwcnet_dca_g.py:117-124, closed form:
    mask channel c = n*16 + i*4 + j, n = 3*(dy+1)+(dx+1); softmax over n;
    out[4h+i,4w+j] = sum_n m[n,i,j,h,w] * 4*disp[h+dy,w+dx] (zero outside the image)."""
    B, _, H, W = disp.shape
    m = ctx.bn(ctx.conv2d(g, prefix + ".conv.0.0.weight"), prefix + ".conv.0.1")
    m = ctx.conv2d(F.relu(m), prefix + ".conv.2.weight")
    m = F.softmax(m.view(B, 9, 4, 4, H, W), dim=1)
    dp = F.pad(4.0 * disp[:, 0], (1, 1, 1, 1))
    out = disp.new_zeros(B, 4, 4, H, W)
    for n in range(9):
        dy, dx = n // 3, n % 3
        out = out + m[:, n] * dp[:, None, None, dy:dy + H, dx:dx + W]
    return out.permute(0, 3, 1, 4, 2).reshape(B, 1, 4 * H, 4 * W)


def hot_path(sd, gwc_l, gwc_r, cat_l, cat_r, g, maxdisp, num_groups=40, operand_bits=None,
             calibrate=False, collect=None, num_cva=3, pv_stage=2):
    """GwcNet.forward (eval) from feature maps to (pred4 [B,1,H,W], prob_volume2 [B,D8,H8,W8]).
    gwcnet_dca_g.py:216-240,282.  `collect` (dict) receives every boundary tensor.
    Stage-count variants (same graph, N cva stages, head classif<N>, class logits of stage `pv_stage` returned):
    gwcnet_dca0_g.py:154-190 (N=0: returns the 1/4-res head logits), gwcnet_dca1_g.py:160-210 (N=1, stage 1),
    gwcnet_dca2_g.py:167-235 (N=2, stage 2), gwcnet_dca4_g.py:214-302 (N=4, stage 4)."""
    ctx = _Ctx(sd, operand_bits, calibrate)
    D4 = maxdisp // 4
    vol = torch.cat([build_gwc_volume(gwc_l, gwc_r, D4, num_groups),
                     build_concat_volume(cat_l, cat_r, D4)], dim=1)
    c = ctx.convbn3d(vol, "dres0.0", 1, 1, "relu")
    c = ctx.convbn3d(c, "dres0.2", 1, 1, "relu")
    r = ctx.convbn3d(c, "dres1.0", 1, 1, "relu")
    cost0 = ctx.convbn3d(r, "dres1.2", 1, 1, None) + c
    cur, out1, pvs = cost0, None, []
    for i in range(num_cva):
        pv, out = cva_forward(ctx, f"cva{i + 1}", cur, collect)
        cur = cost0 + out if i == 0 else out
        if i == 0:
            out1 = cur
        pvs.append(pv)
    h = ctx.convbn3d(cur, f"classif{num_cva}.0", 1, 1, "relu")
    logits = ctx.conv3d(h, f"classif{num_cva}.2.weight", 1, 1).squeeze(1)
    pred_q = disparity_regression(F.softmax(logits, dim=1), D4)
    pred4 = convex_upsample(ctx, "prop", g, pred_q)
    pv_out = logits if num_cva == 0 else pvs[pv_stage - 1].squeeze(1)
    if collect is not None:
        collect.update(volume=vol, dres0=c, cost0=cost0, out1=out1, classif3_logits=logits,
                       pred_quarter=pred_q, pred4=pred4, prob_volume2=pv_out)
    return pred4, pv_out


# --------------------------------------------------------------------------------------------
# the plain-GwcNet baseline models/gwcnet.py
# --------------------------------------------------------------------------------------------
def hourglass_forward(ctx: _Ctx, prefix, x):
    """Full hourglass, gwcnet.py:67-103: two stride-2 stages down, two transposed convs up, 1x1x1 shortcuts."""
    c1 = ctx.convbn3d(x, prefix + ".conv1.0", 2, 1, "relu")
    c2 = ctx.convbn3d(c1, prefix + ".conv2.0", 1, 1, "relu")
    c3 = ctx.convbn3d(c2, prefix + ".conv3.0", 2, 1, "relu")
    c4 = ctx.convbn3d(c3, prefix + ".conv4.0", 1, 1, "relu")
    c5 = F.relu(ctx.bn(ctx.deconv3d(c4, prefix + ".conv5.0.weight"), prefix + ".conv5.1")
                + ctx.convbn3d(c2, prefix + ".redir2", 1, 0))
    return F.relu(ctx.bn(ctx.deconv3d(c5, prefix + ".conv6.0.weight"), prefix + ".conv6.1")
                  + ctx.convbn3d(x, prefix + ".redir1", 1, 0))


def baseline_hot_path(sd, gwc_l, gwc_r, cat_l, cat_r, maxdisp, num_groups=40, calibrate=False, collect=None,
                      vis_size=(24, 67, 120)):
    """models/gwcnet.py GwcNet.forward in eval mode, from the feature maps on (gwcnet.py:199-216, 236-244): volume ->
    dres0 -> dres1 (+) -> dres2 -> dres3 -> vis_tsne1: classif2 logits, rows from 2 on, adaptive average pooling to the
    hard-coded (48//2, 134//2, 240//2) grid (gwcnet.py:186-190).  (dres4 / classif3 are computed by the reference's
    eval branch too, but nothing reads them.)  -> [B, 24, 67, 120]."""
    ctx = _Ctx(sd, None, calibrate)
    D4 = maxdisp // 4
    vol = build_gwc_volume(gwc_l, gwc_r, D4, num_groups)
    if cat_l is not None:
        vol = torch.cat([vol, build_concat_volume(cat_l, cat_r, D4)], dim=1)
    c = ctx.convbn3d(vol, "dres0.0", 1, 1, "relu")
    c = ctx.convbn3d(c, "dres0.2", 1, 1, "relu")
    r = ctx.convbn3d(c, "dres1.0", 1, 1, "relu")
    cost0 = ctx.convbn3d(r, "dres1.2", 1, 1, None) + c
    out1 = hourglass_forward(ctx, "dres2", cost0)
    out2 = hourglass_forward(ctx, "dres3", out1)
    h = ctx.convbn3d(out2, "classif2.0", 1, 1, "relu")
    logits = ctx.conv3d(h, "classif2.2.weight", 1, 1)
    vis = F.adaptive_avg_pool3d(logits[:, :, :, 2:, :], vis_size).squeeze(1)
    if collect is not None:
        collect.update(volume=vol, cost0=cost0, out1=out1, out2=out2, classif2_logits=logits.squeeze(1))
    return vis


def synth_state_dict_from_keys(keyfile, seed, skip=("feature_extraction.", "guidance.")):
    """Seeded weights for every key of a `key shape` listing (tests/golden/state_dict_keys_*.txt) outside the front end:
    conv weights ~ N(0, 2/fan_in) (ConvTranspose3d [Cin,Cout,k,k,k]: fan_in = Cin*27/8), BN weight ~ U(.75, 1.25), bias ~
    N(0, .1), running statistics 0 / 1 (to be calibrated).  Lets a fixture store the calibrated BN statistics only."""
    import ast
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for line in open(keyfile):
        k, shp = line.strip().split(" ", 1)
        if k.startswith(tuple(skip)):
            continue
        shape = ast.literal_eval(shp)
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = torch.zeros(shape)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(shape)
        elif len(shape) == 1 and k.endswith(".weight"):
            sd[k] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif len(shape) == 1:
            sd[k] = torch.randn(shape, generator=g) * 0.1
        else:
            taps = 1
            for d in shape[2:]:
                taps *= d
            transposed = ".conv5.0." in k or ".conv6.0." in k or ".conv3.0.weight" in k and len(shape) == 5 and shape[0] > shape[1]
            fan_in = shape[0] * taps / 8.0 if transposed else shape[1] * taps
            sd[k] = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    return sd


# --------------------------------------------------------------------------------------------
# synthetic, calibrated hot-path checkpoint -- reference-independent
# --------------------------------------------------------------------------------------------
def hot_path_param_shapes(num_groups=40, concat_channels=12, num_cva=3):
    """Reference-layout state_dict entries of the hot path (dres*, cva*, classif*, prop).
    gwcnet_dca_g.py:141-171, cva.py:14-56, SelfAttention_bn.py:20-52.  Stages beyond the third (gwcnet_dca4_g.py:173,191)
    are appended AFTER prop so that the seeded values of everything else do not move."""
    shapes = {}

    def conv3(key, co, ci, k=3):
        shapes[key] = (co, ci, k, k, k)

    def bn(prefix, c):
        for n in ("weight", "bias", "running_mean", "running_var"):
            shapes[f"{prefix}.{n}"] = (c,)
        shapes[f"{prefix}.num_batches_tracked"] = ()

    def cb(prefix, ci, co, k=3):
        conv3(prefix + ".0.weight", co, ci, k)
        bn(prefix + ".1", co)

    cin = num_groups + 2 * concat_channels
    cb("dres0.0", cin, 32); cb("dres0.2", 32, 32)
    cb("dres1.0", 32, 32); cb("dres1.2", 32, 32)
    for i in range(4):
        cb(f"classif{i}.0", 32, 32); conv3(f"classif{i}.2.weight", 1, 32)

    def cva_stage(s):
        p = f"cva{s}"
        cb(p + ".downsample.1", 32, 32)
        a = p + ".slc_net.cross_attention"
        for proj in ("key_project", "query_project"):
            cb(f"{a}.{proj}.0", 32, 32, 1); cb(f"{a}.{proj}.1", 32, 32, 1)
        cb(a + ".value_project", 32, 32, 1); cb(a + ".out_project", 32, 32, 1)
        cb(p + ".classify.0", 32, 32); conv3(p + ".classify.2.weight", 1, 32)
        cb(p + ".fuse.0", 64, 32, 1)
        cb(p + ".cost_agg.conv1.0", 32, 64); cb(p + ".cost_agg.conv2.0", 64, 64)
        shapes[p + ".cost_agg.conv3.0.weight"] = (64, 32, 3, 3, 3)   # ConvTranspose3d: [Cin,Cout,...]
        bn(p + ".cost_agg.conv3.1", 32)
        cb(p + ".cost_agg.redir", 32, 32, 1)

    for s in (1, 2, 3):
        cva_stage(s)
    shapes["prop.conv.0.0.weight"] = (128, 64, 3, 3)
    bn("prop.conv.0.1", 128)
    shapes["prop.conv.2.weight"] = (144, 128, 3, 3)
    for s in range(4, num_cva + 1):
        cb(f"classif{s}.0", 32, 32); conv3(f"classif{s}.2.weight", 1, 32)
        cva_stage(s)
    return shapes


def synth_state_dict(seed=0, num_groups=40, concat_channels=12, num_cva=3):
    """Random hot-path weights with the reference's init law (normal(0, sqrt(2/(k^3*Cout))),
    gwcnet_dca_g.py:173-178) and BN gamma~U(.75,1.25), beta~N(0,.1^2); running stats are
    placeholders until `calibrate_state_dict` runs.  numpy PCG64 so the values are identical on
    every box."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shp in hot_path_param_shapes(num_groups, concat_channels, num_cva).items():
        leaf = key.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            sd[key] = torch.tensor(1, dtype=torch.int64)
        elif len(shp) >= 4:
            co = shp[1] if "conv3.0.weight" in key else shp[0]
            n = co * int(np.prod(shp[2:]))
            sd[key] = torch.from_numpy(rng.normal(0.0, math.sqrt(2.0 / n), shp).astype(np.float32))
        elif leaf == "weight":
            sd[key] = torch.from_numpy(rng.uniform(0.75, 1.25, shp).astype(np.float32))
        elif leaf == "bias":
            sd[key] = torch.from_numpy(rng.normal(0.0, 0.1, shp).astype(np.float32))
        elif leaf == "running_mean":
            sd[key] = torch.zeros(shp)
        else:
            sd[key] = torch.ones(shp)
    return sd


def synth_features(seed, B, H4, W4, shift=3, C=320, Cc=12, Cg=64):
    """Smooth synthetic 1/4-res feature maps: right = left rolled by `shift` columns + noise, so the
    correlation volume has a clear ridge (stands in for the extractor outputs; the 2-D front end is
    out of scope)."""
    g = torch.Generator().manual_seed(seed)

    def smooth(c):
        lo = torch.randn(B, c, max(H4 // 4, 1), max(W4 // 4, 1), generator=g)
        x = F.interpolate(lo, size=(H4, W4), mode="bilinear", align_corners=False)
        return x + 0.1 * torch.randn(B, c, H4, W4, generator=g)

    gl = F.relu(smooth(C))
    gr = torch.roll(gl, -shift, dims=3) + 0.05 * torch.randn(B, C, H4, W4, generator=g)
    cl = 0.5 * smooth(Cc)
    cr = torch.roll(cl, -shift, dims=3) + 0.02 * torch.randn(B, Cc, H4, W4, generator=g)
    gd = smooth(Cg)
    return gl.contiguous(), gr.contiguous(), cl.contiguous(), cr.contiguous(), gd.contiguous()


def calibrate_state_dict(sd, feats, maxdisp, num_groups=40, num_cva=3):
    """One train-mode-BN pass so running stats equal batch stats (unit-scale activations).
    Mutates and returns sd."""
    with torch.no_grad():
        hot_path(sd, *feats, maxdisp=maxdisp, num_groups=num_groups, calibrate=True, num_cva=num_cva)
    return sd
