"""Mirror of the reference's plain-GwcNet baseline, models/gwcnet.py: the same volume and
dres0 / dres1 stem as the DCA model, then three FULL hourglass blocks (`hourglass`, up to 128 channels) instead of the cva
stages.  Same constructor, attribute names and state_dict layout (tests/golden/state_dict_keys_gwcnet_{g,gc}.txt); runs on
the same kernels: the volume kernel, the tcgen05 conv family where its channel counts are covered (32 / 64) and the
CUDA-core member for the 128-channel convs.

The reference's eval branch does not return a disparity: it returns `vis_tsne1(out2)` (gwcnet.py:186-190, 236-244), the
classif2 logits of the second hourglass, rows from 2 on, adaptively average-pooled to a hard-coded (24, 67, 120) grid.  This
mirror returns the same tensor (parity: tests/test_baseline_gwcnet.py against a fixture made by the reference module).
"""
import torch
import torch.nn as nn

from . import _lib, engine
from .gwcnet_dca_g import feature_extraction, hourglass
from .submodule import convbn_3d

VIS_SIZE = (48 // 2, 134 // 2, 240 // 2)        # gwcnet.py:188


class GwcNet(nn.Module):
    def __init__(self, maxdisp, use_concat_volume=False, precision="parity"):
        super().__init__()
        self.maxdisp = maxdisp
        self.use_concat_volume = use_concat_volume
        self.num_groups = 40
        if self.use_concat_volume:
            self.concat_channels = 12
            self.feature_extraction = feature_extraction(concat_feature=True, concat_feature_channel=self.concat_channels)
        else:
            self.concat_channels = 0
            self.feature_extraction = feature_extraction(concat_feature=False)
        self.dres0 = nn.Sequential(convbn_3d(self.num_groups + self.concat_channels * 2, 32, 3, 1, 1),
                                   nn.ReLU(inplace=True), convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True))
        self.dres1 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True), convbn_3d(32, 32, 3, 1, 1))
        self.dres2 = hourglass(32)
        self.dres3 = hourglass(32)
        self.dres4 = hourglass(32)
        for i in range(4):
            setattr(self, f"classif{i}", nn.Sequential(
                convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True),
                nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False)))
        assert precision in ("parity", "fast")
        self._planes = 2 if precision == "parity" else 1
        for m in self.modules():
            if hasattr(m, "precision_planes"):
                m.precision_planes = self._planes

    def _pack(self):
        E, P = engine, self._planes
        pk = {}
        w0 = self.dres0[0][0].weight
        if w0.shape[1] % 32:        # 40 input channels (no concat volume): zero-pad to the 64-channel plane row
            wp = torch.zeros((w0.shape[0], 64) + tuple(w0.shape[2:]), dtype=w0.dtype, device=w0.device)
            wp[:, :w0.shape[1]] = w0.detach()
            w0 = wp
        pk["dres0_0"] = E.PackedConv(w0, self.dres0[0][1])
        pk["dres0_2"] = E.pack_convbn(self.dres0[2])
        pk["dres1_0"] = E.pack_convbn(self.dres1[0])
        pk["dres1_2"] = E.pack_convbn(self.dres1[2])
        pk["cls2_0"] = E.pack_convbn(self.classif2[0])
        for k in ("dres0_0", "dres0_2", "dres1_0", "dres1_2", "cls2_0"):
            pk[k].pack_tc(P)
        pk["cls2_2"] = E.PackedCout1(self.classif2[2].weight, P)
        pk["hg"] = [h._pack(P) for h in (self.dres2, self.dres3)]
        return pk

    def hot_path(self, gwc_l, gwc_r, cat_l=None, cat_r=None):
        """Feature maps -> the eval output of gwcnet.py (vis_tsne1 of the second hourglass): fp32 [B, 24, 67, 120]."""
        E, P = engine, self._planes
        E._require_cuda(gwc_l, gwc_r, cat_l, cat_r)
        if (cat_l is None) == self.use_concat_volume:
            raise _lib.DcaError("concat features: needed exactly when use_concat_volume")
        D4 = self.maxdisp // 4
        if D4 % 4 or gwc_l.shape[2] % 4 or gwc_l.shape[3] % 4 or gwc_l.shape[2] <= 2:
            raise _lib.DcaError("the full hourglass needs D/4, H/4 and W/4 to be multiples of 4 (and more than 2 rows)")
        pk = E.cached_pack(self, ("baseline", P), self._pack)
        f = E._f32c
        vol = E.fused_volume(f(gwc_l), f(gwc_r), None if cat_l is None else f(cat_l), None if cat_r is None else f(cat_r),
                             D4, self.num_groups, P, Cv=64)
        c = E.conv(vol, pk["dres0_0"], E.K3S1, E.ACT_RELU)
        c = E.conv(c, pk["dres0_2"], E.K3S1, E.ACT_RELU)
        r = E.conv(c, pk["dres1_0"], E.K3S1, E.ACT_RELU)
        cost0 = E.conv(r, pk["dres1_2"], E.K3S1, E.ACT_NONE, res_post=c)
        out1 = hourglass.forward_planes(cost0, pk["hg"][0])
        out2 = hourglass.forward_planes(out1, pk["hg"][1])
        # (the reference's eval branch also runs dres4 on out2; nothing reads its result, gwcnet.py:214,236-244)
        P27 = E.conv_taps27(out2, pk["cls2_0"], pk["cls2_2"])
        if P27 is not None:
            logits = E.tap_gather(P27)
        else:
            logits = E.conv_cout1_any(E.conv(out2, pk["cls2_0"], E.K3S1, E.ACT_RELU), pk["cls2_2"])
        B, D, H, W = logits.shape
        vis = torch.empty((B,) + VIS_SIZE, dtype=torch.float32, device=logits.device)
        _lib.call("dca_adaptive_avgpool3d_rows", logits.data_ptr(), vis.data_ptr(), B, D, H, W, 2, *VIS_SIZE, E._stream())
        return vis

    def forward(self, left, right, disp_true_down=None):
        if self.training:
            raise NotImplementedError("dcanet_b200 is an inference engine: call .eval()")
        from . import frontend
        B = left.shape[0]
        with frontend._no_tf32():
            f = self.feature_extraction(torch.cat((left, right), dim=0))
        gwc, cat = f["gwc_feature"], f.get("concat_feature")
        return self.hot_path(gwc[:B], gwc[B:], None if cat is None else cat[:B], None if cat is None else cat[B:])


def GwcNet_G(d):
    return GwcNet(d, use_concat_volume=False)


def GwcNet_GC(d):
    return GwcNet(d, use_concat_volume=True)
