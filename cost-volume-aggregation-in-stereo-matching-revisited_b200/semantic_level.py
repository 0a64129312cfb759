"""SemanticLevelContext mirror (reference models/augment/semantic_level.py:14-128)."""
import torch.nn as nn

from . import engine
from .self_attention import SelfAttentionBlock


class SemanticLevelContext(nn.Module):
    def __init__(self, feats_channels, transform_channels, reduction=8, concat_input=True, **kwargs):
        super().__init__()
        self.cross_attention = SelfAttentionBlock(
            key_in_channels=feats_channels, query_in_channels=feats_channels, transform_channels=transform_channels,
            out_channels=feats_channels, share_key_query=False, query_downsample=None, key_downsample=None,
            key_query_num_convs=2, value_out_num_convs=1, key_query_norm=True, value_out_norm=True,
            matmul_norm=True, with_out_project=True)
        self.precision_planes = 2
        self.last_class_map = None

    def forward(self, x, preds):
        """x fp32 [B,32,D,H,W], preds = class logits [B,D,H,W] -> context [B,32,D,H,W]."""
        engine._require_cuda(x, preds)
        xp = engine.Planes.from_ncdhw(x, self.precision_planes)
        cls, e, S = engine.class_stats(preds.contiguous().float())
        self.last_class_map = cls
        pk = engine.cached_pack(self.cross_attention, ("attn", self.precision_planes),
                                lambda: engine.PackedAttention(self.cross_attention))
        return engine.disp_attention(xp, cls, e, S, pk.buf, False).to_ncdhw()
