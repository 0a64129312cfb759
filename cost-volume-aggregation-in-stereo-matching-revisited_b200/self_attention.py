"""SelfAttentionBlock mirror (reference models/augment/SelfAttention_bn.py:14-160)."""
import torch.nn as nn

from . import engine


class SelfAttentionBlock(nn.Module):
    def __init__(self, key_in_channels, query_in_channels, transform_channels, out_channels, share_key_query,
                 query_downsample, key_downsample, key_query_num_convs, value_out_num_convs, key_query_norm,
                 value_out_norm, matmul_norm, with_out_project, **kwargs):
        super().__init__()
        if (share_key_query or query_downsample is not None or key_downsample is not None
                or key_query_num_convs != 2 or value_out_num_convs != 1 or not key_query_norm or not value_out_norm
                or not matmul_norm or not with_out_project
                or not (key_in_channels == query_in_channels == transform_channels == out_channels == 32)):
            raise NotImplementedError("the B200 kernel implements the configuration DCANet instantiates "
                                      "(semantic_level.py:20-34): 32 channels, 2/1 convs, norms on, out project")
        self.key_project = self.buildproject(key_in_channels, transform_channels, key_query_num_convs, key_query_norm)
        self.query_project = self.buildproject(query_in_channels, transform_channels, key_query_num_convs,
                                               key_query_norm)
        self.value_project = self.buildproject(key_in_channels, transform_channels, value_out_num_convs,
                                               value_out_norm)
        self.out_project = self.buildproject(transform_channels, out_channels, value_out_num_convs, value_out_norm)
        self.query_downsample, self.key_downsample = query_downsample, key_downsample
        self.matmul_norm, self.transform_channels = matmul_norm, transform_channels
        self.precision_planes = 2

    def buildproject(self, in_channels, out_channels, num_convs, use_norm):
        def unit(ci):
            return nn.Sequential(nn.Conv3d(ci, out_channels, kernel_size=1, stride=1, padding=0, bias=False),
                                 nn.BatchNorm3d(out_channels), nn.LeakyReLU(0.1, inplace=True))
        convs = [unit(in_channels)] + [unit(out_channels) for _ in range(num_convs - 1)]
        return nn.Sequential(*convs) if len(convs) > 1 else convs[0]

    def forward(self, query_feats, key_feats):
        """query_feats, key_feats fp32 [B,32,D,H,W] -> context [B,32,D,H,W] (SelfAttention_bn.py:62-98): two-layer
        query/key projections, one-layer value/out projections (1x1x1 conv + BN + LeakyReLU 0.1), 4 heads x 8 channels,
        softmax(q k^T / sqrt(8)) v over the D axis per pixel.  DCANet itself goes through SemanticLevelContext, whose
        kernel call builds the key from the class statistics in place; this is the generic two-input form."""
        engine._require_cuda(query_feats, key_feats)
        if query_feats.shape != key_feats.shape or query_feats.dim() != 5 or query_feats.shape[1] != 32:
            raise engine._lib.DcaError("SelfAttentionBlock.forward: query and key must both be [B,32,D,H,W]")
        P = getattr(self, "precision_planes", 2)
        pk = engine.cached_pack(self, ("attn", P), lambda: engine.PackedAttention(self))
        q = engine.Planes.from_ncdhw(query_feats, P)
        k = engine.Planes.from_ncdhw(key_feats, P)
        return engine.self_attention(q, k, pk.buf).to_ncdhw()
