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

    def buildproject(self, in_channels, out_channels, num_convs, use_norm):
        def unit(ci):
            return nn.Sequential(nn.Conv3d(ci, out_channels, kernel_size=1, stride=1, padding=0, bias=False),
                                 nn.BatchNorm3d(out_channels), nn.LeakyReLU(0.1, inplace=True))
        convs = [unit(in_channels)] + [unit(out_channels) for _ in range(num_convs - 1)]
        return nn.Sequential(*convs) if len(convs) > 1 else convs[0]

    def forward(self, query_feats, key_feats):
        """Generic two-input form is not what DCANet calls; SemanticLevelContext drives the fused kernel
        (key = query * class-wise scale).  Kept for API parity: key_feats must be that scaled query."""
        raise NotImplementedError("use SemanticLevelContext.forward (fused key construction + attention)")
