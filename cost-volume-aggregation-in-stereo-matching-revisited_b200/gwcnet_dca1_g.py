"""Mirror of the reference's models/gwcnet_dca1_g.py:108-210: one cva stage, head classif1; eval returns (pred1.squeeze(1), prob_volume1).
Same ctor and 2-argument forward as the reference, same state_dict key layout (tests/golden/state_dict_keys_dca1.txt),
same kernels as the 3-stage model: only the stage graph differs (engine.PackedHotPath reads num_cva / pv_stage)."""
from .gwcnet_dca_g import GwcNet as _GwcNet3
from .gwcnet_dca_g import feature_extraction, hourglass  # noqa: F401  (the reference module defines them too)


class GwcNet(_GwcNet3):
    NUM_CVA = 1
    PV_STAGE = 1
    SQUEEZE_PRED = True

    def forward(self, left, right, disp_true=None):     # reference signature: forward(left, right)
        return super().forward(left, right)


def GwcNet_G(d):
    return GwcNet(d, use_concat_volume=False)


def GwcNet_GC(d):
    return GwcNet(d, use_concat_volume=True)
