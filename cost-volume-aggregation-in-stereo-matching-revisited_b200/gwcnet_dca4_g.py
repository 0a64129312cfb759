"""Mirror of the reference's models/gwcnet_dca4_g.py:146-302: four cva stages (built there from cva_bn / submodule_bn, which differ from cva / submodule only in importing the SyncBN names), head classif4; eval returns (pred4 [B,1,H,W], prob_volume4).
Same ctor and 2-argument forward as the reference, same state_dict key layout (tests/golden/state_dict_keys_dca4.txt),
same kernels as the 3-stage model: only the stage graph differs (engine.PackedHotPath reads num_cva / pv_stage)."""
from .gwcnet_dca_g import GwcNet as _GwcNet3
from .gwcnet_dca_g import feature_extraction, hourglass  # noqa: F401  (the reference module defines them too)


class GwcNet(_GwcNet3):
    NUM_CVA = 4
    PV_STAGE = 4
    SQUEEZE_PRED = False

    def forward(self, left, right, disp_true=None):     # reference signature: forward(left, right)
        return super().forward(left, right)


def GwcNet_G(d):
    return GwcNet(d, use_concat_volume=False)


def GwcNet_GC(d):
    return GwcNet(d, use_concat_volume=True)
