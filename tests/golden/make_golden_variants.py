"""Golden data for the stage-count variants models/gwcnet_dca{0,1,2,4}_g.py, FROM THE REFERENCE (build container only).

    python tests/golden/make_golden_variants.py

Writes per variant: `state_dict_keys_dca<N>.txt` (key + shape, in the reference's order) and
`variant_dca<N>_32x64_d48.npz` (the reference's eval outputs on one seeded 32x64 pair after a train-mode BN
calibration pass, with the front-end outputs and BN tensors needed to replay the hot path).
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import ROOT, import_reference, synthetic_pair  # noqa: E402


def main():
    from oracle import dcanet_oracle as O
    import_reference()
    torch.set_num_threads(8)
    H, W, maxdisp, seed = 32, 64, 48, 5
    for n in (0, 1, 2, 4):
        if n == 4:      # gwcnet_dca4_g imports the vendored SyncBN package only for names it never instantiates
            import types
            stub = types.ModuleType("models.lib.nn")
            stub.SynchronizedBatchNorm2d = torch.nn.BatchNorm2d
            stub.SynchronizedBatchNorm3d = torch.nn.BatchNorm3d
            sys.modules.setdefault("models.lib", types.ModuleType("models.lib"))
            sys.modules.setdefault("models.lib.nn", stub)
        mod = importlib.import_module(f"models.gwcnet_dca{n}_g")
        torch.manual_seed(seed)
        net = mod.GwcNet(maxdisp)
        with open(os.path.join(HERE, f"state_dict_keys_dca{n}.txt"), "w") as f:
            for k, v in net.state_dict().items():
                f.write(f"{k} {tuple(v.shape)}\n")
        # seeded hot-path weights of this repo + BN affine noise, BN calibrated by one train-mode forward
        sd = O.synth_state_dict(seed, num_cva=max(n, 3))
        own = net.state_dict()
        for k in own:
            if k in sd and not k.startswith(("feature_extraction.", "guidance.")):
                own[k] = sd[k].clone()
        net.load_state_dict(own)
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
                m.momentum = 1.0
        left, right = synthetic_pair(seed, H, W, 8)
        feats = {}
        net.feature_extraction.register_forward_hook(
            lambda m, i, o: feats.setdefault("L" if "L" not in feats else "R", o))
        net.guidance.register_forward_hook(lambda m, i, o: feats.__setitem__("g", o["g"]))
        with torch.no_grad():
            net.train()
            try:
                net(left, right)
            except Exception as e:      # the train-mode returns of some variants reference undefined names
                print(f"dca{n}: train-mode forward raised after the BN statistics were taken: {type(e).__name__}")
            net.eval()
            feats.clear()
            pred, pv = net(left, right)
        out = {"meta": np.array([H, W, maxdisp, seed, n]), "pred": pred.numpy(), "pv": pv.numpy(),
               "gwc_l": feats["L"]["gwc_feature"].numpy(), "gwc_r": feats["R"]["gwc_feature"].numpy(),
               "cat_l": feats["L"]["concat_feature"].numpy(), "cat_r": feats["R"]["concat_feature"].numpy(),
               "g": feats["g"].numpy()}
        for k, v in net.state_dict().items():
            if not k.startswith(("feature_extraction.", "guidance.")) and ("running_" in k):
                out["bn:" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, f"variant_dca{n}_32x64_d48.npz"), **out)
        print(f"dca{n}: {len(net.state_dict())} keys, pred {tuple(pred.shape)}, pv {tuple(pv.shape)}")


if __name__ == "__main__":
    main()
