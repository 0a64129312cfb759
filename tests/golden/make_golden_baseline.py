"""Golden data for the plain-GwcNet baseline models/gwcnet.py, FROM THE REFERENCE (build container only).

    python tests/golden/make_golden_baseline.py

Writes `state_dict_keys_gwcnet_{g,gc}.txt` (key + shape in the reference's order; g: `GwcNet(d, use_concat_volume=False)`,
the default; gc: with the concat volume) and `baseline_gwcnet_gc_32x64_d48.npz`: the reference's eval output (`vis_tsne1`, gwcnet.py:186-190,
236-244) on one seeded 32x64 pair after a train-mode BN calibration pass, with the front-end outputs and the calibrated BN
statistics (the weights are rebuilt from the seed: oracle.synth_state_dict_from_keys).
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import import_reference, synthetic_pair  # noqa: E402


def main():
    import_reference()
    torch.set_num_threads(8)
    mod = importlib.import_module("models.gwcnet")
    H, W, maxdisp, seed = 32, 64, 48, 9
    for tag, ctor in (("g", lambda d: mod.GwcNet(d, False)), ("gc", lambda d: mod.GwcNet(d, True))):
        torch.manual_seed(seed)
        net = ctor(maxdisp)
        with open(os.path.join(HERE, f"state_dict_keys_gwcnet_{tag}.txt"), "w") as f:
            for k, v in net.state_dict().items():
                f.write(f"{k} {tuple(v.shape)}\n")
    # gc variant: this repo's seeded hot-path weights (rebuilt from the key listing by the tests), BN momentum 1 so that one
    # train-mode forward leaves the batch statistics in running_mean / running_var
    from oracle import dcanet_oracle as O
    own = net.state_dict()
    own.update(O.synth_state_dict_from_keys(os.path.join(HERE, "state_dict_keys_gwcnet_gc.txt"), seed))
    net.load_state_dict(own)
    for m in net.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            m.momentum = 1.0
    left, right = synthetic_pair(seed, H, W, 8)
    feats = {}
    net.feature_extraction.register_forward_hook(lambda m, i, o: feats.setdefault("L" if "L" not in feats else "R", o))
    with torch.no_grad():
        net.train()
        net(left, right, None)
        net.eval()
        feats.clear()
        vis = net(left, right, None)
    out = {"meta": np.array([H, W, maxdisp, seed]), "vis": vis.numpy(),
           "gwc_l": feats["L"]["gwc_feature"].numpy(), "gwc_r": feats["R"]["gwc_feature"].numpy(),
           "cat_l": feats["L"]["concat_feature"].numpy(), "cat_r": feats["R"]["concat_feature"].numpy()}
    for k, v in net.state_dict().items():
        if not k.startswith("feature_extraction.") and "running_" in k:
            out["bn:" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "baseline_gwcnet_gc_32x64_d48.npz"), **out)
    print(f"gwcnet_gc: {len(net.state_dict())} keys, vis {tuple(vis.shape)}")


if __name__ == "__main__":
    main()
