"""Reference arm of bench.py: the reference's OWN `models.gwcnet_dca_g.GwcNet(maxdisp).eval()` torch forward, fp32, on
the host cores, imported from the staged copy baseline/_ref/ (baseline/stage_reference.py).  None of this repo's
kernels, modules or oracle is on that path; the hot path's share of the forward (feature maps -> disparity) is isolated
with forward hooks on the reference's `feature_extraction` / `guidance` modules, as in SURVEY.md section 8d."""
import os
import sys
import time
import types

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "models", "gwcnet_dca_g.py"))


def import_reference():
    """The two shims of SURVEY 8c: a stub `models` package (models/__init__.py imports a file that is not in the tree)
    and stub matplotlib modules (only used by a visualisation helper)."""
    sys.dont_write_bytecode = True
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = pkg
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import models.gwcnet_dca_g as ref_model
    return ref_model


def synthetic_pair(seed, B, H, W, shift=12):
    g = torch.Generator().manual_seed(seed)
    lo = torch.randn(B, 3, H // 8, W // 8, generator=g)
    left = F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False) + 0.1 * torch.randn(B, 3, H, W, generator=g)
    right = torch.roll(left, -shift, dims=3) + 0.05 * torch.randn(B, 3, H, W, generator=g)
    return left, right


class _FrontEndClock:
    """Accumulates the wall time spent inside the front-end modules (hooks run on the calling thread)."""

    def __init__(self, modules):
        self.total = 0.0
        self._t0 = None
        for m in modules:
            m.register_forward_pre_hook(self._pre)
            m.register_forward_hook(self._post)

    def _pre(self, *_):
        self._t0 = time.perf_counter()

    def _post(self, *_):
        self.total += time.perf_counter() - self._t0


def run(H, W, maxdisp, B, steps, warmup, state_dict, budget_s=150.0):
    """Times `steps` reference forwards on a bounded sample of the workload: a top crop of the pair with the full width
    and disparity range, sized so that (warmup + steps) forwards fit `budget_s`; the hot-path time is scaled by the row
    fraction (every op on the path is linear in the rows).  Returns (pairs_per_s, ms_per_pair, cores, sample, total_ms)."""
    ref_model = import_reference()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    model = ref_model.GwcNet(maxdisp)
    model.load_state_dict(state_dict, strict=True)
    model.eval()
    clock = _FrontEndClock([model.feature_extraction, model.guidance])

    def one(rows):
        left, right = synthetic_pair(0, B, rows, W)
        clock.total = 0.0
        t0 = time.perf_counter()
        with torch.no_grad():
            model(left, right, None)
        total = time.perf_counter() - t0
        return total - clock.total, total

    probe = max(32, (H // 8) // 32 * 32)
    one(probe)                                   # first call: lazy initialisations
    hot_probe, tot_probe = one(probe)
    n_total = max(1, steps + warmup)
    rows = int(min(H, max(32, (budget_s / n_total / (tot_probe / probe)) // 32 * 32)))
    hot, tot = [], []
    for i in range(warmup + steps):
        h, t = one(rows)
        if i >= warmup:
            hot.append(h)
            tot.append(t)
    scale = H / rows
    t_pair = sum(hot) / len(hot) * scale / B
    sample = (f"reference models.gwcnet_dca_g.GwcNet({maxdisp}).eval() torch forward, fp32, {cores} threads; "
              f"{rows}x{W} top crop ({rows}/{H} rows, full W and maxdisp) of the {H}x{W} pair, hot-path time "
              f"(total minus feature_extraction x2 and guidance, forward hooks) scaled by {scale:.2f}; {warmup} warm-up + "
              f"{steps} timed; torch {torch.__version__}")
    return 1.0 / t_pair, t_pair * 1e3, cores, sample, sum(tot) / len(tot) * scale / B * 1e3
