"""CPU: size-independent properties of the oracle's closed forms (the edge cases the reference's formulation implies:
disparities beyond the image width, ties in the class argmax, empty classes, zero padding of the convex upsampling)."""
import torch
import torch.nn.functional as F

from oracle import dcanet_oracle as O


def test_volume_is_zero_left_of_the_disparity_and_linear_in_each_image():
    g = torch.Generator().manual_seed(0)
    L, R = torch.randn(2, 16, 3, 10, generator=g), torch.randn(2, 16, 3, 10, generator=g)
    D = 14                                              # more disparities than columns (submodule.py:160 loop still runs)
    vol = O.build_gwc_volume(L, R, D, 4)
    cat = O.build_concat_volume(L[:, :6], R[:, :6], D)
    assert vol.shape == (2, 4, D, 3, 10) and cat.shape == (2, 12, D, 3, 10)
    for d in range(D):
        assert float(vol[:, :, d, :, :min(d, 10)].abs().max() if d else 0.0) == 0.0
        assert float(cat[:, :, d, :, :min(d, 10)].abs().max() if d else 0.0) == 0.0   # BOTH halves, submodule.py:141-142
    assert float(vol[:, :, 10:].abs().max()) == 0.0    # d >= W: nothing is ever written
    torch.testing.assert_close(O.build_gwc_volume(2.5 * L, R, D, 4), 2.5 * vol)
    torch.testing.assert_close(O.build_gwc_volume(L, -3.0 * R, D, 4), -3.0 * vol)
    # group mean: one group over all channels = mean of the per-channel products
    one = O.build_gwc_volume(L, R, 1, 1)[:, 0, 0]
    torch.testing.assert_close(one, (L * R).mean(1))


def test_regression_of_a_one_hot_is_its_index_and_softmax_is_shift_invariant():
    D = 12
    for k in (0, 5, D - 1):
        p = torch.zeros(1, D, 2, 3)
        p[:, k] = 1.0
        assert torch.equal(O.disparity_regression(p, D), torch.full((1, 1, 2, 3), float(k)))
    logits = torch.randn(1, D, 2, 3, generator=torch.Generator().manual_seed(1))
    a = O.disparity_regression(F.softmax(logits, 1), D)
    b = O.disparity_regression(F.softmax(logits + 7.0, 1), D)
    torch.testing.assert_close(a, b, rtol=0, atol=1e-5)
    assert float(a.min()) >= 0.0 and float(a.max()) <= D - 1


def test_class_stats_ties_empty_classes_and_mass():
    logits = torch.zeros(1, 4, 2, 2)                   # all tied -> class 0 everywhere (torch.argmax: first index)
    logits[0, :, 1, 1] = torch.tensor([0.0, 2.0, 2.0, 1.0])       # tie between 1 and 2 -> 1
    P, k, e, S, w = O.class_stats(logits)
    assert k.tolist() == [[[0, 0], [0, 1]]]
    assert float(S[0, 2]) == 0.0 and float(S[0, 3]) == 0.0         # empty classes keep a zero sum
    torch.testing.assert_close(S.sum(), e.sum())                   # every pixel lands in exactly one class
    torch.testing.assert_close(w[0, 1, 1], torch.tensor(1.0))      # a class with one pixel: weight 1
    torch.testing.assert_close(w[0, 0, 0] + w[0, 0, 1] + w[0, 1, 0], torch.tensor(1.0))


def test_semantic_key_differs_from_x_only_at_the_class_plane():
    g = torch.Generator().manual_seed(2)
    x, logits = torch.randn(1, 8, 5, 3, 4, generator=g), torch.randn(1, 5, 3, 4, generator=g)
    key, k = O.semantic_level_key(x, logits)
    onehot = F.one_hot(k, 5).permute(0, 3, 1, 2).bool()            # [B,D,H,W]
    assert torch.equal(key[:, :, ~onehot[0]], x[:, :, ~onehot[0]])
    assert bool((key[:, :, onehot[0]].abs() >= x[:, :, onehot[0]].abs()).all())     # scaled by 1 + w, w > 0


def test_convex_upsample_of_a_constant_is_4x_inside_and_sees_zero_padding_at_the_border():
    sd = O.synth_state_dict(0)
    ctx = O._Ctx(sd)
    g = torch.randn(1, 64, 6, 7, generator=torch.Generator().manual_seed(3))
    disp = torch.full((1, 1, 6, 7), 3.0)
    up = O.convex_upsample(ctx, "prop", g, disp)
    assert up.shape == (1, 1, 24, 28)
    torch.testing.assert_close(up[:, :, 4:-4, 4:-4], torch.full((1, 1, 16, 20), 12.0))   # convex weights sum to 1
    assert float(up.max()) <= 12.0 + 1e-4 and float(up[:, :, :4].min()) < 12.0           # F.unfold zero padding
