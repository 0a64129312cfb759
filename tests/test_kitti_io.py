"""CPU: image side of the inference driver (kitti_io.py) against the semantics of my_img.py:47-110 / main_dca.py:153-175."""
import importlib

import numpy as np
import torch
from PIL import Image

io = importlib.import_module("cost-volume-aggregation-in-stereo-matching-revisited_b200.kitti_io")


def _rgb(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def test_load_pair_standardises_each_channel_of_each_image(tmp_path):
    l, r = _rgb(0, 37, 53), _rgb(1, 37, 53)
    Image.fromarray(l).save(tmp_path / "l.png")
    Image.fromarray(r).save(tmp_path / "r.png")
    pair = io.load_pair(str(tmp_path / "l.png"), str(tmp_path / "r.png"))
    assert pair.shape == (6, 37, 53) and pair.dtype == np.float32
    for c in range(6):
        assert abs(float(pair[c].mean())) < 1e-5 and abs(float(pair[c].std()) - 1.0) < 1e-5
    src = l[:, :, 1].astype(np.float64)
    np.testing.assert_allclose(pair[1], (src - src.mean()) / src.std(), rtol=0, atol=1e-5)


def test_small_images_are_padded_top_and_right_and_cropped_back():
    pair = np.random.default_rng(2).standard_normal((6, 370, 1226)).astype(np.float32)     # a KITTI 2015 size
    left, right, h, w = io.fit_to_crop(pair)
    assert left.shape == right.shape == (1, 3, 384, 1248) and (h, w) == (370, 1226)
    assert torch.equal(left[0, :, 14:, :1226], torch.from_numpy(pair[0:3]))
    assert torch.equal(right[0, :, 14:, :1226], torch.from_numpy(pair[3:6]))
    assert float(left[0, :, :14].abs().max()) == 0 and float(left[0, :, :, 1226:].abs().max()) == 0
    disp = np.arange(384 * 1248, dtype=np.float32).reshape(384, 1248)
    back = io.crop_prediction(disp, h, w)
    assert back.shape == (370, 1226) and back[0, 0] == disp[14, 0] and back[-1, -1] == disp[383, 1225]


def test_large_images_are_cropped_rows_centred_columns_from_zero():
    pair = np.random.default_rng(3).standard_normal((6, 400, 1300)).astype(np.float32)
    left, _, h, w = io.fit_to_crop(pair)
    assert left.shape == (1, 3, 384, 1248)
    assert torch.equal(left[0], torch.from_numpy(pair[0:3, 8:392, :1248]))
    disp = np.zeros((384, 1248), np.float32)
    assert io.crop_prediction(disp, h, w).shape == (384, 1248)          # the reference leaves these un-cropped


def test_pad_to_multiple_of_16_is_top_right():
    x = torch.arange(2 * 3 * 375 * 1242, dtype=torch.float32).view(2, 3, 375, 1242)
    y, top, right = io.pad_to_multiple(x)
    assert (top, right) == (9, 6) and y.shape == (2, 3, 384, 1248)
    assert torch.equal(y[:, :, 9:, :1242], x) and float(y[:, :, :9].abs().max()) == 0
    z, top, right = io.pad_to_multiple(torch.zeros(1, 3, 384, 1248))
    assert (top, right) == (0, 0) and z.shape == (1, 3, 384, 1248)


def test_kitti_png_is_uint16_of_disp_times_256(tmp_path):
    disp = np.random.default_rng(4).uniform(0, 191.99, (20, 31)).astype(np.float32)
    io.save_disparity_png(str(tmp_path / "d.png"), disp)
    back = np.asarray(Image.open(tmp_path / "d.png"))
    assert back.dtype == np.uint16 or back.dtype == np.int32
    np.testing.assert_array_equal(back.astype(np.int64), (disp * 256.0).astype("uint16").astype(np.int64))
    assert float(np.abs(back / 256.0 - disp).max()) < 1.0 / 256.0


def test_predict_files_drives_a_model_like_my_img(tmp_path):
    l, r = _rgb(5, 40, 60), _rgb(6, 40, 60)
    Image.fromarray(l).save(tmp_path / "l.png")
    Image.fromarray(r).save(tmp_path / "r.png")

    class Fake(torch.nn.Module):           # stands in for GwcNet: returns (pred4 [B,1,H,W], prob_volume2)
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        def forward(self, left, right):
            assert left.shape == (1, 3, 48, 64)
            rows = torch.arange(48.0).view(1, 1, 48, 1).expand(1, 1, 48, 64)
            return rows + self.p, None

    disp = io.predict_files(Fake(), str(tmp_path / "l.png"), str(tmp_path / "r.png"), str(tmp_path / "o.png"), 48, 64)
    assert disp.shape == (40, 60) and disp[0, 0] == 8.0 and disp[-1, 0] == 47.0
    assert np.asarray(Image.open(tmp_path / "o.png"))[0, 0] == 8 * 256


def test_against_the_references_own_load_and_transform():
    """tests/golden/io_my_img.npz: outputs of my_img.py's load_data / my_transform on seeded images (crop 48x64)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io_my_img.npz"))
    for tag in ("small", "large"):
        pair = io.load_pair(Image.fromarray(z[f"{tag}:l"]), Image.fromarray(z[f"{tag}:r"]))
        np.testing.assert_allclose(pair, z[f"{tag}:data"], rtol=0, atol=2e-6)
        left, right, h, w = io.fit_to_crop(z[f"{tag}:data"], 48, 64)
        assert [h, w] == z[f"{tag}:hw"].tolist()
        np.testing.assert_array_equal(left.numpy(), z[f"{tag}:left"])
        np.testing.assert_array_equal(right.numpy(), z[f"{tag}:right"])


class _FakeStereo(torch.nn.Module):
    """A stand-in model with GwcNet's call signature and return shapes: 'disparity' = a fixed function of both images."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor([0.5, 1.5, 2.5]))

    def forward(self, left, right, disp_true=None):
        d = (left * self.w.view(1, 3, 1, 1)).sum(1, keepdim=True).abs() + right.mean(1, keepdim=True).abs()
        return d, d[:, :, ::8, ::8]


def test_image_pipeline_equals_the_serial_driver(tmp_path):
    """ImagePipeline (loader threads + slot ring + writer thread) = predict_files pair by pair, PNG bytes included; more
    pairs than slots, mixed sizes (padded, exact, cropped)."""
    sizes = [(37, 53), (48, 64), (60, 70), (20, 64), (48, 30), (37, 53), (52, 80)]
    triples, ref = [], []
    model = _FakeStereo()
    for i, (h, w) in enumerate(sizes):
        Image.fromarray(_rgb(10 + i, h, w)).save(tmp_path / f"l{i}.png")
        Image.fromarray(_rgb(50 + i, h, w)).save(tmp_path / f"r{i}.png")
        triples.append((str(tmp_path / f"l{i}.png"), str(tmp_path / f"r{i}.png"), str(tmp_path / f"p{i}.png")))
        ref.append(io.predict_files(model, triples[-1][0], triples[-1][1], str(tmp_path / f"q{i}.png"), 48, 64))
    pipe = io.ImagePipeline(model, crop_height=48, crop_width=64, depth=2, workers=3)
    got = pipe.run(triples)
    assert len(got) == len(sizes)
    for i, (g, r) in enumerate(zip(got, ref)):
        assert g.shape == r.shape and np.array_equal(g, r), i
        assert np.array_equal(np.asarray(Image.open(tmp_path / f"p{i}.png")), np.asarray(Image.open(tmp_path / f"q{i}.png")))
    again = pipe.run(triples[:3])                       # the ring is reusable
    assert all(np.array_equal(a, b) for a, b in zip(again, ref[:3]))


def test_image_pipeline_surfaces_loader_errors(tmp_path):
    import pytest
    Image.fromarray(_rgb(1, 20, 30)).save(tmp_path / "l.png")
    Image.fromarray(_rgb(2, 21, 30)).save(tmp_path / "r.png")          # sizes differ -> load_pair raises
    pipe = io.ImagePipeline(_FakeStereo(), 32, 32)
    with pytest.raises(ValueError):
        pipe.run([(str(tmp_path / "l.png"), str(tmp_path / "r.png"), None)])
