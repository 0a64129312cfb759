"""Image side of the reference's inference driver: load + normalise a stereo pair, fit it to the
network's input size, crop the prediction back, write the KITTI 16-bit disparity PNG.  Host-only (numpy / PIL); the
only device work is the model call itself.

Mirrors, function by function:
  load_pair            my_img.py:47-71   per-image, per-channel (x - mean) / std of the 8-bit RGB values
  fit_to_crop          my_img.py:73-89   smaller images: zero-padded at the TOP and RIGHT into crop_h x crop_w;
                                         larger: rows cropped around the centre, columns from 0
  pad_to_multiple      main_dca.py:153-166  top / right zero padding to multiples of 16
  crop_prediction      my_img.py:105-108, main_dca.py:171-174
  save_disparity_png   my_img.py:110     uint16(disp * 256), the KITTI submission format
  predict_files        my_img.py:91-110
"""
import numpy as np
import torch


def _normalised(img):
    """uint8 [H,W,>=3] -> float32 [3,H,W], each channel standardised by ITS OWN mean/std (population std, np.std)."""
    a = np.asarray(img)
    if a.ndim != 3 or a.shape[2] < 3:
        raise ValueError("expected an RGB image [H, W, 3]")
    out = np.empty((3, a.shape[0], a.shape[1]), np.float32)
    for c in range(3):
        ch = a[:, :, c]
        out[c] = (ch - np.mean(ch)) / np.std(ch)
    return out


def load_pair(left, right):
    """Two file names (or PIL images / uint8 arrays) -> float32 [6,H,W]: left RGB then right RGB, standardised."""
    from PIL import Image
    imgs = [Image.open(x) if isinstance(x, (str, bytes)) or hasattr(x, "__fspath__") else x for x in (left, right)]
    l, r = _normalised(imgs[0]), _normalised(imgs[1])
    if l.shape != r.shape:
        raise ValueError(f"left {l.shape[1:]} and right {r.shape[1:]} images differ in size")
    return np.concatenate([l, r], axis=0)


def fit_to_crop(pair, crop_height=384, crop_width=1248):
    """float32 [6,h,w] -> (left [1,3,crop_h,crop_w], right, h, w) as torch tensors."""
    _, h, w = pair.shape
    if h <= crop_height and w <= crop_width:
        fitted = np.zeros((6, crop_height, crop_width), np.float32)
        fitted[:, crop_height - h:, :w] = pair
    else:
        start_y = int((h - crop_height) / 2)
        fitted = pair[:, start_y:start_y + crop_height, :crop_width]
    fitted = np.ascontiguousarray(fitted, dtype=np.float32)
    return torch.from_numpy(fitted[None, 0:3].copy()), torch.from_numpy(fitted[None, 3:6].copy()), h, w


def pad_to_multiple(img, multiple=16):
    """[B,C,H,W] -> (zero-padded at the top and right to multiples of `multiple`, top_pad, right_pad)."""
    H, W = img.shape[2], img.shape[3]
    top = (-H) % multiple
    right = (-W) % multiple
    return torch.nn.functional.pad(img, (0, right, top, 0)), top, right


def crop_prediction(disp, h, w, crop_height=384, crop_width=1248):
    """Undo fit_to_crop on a [crop_h, crop_w] disparity map (no-op for images that were cropped, as in the reference)."""
    if h <= crop_height and w <= crop_width:
        return disp[crop_height - h:crop_height, 0:w]
    return disp


def save_disparity_png(path, disp):
    """KITTI format: 16-bit grayscale PNG of disp * 256 (truncated like numpy's astype('uint16'))."""
    from PIL import Image
    a = (np.asarray(disp, dtype=np.float32) * 256.0).astype("uint16")
    Image.fromarray(a).save(path, format="PNG")       # uint16 arrays map to PIL mode I;16
    return a


def predict_files(model, leftname, rightname, savename=None, crop_height=384, crop_width=1248):
    """my_img.py's `my()`: files in, disparity map [h,w] float32 out (and the PNG written when `savename` is given)."""
    left, right, h, w = fit_to_crop(load_pair(leftname, rightname), crop_height, crop_width)
    dev = next(model.parameters()).device
    model.eval()
    with torch.no_grad():
        out = model(left.to(dev), right.to(dev))
    pred = out[0] if isinstance(out, (tuple, list)) else out      # this repo's GwcNet returns (pred4, prob_volume2)
    disp = crop_prediction(pred.squeeze().detach().float().cpu().numpy(), h, w, crop_height, crop_width)
    if savename is not None:
        save_disparity_png(savename, disp)
    return disp
