"""Host-side helpers with the reference's names and semantics (reference: SinDDM/functions.py).

Only the helpers on the hot path's boundary are here: small utilities, the cosine schedule, `extract`,
`noise_like` and the pyramid builder `create_img_scales` (called by main.py before any model exists).
`dilate_mask` and `match_histograms` serve image2image (harmonization / style transfer); the
reference takes them from scikit-image 0.19.3 (requirements.txt:26), which this image does not have, so their
published algorithms are restated on scipy.ndimage / numpy.  CLIP / ROI helpers (thresholded_grad, stat_from_bbs,
extract_patch) are out of scope.
"""
from __future__ import annotations

import inspect
from pathlib import Path

import numpy as np
import torch
from PIL import Image

APEX_AVAILABLE = False  # the apex fp16 path is dead in the reference (fp16=False hard-coded, main.py:122)


def exists(x):
    """functions.py:72"""
    return x is not None


def default(val, d):
    """functions.py:76 -- `d` may be a value or a zero-argument function."""
    if val is not None:
        return val
    return d() if inspect.isfunction(d) else d


def cycle(dl):
    """functions.py:82 -- endless iteration over a data loader."""
    while True:
        yield from dl


def num_to_groups(num, divisor):
    """functions.py:88 -- [divisor, divisor, ..., remainder]."""
    full, rem = divmod(num, divisor)
    return [divisor] * full + ([rem] if rem > 0 else [])


def loss_backwards(fp16, loss, optimizer, **kwargs):
    """functions.py:97 -- the apex branch cannot be taken here (no apex, fp16 is always False)."""
    if fp16:
        raise RuntimeError("fp16/apex training is not supported (dead path in the reference)")
    loss.backward(**kwargs)


def extract(a, t, x_shape):
    """functions.py:105 -- a[t] reshaped to broadcast over x_shape."""
    picked = a.gather(-1, t)
    return picked.reshape(t.shape[0], *([1] * (len(x_shape) - 1)))


def noise_like(shape, device, repeat=False):
    """functions.py:111 -- same torch.randn call shapes as the reference (RNG-stream parity, quirk Q7)."""
    if repeat:
        one = torch.randn((1, *shape[1:]), device=device)
        return one.repeat(shape[0], *([1] * (len(shape) - 1)))
    return torch.randn(shape, device=device)


def cosine_beta_schedule(timesteps, s=0.008):
    """functions.py:117 -- cosine schedule; keeps the reference's linspace(0, steps, steps) (quirk Q4)."""
    steps = timesteps + 1
    grid = np.linspace(0, steps, steps)
    alpha_bar = np.cos(((grid / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    alpha_bar = alpha_bar / alpha_bar[0]
    betas = 1 - (alpha_bar[1:] / alpha_bar[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


def create_img_scales(foldername, filename, scale_factor=1.411, image_size=None, create=False, auto_scale=None,
                      return_pyramid=False):
    """Pyramid builder, functions.py:130-192.

    return_pyramid=True (new) appends a fifth result: [(level_i, blurry_i or None)] as PIL
    images -- exactly what create=True writes to scale_i/ and scale_i_recon/ (PNG is lossless), so
    MultiscaleTrainer(pyramid=...) can train without a writable dataset folder.

    Returns (sizes [(W, H) per scale], rescale_losses, adjusted scale_factor, n_scales) and, with create=True,
    writes <folder>/scale_i/<name>.png and <folder>/scale_i_recon/<name>.png.  Bit-compatible with the
    reference, including the uint8 wrap-around in the rescale losses (quirk Q1: np.subtract on two PIL
    images stays uint8).
    """
    source = Image.open(foldername + filename)
    png_name = filename.rsplit(".", 1)[0] + ".png"
    if image_size is None:
        image_size = source.size
    if auto_scale is not None:
        shrink = np.sqrt((image_size[0] * image_size[1]) / auto_scale)
        if shrink > 1:
            image_size = (int(image_size[0] / shrink), int(image_size[1] / shrink))

    # coarsest scale: area ~3110 px so the 35-px receptive field covers ~40 % of it, short side in [42, 55]
    short, long_ = min(image_size), max(image_size)
    coarse = int(round(np.sqrt(3110 * short / long_)))
    coarse = min(max(coarse, 42), 55)
    n_scales = int(round(np.log(short / coarse) / np.log(scale_factor)) + 1)
    scale_factor = np.exp(np.log(short / coarse) / (n_scales - 1))

    sizes, pyramid = [], []
    for i in range(n_scales):
        shrink = np.power(scale_factor, n_scales - i - 1)
        size_i = (int(round(image_size[0] / shrink)), int(round(image_size[1] / shrink)))
        level = source.resize(size_i, Image.LANCZOS)
        if create:
            out_dir = Path(foldername + "scale_" + str(i) + "/")
            out_dir.mkdir(parents=True, exist_ok=True)
            level.save(str(out_dir / png_name))
        sizes.append(size_i)
        pyramid.append(level)

    rescale_losses = []
    recons = [None]
    for i in range(n_scales - 1):
        blurry = pyramid[i].resize(sizes[i + 1], Image.BILINEAR)
        diff = np.subtract(pyramid[i + 1], blurry)          # uint8 arithmetic, wraps (Q1)
        rescale_losses.append(np.linalg.norm(diff) / np.asarray(blurry).size)
        recons.append(blurry)
        if create:
            out_dir = Path(foldername + "scale_" + str(i + 1) + "_recon/")
            out_dir.mkdir(parents=True, exist_ok=True)
            blurry.save(str(out_dir / png_name))
    if return_pyramid:
        return sizes, rescale_losses, scale_factor, n_scales, list(zip(pyramid, recons))
    return sizes, rescale_losses, scale_factor, n_scales


# ---------------------------------------------------------------------------------------------------
# image2image helpers (reference SinDDM/functions.py:21-33 and skimage.exposure.match_histograms)
# ---------------------------------------------------------------------------------------------------

def _disk(radius: int) -> np.ndarray:
    """skimage.morphology.disk: (2r+1)^2 footprint of the pixels with x^2 + y^2 <= r^2."""
    yy, xx = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    return (xx * xx + yy * yy) <= radius * radius


def dilate_mask(mask, mode):
    """functions.py:21-33: binary dilation by a disk (radius 7 for harmonization, 20 for editing), Gaussian blur
    (sigma 5, skimage defaults: mode='nearest', truncate=4), min-max normalisation.  `mask` is a [C,H,W] tensor in
    [0,1] (transforms.ToTensor of the mask image); only channel 0 is used, any non-zero value is foreground.
    Returns a float64 numpy array [1,1,H,W] like the reference."""
    from scipy import ndimage as ndi
    if mode == "harmonization":
        element = _disk(7)
    elif mode == "editing":
        element = _disk(20)
    else:
        raise ValueError(f"dilate_mask: unknown mode {mode!r}")
    m = np.asarray(mask.detach().cpu() if torch.is_tensor(mask) else mask)[0] != 0
    m = ndi.binary_dilation(m, structure=element)
    m = ndi.gaussian_filter(m.astype(np.float64), sigma=5, mode="nearest", truncate=4.0)
    m = m[None, None, :, :]
    return (m - m.min()) / (m.max() - m.min())


def _match_cumulative_cdf(source: np.ndarray, template: np.ndarray) -> np.ndarray:
    """skimage/exposure/histogram_matching.py::_match_cumulative_cdf (0.19): map every source value to the template
    value of the same quantile (linear interpolation between template quantiles)."""
    if source.dtype.kind == "u":
        src_lookup = source.reshape(-1)
        src_counts = np.bincount(src_lookup)
        tmpl_counts = np.bincount(template.reshape(-1))
        tmpl_values = np.nonzero(tmpl_counts)[0]
        tmpl_counts = tmpl_counts[tmpl_values]
    else:
        _, src_lookup, src_counts = np.unique(source.reshape(-1), return_inverse=True, return_counts=True)
        tmpl_values, tmpl_counts = np.unique(template.reshape(-1), return_counts=True)
    src_quantiles = np.cumsum(src_counts) / source.size
    tmpl_quantiles = np.cumsum(tmpl_counts) / template.size
    interp_a_values = np.interp(src_quantiles, tmpl_quantiles, tmpl_values)
    return interp_a_values[src_lookup].reshape(source.shape)


def match_histograms(image: np.ndarray, reference: np.ndarray, channel_axis=None) -> np.ndarray:
    """skimage.exposure.match_histograms (0.19.3, what trainer.py:313 calls with channel_axis=2): per channel CDF
    matching; the result has the dtype of `image` (uint8 images come back as uint8, values truncated)."""
    if image.ndim != reference.ndim:
        raise ValueError("Image and reference must have the same number of channels.")
    if channel_axis is None:
        return _match_cumulative_cdf(image, reference).astype(image.dtype, copy=False)
    image = np.moveaxis(image, channel_axis, -1)
    reference = np.moveaxis(reference, channel_axis, -1)
    if image.shape[-1] != reference.shape[-1]:
        raise ValueError("Number of channels in the input image and reference image must match!")
    matched = np.empty(image.shape, dtype=image.dtype)
    for ch in range(image.shape[-1]):
        matched[..., ch] = _match_cumulative_cdf(image[..., ch], reference[..., ch])
    return np.moveaxis(matched, -1, channel_axis)
