"""Host-side logic that needs no GPU: module surface / state_dict contract, EMA, trainer bookkeeping pieces."""
import copy
import tempfile

import numpy as np
import pytest
import torch

from conftest import GOLDEN_SCALE_LOSSES, GOLDEN_SIZES
from oracle import sinddm_oracle as orc


def make(dim=160):
    from sinddm_b200 import MultiScaleGaussianDiffusion, SinDDMNet
    net = SinDDMNet(dim=dim, multiscale=True)
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=5, scale_factor=1.36, image_sizes=GOLDEN_SIZES,
                                      timesteps=100, train_full_t=True, scale_losses=GOLDEN_SCALE_LOSSES,
                                      results_folder=tempfile.mkdtemp())
    return net, dif


def test_state_dict_contract_65_keys():
    """13 buffers + 52 denoise_fn.* parameters, names and shapes of the reference."""
    net, dif = make()
    sd = dif.state_dict()
    assert len(sd) == 65
    expect = [("denoise_fn." + n, s) for n, s in orc.param_shapes(160)]
    got = [(k, tuple(v.shape)) for k, v in sd.items() if k.startswith("denoise_fn.")]
    assert got == expect
    assert [n for n, _ in net.named_parameters()] == [n for n, _ in orc.param_shapes(160)]
    assert sum(p.numel() for p in net.parameters()) == 1_106_772


def test_drop_in_import_paths():
    import SinDDM.functions as F
    import SinDDM.models as M
    import SinDDM.trainer as T
    from text2live_util.util import get_augmentations_template  # noqa: F401
    from text2live_util.clip_extractor import ClipExtractor  # noqa: F401
    for name in ("SinDDMNet", "MultiScaleGaussianDiffusion", "EMA", "SinDDMConvBlock", "SinusoidalPosEmb"):
        assert hasattr(M, name)
    assert hasattr(T, "MultiscaleTrainer") and hasattr(T, "Dataset")
    for name in ("create_img_scales", "extract", "cosine_beta_schedule", "noise_like", "default", "exists",
                 "num_to_groups", "cycle", "loss_backwards"):
        assert hasattr(F, name)


def test_helpers_match_reference_semantics():
    from sinddm_b200.functions import cosine_beta_schedule, default, extract, num_to_groups
    assert num_to_groups(16, 32) == [16]
    assert num_to_groups(70, 32) == [32, 32, 6]
    assert default(None, lambda: 5) == 5 and default(3, 4) == 3
    a = torch.arange(10.0)
    t = torch.tensor([1, 7])
    assert extract(a, t, (2, 3, 4, 4)).shape == (2, 1, 1, 1)
    np.testing.assert_array_equal(cosine_beta_schedule(100), orc.cosine_beta_schedule(100))


def test_sinusoidal_embedding_matches_oracle():
    from sinddm_b200 import SinusoidalPosEmb
    x = torch.tensor([0, 3, 99])
    assert torch.equal(SinusoidalPosEmb(32)(x), orc.sinusoidal_pos_emb(x, 32))


def test_deepcopy_and_ema_semantics():
    from sinddm_b200 import EMA
    net, dif = make(dim=16)
    twin = copy.deepcopy(dif)                        # trainer.py:100
    assert twin.denoise_fn._runtime is not dif.denoise_fn._runtime
    with torch.no_grad():
        for p in dif.parameters():
            p.add_(1.0)
    before = [p.clone() for p in twin.parameters()]
    EMA(0.995).update_model_average(twin, dif)
    for b, new, cur in zip(before, twin.parameters(), dif.parameters()):
        assert torch.allclose(new, b * 0.995 + (1 - 0.995) * cur)


def test_checkpoint_keys_load_like_the_authors_files():
    """A state_dict with the reference's 65 keys loads strictly (the shipped model-12.pt files have them)."""
    net, dif = make()
    sd = {("denoise_fn." + k): v for k, v in orc.synthetic_params(5, 160).items()}
    sd.update(orc.Schedule(5, GOLDEN_SCALE_LOSSES, 100, train_full_t=True).buffers())
    missing = dif.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys


def test_guidance_modes_refuse_loudly():
    from sinddm_b200 import MultiscaleTrainer
    net, dif = make(dim=16)
    dif.clip_guided_sampling = True
    with pytest.raises(NotImplementedError):
        dif._refuse_guidance()
    for name in ("clip_sampling", "clip_roi_sampling", "roi_guided_sampling"):
        with pytest.raises(NotImplementedError):
            getattr(MultiscaleTrainer, name)(None)


def test_dataset_yields_first_image_replicated(tmp_path):
    from PIL import Image
    from sinddm_b200.trainer import Dataset
    d0 = tmp_path / "scale_1"
    d1 = tmp_path / "scale_1_recon"
    d0.mkdir(); d1.mkdir()
    rs = np.random.RandomState(0)
    a = rs.randint(0, 255, (9, 11, 3), dtype=np.uint8)
    b = rs.randint(0, 255, (9, 11, 3), dtype=np.uint8)
    Image.fromarray(a).save(d0 / "img.png")
    Image.fromarray(b).save(d1 / "img.png")
    ds = Dataset(str(d0), None, blurry_img=True)
    assert len(ds) == 128
    orig, blur = ds.batch(200)                       # the reference's DataLoader caps at len(ds) = 128
    assert orig.shape == (128, 3, 9, 11) and blur.shape == (128, 3, 9, 11)
    assert torch.equal(orig[0], torch.from_numpy(a).permute(2, 0, 1).float() / 255 * 2 - 1)
    assert torch.equal(blur[5], torch.from_numpy(b).permute(2, 0, 1).float() / 255 * 2 - 1)


def test_cli_accepts_the_reference_flags():
    """main.py parses every flag of the reference CLI (main.py:15-58)."""
    import importlib.util
    import pathlib
    spec = importlib.util.spec_from_file_location("sinddm_main", pathlib.Path(__file__).parents[1] / "main.py")
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    args = cli.build_parser().parse_args(
        "--scope balloons --mode train --input_image a.png --start_t_harm 5 --start_t_style 15 --harm_mask m.png "
        "--clip_text hello --fill_factor 0.5 --strength 0.3 --roi_n_tar 2 --dataset_folder ./d/ --image_name b.png "
        "--results_folder ./r/ --dim 160 --scale_factor 1.411 --timesteps 100 --train_batch_size 32 "
        "--grad_accumulate 1 --train_num_steps 120001 --save_and_sample_every 10000 --avg_window 100 "
        "--train_lr 0.001 --sched_k_milestones 20 40 70 80 90 110 --load_milestone 0 --sample_batch_size 16 "
        "--scale_mul 1 1 --sample_t_list 52 41 31 22 --device_num 0 --sample_limited_t --omega 0 "
        "--loss_factor 1".split())
    assert args.sched_k_milestones == [20, 40, 70, 80, 90, 110] and args.scale_mul == [1.0, 1.0]
    with pytest.raises(SystemExit):
        cli.main(["--mode", "clip_content"])


def test_dilate_mask_matches_the_published_skimage_pipeline():
    """functions.py:21-33 restated on scipy (skimage 0.19.3 is absent): disk dilation + sigma-5 blur + min-max."""
    from scipy import ndimage as ndi
    from sinddm_b200.functions import dilate_mask
    m = torch.zeros(3, 60, 80)
    m[:, 25:35, 30:50] = 0.7            # any non-zero value is foreground
    out = dilate_mask(m, mode="harmonization")
    assert out.shape == (1, 1, 60, 80) and out.dtype == np.float64
    assert out.min() == 0.0 and out.max() == 1.0
    assert out[0, 0, 30, 40] == 1.0 and out[0, 0, 0, 0] < 1e-6
    # radius-7 disk: the binary support grows by exactly 7 pixels along the axes before the blur
    yy, xx = np.mgrid[-7:8, -7:8]
    dil = ndi.binary_dilation(m[0].numpy() != 0, structure=(xx * xx + yy * yy) <= 49)
    assert dil[25 - 7, 40] and not dil[25 - 8, 40] and dil[30, 30 - 7] and not dil[30, 30 - 8]
    assert not dil[25 - 6, 30 - 6]                                   # disk, not square
    ref = ndi.gaussian_filter(dil.astype(np.float64), 5, mode="nearest", truncate=4.0)
    assert np.allclose(out[0, 0], (ref - ref.min()) / (ref.max() - ref.min()))
    big = dilate_mask(m, mode="editing")
    assert (big >= out - 1e-12).all() and big.sum() > out.sum()
    with pytest.raises(ValueError):
        dilate_mask(m, mode="nope")


def test_match_histograms_is_cdf_matching_per_channel():
    """skimage.exposure.match_histograms (trainer.py:313) restated: uint8 in -> uint8 out, per channel, identity on
    itself, monotone, and the matched image takes the reference's quantiles."""
    from sinddm_b200.functions import match_histograms
    rs = np.random.RandomState(0)
    img = rs.randint(0, 120, size=(40, 50, 3)).astype(np.uint8)
    ref = (rs.beta(2, 5, size=(30, 70, 3)) * 255).astype(np.uint8)
    out = match_histograms(img, ref, channel_axis=2)
    assert out.dtype == np.uint8 and out.shape == img.shape
    assert np.array_equal(match_histograms(img, img, channel_axis=2), img)
    for c in range(3):
        order = np.argsort(img[..., c].ravel(), kind="stable")
        assert (np.diff(out[..., c].ravel()[order].astype(int)) >= 0).all()          # monotone mapping
        for q in (0.1, 0.5, 0.9):
            assert abs(np.quantile(out[..., c], q) - np.quantile(ref[..., c], q)) <= 3
    f = match_histograms(img.astype(np.float64), ref.astype(np.float64), channel_axis=2)
    assert f.dtype == np.float64 and np.abs(f - out).max() < 1.0 + 1e-9               # uint8 result = truncation
    with pytest.raises(ValueError):
        match_histograms(img, ref[..., :2], channel_axis=2)


def test_in_memory_pyramid_equals_the_png_round_trip(tmp_path):
    """MultiscaleTrainer(pyramid=...) holds bit-identical training tensors to the reference flow
    (create_img_scales(create=True) -> scale_i/*.png -> Dataset), for an RGB and an RGBA source image."""
    from PIL import Image
    from sinddm_b200 import MultiScaleGaussianDiffusion, MultiscaleTrainer, SinDDMNet, create_img_scales
    rs = np.random.RandomState(3)
    for mode in ("RGB", "RGBA"):
        folder = str(tmp_path / mode) + "/"
        (tmp_path / mode).mkdir()
        arr = rs.randint(0, 255, (120, 160, len(mode)), dtype=np.uint8)
        Image.fromarray(arr, mode).save(folder + "img.png")
        sizes, losses, sf, ns, pyr = create_img_scales(folder, "img.png", create=True, auto_scale=50000,
                                                       return_pyramid=True)
        assert len(pyr) == ns and pyr[0][1] is None and [p[0].size for p in pyr] == sizes
        assert (sizes, losses, sf, ns) == create_img_scales(folder, "img.png", create=False, auto_scale=50000)
        net = SinDDMNet(dim=16, multiscale=True)
        dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=ns, scale_factor=sf, image_sizes=sizes,
                                          timesteps=100, train_full_t=True, scale_losses=losses,
                                          results_folder=str(tmp_path / "r"))
        kw = dict(n_scales=ns, scale_factor=sf, image_sizes=sizes, train_batch_size=3,
                  results_folder=str(tmp_path / "r"), device="cpu")
        from_files = MultiscaleTrainer(dif, folder, **kw)
        from_memory = MultiscaleTrainer(dif, None, pyramid=pyr, **kw)
        for (a0, a1), (b0, b1) in zip(from_files.data_list, from_memory.data_list):
            assert torch.equal(a0, b0) and torch.equal(a1, b1) and a0.shape[0] == 3


def test_flattened_parameters_keep_the_module_surface():
    """The fused optimizer step re-points every nn.Parameter at one flat buffer (fused_optim._flatten_params):
    values, names, shapes and state_dict stay what they were, and the parameters alias the flat vector."""
    from sinddm_b200.fused_optim import _flatten_params
    net, _ = make(dim=16)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    params = list(net.parameters())
    flat = _flatten_params(params, 4)
    assert flat.numel() % 4 == 0 and flat.numel() >= sum(p.numel() for p in params)
    after = net.state_dict()
    assert list(after) == list(before) and all(torch.equal(after[k], before[k]) for k in before)
    off = 0
    for p in params:
        assert p.data_ptr() == flat.data_ptr() + 4 * off and p.is_contiguous()
        off += p.numel()
    flat.add_(1.0)                                    # a kernel writing the flat vector updates every parameter
    assert all(torch.equal(net.state_dict()[k], before[k] + 1.0) for k in before)


# ---- image2image helpers: known answers and an independent restatement ------------------------

def test_dilate_mask_known_answers_and_bruteforce_restatement():
    """functions.py:21-33 = skimage.morphology.binary_dilation(disk(r)) -> skimage.filters.gaussian(sigma=5) ->
    min-max normalisation.  skimage is absent here, so the scipy-based product code is pinned by
      (1) the disk footprint's known lattice-point counts (Gauss circle problem: N(7) = 149, N(20) = 1257),
      (2) a brute-force restatement written from the definitions (OR over footprint shifts; separable Gaussian with
          radius int(4 * sigma + 0.5), weights exp(-x^2 / (2 sigma^2)) normalised, edge replication), and
      (3) symmetry / range properties."""
    import torch
    from sinddm_b200.functions import _disk, dilate_mask
    assert int(_disk(7).sum()) == 149 and int(_disk(20).sum()) == 1257
    assert _disk(7).shape == (15, 15) and _disk(7)[0, 7] and not _disk(7)[0, 6]

    H, W = 45, 61
    mask = np.zeros((3, H, W), np.float32)
    mask[0, 22, 30] = 1.0                 # a single pixel ...
    mask[0, 5:8, 50:58] = 0.3             # ... and a block near the border (any non-zero value is foreground)
    got = dilate_mask(torch.from_numpy(mask), mode="harmonization")
    assert got.shape == (1, 1, H, W) and got.dtype == np.float64

    # brute force
    fg = mask[0] != 0
    dil = np.zeros((H, W), bool)
    for y in range(H):
        for x in range(W):
            if fg[y, x]:
                for dy in range(-7, 8):
                    for dx in range(-7, 8):
                        if dy * dy + dx * dx <= 49 and 0 <= y + dy < H and 0 <= x + dx < W:
                            dil[y + dy, x + dx] = True
    assert int(dil[15:30, 23:38].sum()) == 149                                  # the single pixel became the disk
    r = int(4.0 * 5 + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / 5.0) ** 2)
    k /= k.sum()
    img = dil.astype(np.float64)
    tmp = np.zeros_like(img)
    for y in range(H):                                                           # rows, then columns; edge replication
        for x in range(W):
            tmp[y, x] = sum(k[j + r] * img[y, min(max(x + j, 0), W - 1)] for j in range(-r, r + 1))
    out = np.zeros_like(img)
    for y in range(H):
        for x in range(W):
            out[y, x] = sum(k[j + r] * tmp[min(max(y + j, 0), H - 1), x] for j in range(-r, r + 1))
    want = (out - out.min()) / (out.max() - out.min())
    np.testing.assert_allclose(got[0, 0], want, rtol=0, atol=1e-12)
    assert got.min() == 0.0 and got.max() == 1.0

    # an isolated pixel gives a result symmetric under the square's symmetries
    m2 = np.zeros((1, 41, 41), np.float32)
    m2[0, 20, 20] = 1
    g2 = dilate_mask(torch.from_numpy(m2), mode="editing")[0, 0]
    np.testing.assert_allclose(g2, g2[::-1, :], atol=1e-14)
    np.testing.assert_allclose(g2, g2.T, atol=1e-14)
    assert g2[20, 20] == 1.0
    with pytest.raises(ValueError):
        dilate_mask(torch.from_numpy(m2), mode="nope")


def test_match_histograms_known_answers():
    """skimage.exposure.match_histograms 0.19 (trainer.py:313 calls it with channel_axis=2): every source value maps to
    the reference value of the same quantile, np.interp between reference quantiles, result cast to the input dtype.
    Hand-derived cases."""
    from sinddm_b200.functions import match_histograms
    # one-to-one quantiles
    src = np.array([[0, 1], [2, 3]], np.uint8)
    ref = np.array([[10, 20], [30, 40]], np.uint8)
    np.testing.assert_array_equal(match_histograms(src, ref), ref)
    # ties: source quantiles .5, .75, 1.0 against reference quantiles .25, .5, .75, 1.0
    np.testing.assert_array_equal(match_histograms(np.array([0, 0, 1, 2], np.uint8), np.array([10, 20, 30, 40], np.uint8)),
                                  [20, 20, 30, 40])
    # interpolation: reference {0, 10} has quantiles .5 and 1.0; source quantiles .25 .5 .75 1.0 (exact in binary)
    np.testing.assert_array_equal(match_histograms(np.arange(4, dtype=np.uint8), np.array([0, 10], np.uint8)),
                                  [0, 0, 5, 10])
    # truncation to the input dtype (not rounding): reference {0, 3}: .75 -> 1.5 -> 1
    np.testing.assert_array_equal(match_histograms(np.arange(4, dtype=np.uint8), np.array([0, 3], np.uint8)),
                                  [0, 0, 1, 3])
    # float images keep float values
    out = match_histograms(np.array([0.0, 1.0, 2.0, 3.0]), np.array([0.0, 3.0]))
    np.testing.assert_allclose(out, [0.0, 0.0, 1.5, 3.0])
    # an image matched to itself is unchanged; per-channel matching treats channels independently
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, (17, 13, 3)).astype(np.uint8)
    np.testing.assert_array_equal(match_histograms(img, img, channel_axis=2), img)
    ref3 = rs.randint(0, 256, (9, 11, 3)).astype(np.uint8)
    m = match_histograms(img, ref3, channel_axis=2)
    for c in range(3):
        np.testing.assert_array_equal(m[..., c], match_histograms(img[..., c], ref3[..., c]))
        # matched values never leave the reference channel's range, order of pixels is preserved
        assert m[..., c].min() >= ref3[..., c].min() and m[..., c].max() <= ref3[..., c].max()
        o = np.argsort(img[..., c].reshape(-1), kind="stable")
        assert np.all(np.diff(m[..., c].reshape(-1)[o].astype(int)) >= 0)
    with pytest.raises(ValueError):
        match_histograms(img, ref3[..., :2], channel_axis=2)
