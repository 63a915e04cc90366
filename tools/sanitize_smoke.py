"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck): training steps at two ragged scales in
both math modes, the fused optimizer step, a sampling chain with and without graph replay.
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

from sinddm_b200 import MultiScaleGaussianDiffusion, MultiscaleTrainer, SinDDMNet  # noqa: E402

dev = "cuda:0"
sizes = [(37, 29), (53, 41)]          # (W, H): not multiples of any tile size
for math in ("tf32", "fp32"):
    torch.manual_seed(0)
    net = SinDDMNet(dim=160, multiscale=True, device=dev, math=math).to(dev)
    dif = MultiScaleGaussianDiffusion(denoise_fn=net, n_scales=2, scale_factor=1.4, image_sizes=sizes, timesteps=12,
                                      train_full_t=True, scale_losses=[0.9], device=dev,
                                      results_folder=tempfile.mkdtemp()).to(dev)
    rs = np.random.RandomState(0)
    pyr = [(Image.fromarray(rs.randint(0, 255, (h, w, 3)).astype(np.uint8)),) * 2 for (w, h) in sizes]
    tr = MultiscaleTrainer(dif, None, n_scales=2, image_sizes=sizes, train_batch_size=3, train_lr=1e-3,
                           gradient_accumulate_every=1, step_start_ema=1, update_ema_every=1, avg_window=2,
                           results_folder=tempfile.mkdtemp(), device=dev, pyramid=pyr)
    tr.train_num_steps = 4
    tr.train()
    for graph in (False, True):
        tr.ema_model.use_step_graph = graph
        out = tr.sample_scales(scale_mul=(1, 1), batch_size=2, save_images=False, custom_sample=True)
        assert torch.isfinite(out[-1]).all()
    torch.cuda.synchronize()
    print("ok", math, float(out[-1].abs().mean()))
