"""`SinDDM.models` of the reference -> sinddm_b200 (denoiser + diffusion)."""
from sinddm_b200.denoiser import SinDDMConvBlock, SinDDMNet, SinusoidalPosEmb  # noqa: F401
from sinddm_b200.diffusion import EMA, MultiScaleGaussianDiffusion  # noqa: F401
