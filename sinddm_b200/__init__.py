"""sinddm_b200 -- B200-native (sm_100a) implementation of SinDDM's hot path.

Public surface mirrors the reference (fallenshock/SinDDM): SinDDMNet, MultiScaleGaussianDiffusion, EMA,
MultiscaleTrainer, create_img_scales.  All compute goes through libsinddm_b200.so (include/sinddm_b200.h).
"""
from .denoiser import SinDDMConvBlock, SinDDMNet, SinusoidalPosEmb
from .diffusion import EMA, MultiScaleGaussianDiffusion
from .functions import create_img_scales
from .trainer import Dataset, MultiscaleTrainer

__all__ = ["SinDDMNet", "SinDDMConvBlock", "SinusoidalPosEmb", "MultiScaleGaussianDiffusion", "EMA",
           "MultiscaleTrainer", "Dataset", "create_img_scales"]
