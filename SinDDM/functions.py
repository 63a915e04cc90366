"""`SinDDM.functions` of the reference -> sinddm_b200.functions."""
from sinddm_b200.functions import *  # noqa: F401,F403
from sinddm_b200.functions import (APEX_AVAILABLE, cosine_beta_schedule, create_img_scales, cycle, default,  # noqa: F401
                                   exists, extract, loss_backwards, noise_like, num_to_groups)
