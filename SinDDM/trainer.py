"""`SinDDM.trainer` of the reference -> sinddm_b200.trainer."""
from sinddm_b200.trainer import Dataset, MultiscaleTrainer  # noqa: F401
