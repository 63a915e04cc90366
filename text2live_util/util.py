"""Stub of text2live_util.util: CLIP guidance is outside the sinddm_b200 hot path."""


def get_augmentations_template():
    raise NotImplementedError("CLIP guidance (text2live_util) is outside the sinddm_b200 hot path")
