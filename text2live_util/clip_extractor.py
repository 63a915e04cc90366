"""Stub of text2live_util.clip_extractor: CLIP guidance is outside the sinddm_b200 hot path."""


class ClipExtractor:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("CLIP guidance (ClipExtractor) is outside the sinddm_b200 hot path; "
                                  "it also needs CLIP weights from the network")
