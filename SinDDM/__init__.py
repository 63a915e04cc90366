"""Drop-in import path of the reference package (`from SinDDM.models import ...`), backed by sinddm_b200."""
