"""TEST STUB of the `kiui` package (absent from the image) — only what the reference's infer.py touches at import time and on the
point-cloud path: seed_everything, lo, op.recenter (image path, unused), mesh_utils.clean_mesh (identity for an already clean mesh)."""
import random

import numpy as np
import torch


def seed_everything(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def lo(*a, **k):
    pass
