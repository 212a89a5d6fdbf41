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


def read_image(path, mode='float', order='RGB'):
    """kiui.read_image as infer_dit.py:82 uses it (mode='uint8', order='RGBA')."""
    from PIL import Image
    img = Image.open(path)
    img = img.convert('RGBA' if order == 'RGBA' and img.mode == 'RGBA' else 'RGB')
    arr = np.asarray(img)
    return arr if mode == 'uint8' else arr.astype(np.float32) / 255.0


def write_image(path, img, order='RGB'):
    from PIL import Image
    arr = np.asarray(img)
    if arr.dtype != np.uint8:
        arr = (np.clip(arr, 0, 1) * 255).astype(np.uint8)
    Image.fromarray(arr).save(path)
