"""TEST STUB of `rembg` (absent from the image): the reference's infer_dit.py creates a session at import time and removes the background of
RGB inputs.  The test feeds an RGBA image, so `remove` only has to exist; it returns the input with an opaque alpha channel."""
import numpy as np


def new_session(*a, **k):
    return None


def remove(image, session=None, **k):
    image = np.asarray(image)
    if image.shape[-1] == 4:
        return image
    return np.concatenate([image, np.full(image.shape[:2] + (1,), 255, dtype=image.dtype)], axis=-1)
