"""TEST STUB of kiui.op: `recenter` as infer_dit.py:89 calls it (image pre-processing, not on the measured path): crop to the mask's
bounding box and paste it, scaled to leave `border_ratio` of margin, into a blank canvas of the original size."""
import numpy as np


def recenter(image, mask, border_ratio=0.2):
    from PIL import Image
    H, W, C = image.shape
    ys, xs = np.nonzero(mask)
    if len(ys) == 0:
        return image
    y0, y1, x0, x1 = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
    crop = image[y0:y1, x0:x1]
    size = int(min(H, W) * (1 - border_ratio))
    s = size / max(crop.shape[0], crop.shape[1])
    h2, w2 = max(1, int(round(crop.shape[0] * s))), max(1, int(round(crop.shape[1] * s)))
    small = np.asarray(Image.fromarray(crop).resize((w2, h2), Image.BILINEAR))
    out = np.zeros_like(image)
    oy, ox = (H - h2) // 2, (W - w2) // 2
    out[oy:oy + h2, ox:ox + w2] = small
    return out
