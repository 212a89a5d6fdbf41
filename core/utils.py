"""Host utilities on the decode path.  Mirrors ``/root/reference/core/utils.py``.

Only ``quantize_num_faces`` (:109-136), ``get_tokenizer`` (:98-106), ``normalize_mesh`` (:89-95) and the HF
monkey-patch entry point (:138-161) are on or next to the hot path.  ``load_mesh`` needs trimesh (mesh file
I/O, out of scope for the engine) and imports it lazily so that this module loads without it.
"""

import logging

import numpy as np
import torch

from core.options import Options

_FACE_BUCKET_EDGES = (0, 1000, 2000, 4000, 8000)   # bucket b covers (edge[b-1], edge[b]]; > 8000 -> 5


def quantize_num_faces(n):
    """Face-count conditioning bucket: <=0 -> 0 (unconditional), (0,1000] -> 1, (1000,2000] -> 2,
    (2000,4000] -> 3, (4000,8000] -> 4, else 5.  Accepts an int or an integer tensor."""
    if isinstance(n, int):
        return int(np.searchsorted(_FACE_BUCKET_EDGES, n, side='left'))
    edges = torch.tensor(_FACE_BUCKET_EDGES, device=n.device, dtype=n.dtype)
    return torch.bucketize(n, edges, right=False).to(n.dtype)


def get_tokenizer(opt: Options):
    """-> (tokenizer | None, vocab_size).  The tokenizer is ``meto.Engine`` backed by the native C-ABI library."""
    if opt.use_meto:
        from meto import Engine
        tokenizer = Engine(discrete_bins=opt.discrete_bins, backend=opt.meto_backend)
        return tokenizer, tokenizer.num_tokens + 3
    return None, opt.discrete_bins + 3


def normalize_mesh(vertices, bound=0.95):
    lo, hi = vertices.min(0), vertices.max(0)
    return (vertices - (hi + lo) / 2) * (2 * bound / np.max(hi - lo))


def load_mesh(path):
    """Mesh file -> (vertices, faces) through trimesh (scene graphs flattened).  I/O only; not accelerated."""
    import trimesh
    if path.startswith('s3'):
        import megfile
        with megfile.smart_open(path, 'rb') as f:
            data = trimesh.load(file_obj=trimesh.util.wrap_as_stream(f.read()), file_type=path.split('.')[-1])
    else:
        data = trimesh.load(path)
    if isinstance(data, trimesh.Scene):
        parts = []
        for node in data.graph.to_flattened().values():
            geom = data.geometry.get(node['geometry'])
            if isinstance(geom, trimesh.Trimesh):
                parts.append(geom.apply_transform(node['transform']))
        data = trimesh.util.concatenate(parts)
    return data.vertices, data.faces


def init_logger(filename):
    logger = logging.getLogger(__name__)
    logger.setLevel(logging.DEBUG)
    fmt = logging.Formatter('%(asctime)s [%(levelname)s] %(message)s')
    for h in (logging.FileHandler(filename, mode='w'), logging.StreamHandler()):
        h.setLevel(logging.DEBUG)
        h.setFormatter(fmt)
        logger.addHandler(h)
    return logger


def monkey_patch_transformers():
    """The reference patches HF's PrefixConstrainedLogitsProcessor (utils.py:138-161) before generating.
    Here the constraint mask is applied on the device inside the decode kernel (csrc/decode_kernel.cu,
    ``sampler``), so there is nothing to patch; kept so that ``infer.py`` runs unmodified."""
    print('[INFO] edgerunner_b200: grammar constraint runs on-device; no transformers patch needed')
