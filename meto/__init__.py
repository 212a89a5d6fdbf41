"""``meto`` — mesh tokenizer package, drop-in for ``/root/reference/meto/meto/__init__.py``.

``Engine(discrete_bins, verbose=False, backend='LR_ABSCO')`` keeps the reference's surface (:21-50):
``decode(tokens[N] int) -> (vertices float64 [V,3], faces int [F,3], face_type)`` and the ``num_tokens`` /
``num_base_tokens`` / ``num_special_tokens`` attributes.  The implementation is the native C-ABI function
``er_meto_decode`` of libedgerunner_b200 (no pybind, no per-element Python objects).
"""

import ctypes as C
from typing import Literal

import numpy as np

from edgerunner_b200 import _lib


class Engine:
    def __init__(self, discrete_bins, verbose=False, backend: Literal['CLERS', 'LR', 'LR_ABSCO'] = 'LR_ABSCO'):
        if backend != 'LR_ABSCO':
            raise NotImplementedError(f"meto backend '{backend}': only LR_ABSCO (the ArAE/DiT preset backend) is on the B200 path")
        self.discrete_bins = int(discrete_bins)
        self.verbose = verbose
        self.backend = backend
        self.num_base_tokens = self.discrete_bins
        self.num_special_tokens = 3
        self.num_tokens = self.num_base_tokens + self.num_special_tokens
        self._lib = _lib.load()

    def decode(self, tokens):
        tok = np.ascontiguousarray(np.asarray(tokens).reshape(-1), dtype=np.int32)
        n = tok.shape[0]
        cap = n // 4 + 3
        verts = np.empty((3 * cap, 3), dtype=np.float32)
        faces = np.empty((cap, 3), dtype=np.int32)
        ftype = np.empty(cap, dtype=np.int32)
        nv, nf, nt = C.c_int64(), C.c_int64(), C.c_int64()
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        _lib.check(self._lib.er_meto_decode(self.discrete_bins, p(tok, C.c_int32), n, p(verts, C.c_float), p(faces, C.c_int32),
                                            p(ftype, C.c_int32), C.byref(nv), C.byref(nf), C.byref(nt)))
        # the reference returns np.asarray of Python floats: float64 holding float32-valued numbers
        return verts[:nv.value].astype(np.float64), faces[:nf.value].astype(np.int64), ftype[:nt.value].astype(np.int64)

    def encode(self, vertices, faces):
        raise NotImplementedError('meto encode (training-data side) is the next row of the scope table (SURVEY.md §8f.1)')


def normalize_mesh(vertices, bound=0.95):
    lo, hi = vertices.min(0), vertices.max(0)
    return (vertices - (hi + lo) / 2) * (2 * bound / np.max(hi - lo))
