"""``meto`` — mesh tokenizer package, drop-in for ``/root/reference/meto/meto/__init__.py``.

``Engine(discrete_bins, verbose=False, backend='LR_ABSCO' | 'LR' | 'CLERS')`` keeps the reference's surface (:21-50):
``encode(vertices [V,3], faces [F,3]) -> (tokens, face_order, face_type)``,
``decode(tokens[N] int) -> (vertices float64 [V,3], faces int [F,3], face_type)`` and the ``num_tokens`` /
``num_base_tokens`` / ``num_special_tokens`` attributes.  The implementation is the pair of native C-ABI functions
``er_meto_encode`` / ``er_meto_decode`` of libedgerunner_b200 (no pybind, no per-element Python objects).
"""

import ctypes as C
from typing import Literal

import numpy as np

from edgerunner_b200 import _lib


class Engine:
    def __init__(self, discrete_bins, verbose=False, backend: Literal['CLERS', 'LR', 'LR_ABSCO'] = 'LR_ABSCO'):
        if backend not in ('LR_ABSCO', 'LR', 'CLERS'):
            raise NotImplementedError(f"meto backend '{backend}': LR_ABSCO, LR and CLERS are provided")
        self.discrete_bins = int(discrete_bins)
        self.verbose = verbose
        self.backend = backend
        self._backend_id = {'LR_ABSCO': 0, 'LR': 1, 'CLERS': 2}[backend]          # ER_METO_LR_ABSCO / ER_METO_LR / ER_METO_CLERS
        self.num_base_tokens = self.discrete_bins * (1 if backend == 'LR_ABSCO' else 2)
        self.num_special_tokens = 7 if backend == 'CLERS' else 3                   # reference meto/meto/__init__.py:26-37
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
        _lib.check(self._lib.er_meto_decode(self._backend_id, self.discrete_bins, p(tok, C.c_int32), n, p(verts, C.c_float), p(faces, C.c_int32),
                                            p(ftype, C.c_int32), C.byref(nv), C.byref(nf), C.byref(nt)))
        # the reference returns np.asarray of Python floats: float64 holding float32-valued numbers
        return verts[:nv.value].astype(np.float64), faces[:nf.value].astype(np.int64), ftype[:nt.value].astype(np.int64)

    def encode(self, vertices, faces):
        """Mesh -> (tokens, face_order, face_type) as int arrays, reference meto/meto/__init__.py:40-45.
        vertices are expected in [-1, 1] (normalize_mesh), faces are triangles of vertex indices."""
        v = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 3))   # pybind narrows to float the same way
        f = np.ascontiguousarray(np.asarray(faces).reshape(-1, 3), dtype=np.int32)
        nv, nf = v.shape[0], f.shape[0]
        if nf and (f.min() < 0 or f.max() >= nv):
            raise ValueError('meto.encode: face index out of range')
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        tcap, fcap = max(12 * nf, 1), max(nf, 1)                       # always enough for LR_ABSCO; LR / CLERS may repeat faces: retry once
        for _ in range(2):
            tok = np.empty(tcap, dtype=np.int32)
            order = np.empty(fcap, dtype=np.int32)
            ftype = np.empty(fcap, dtype=np.int32)
            nt, ne = C.c_int64(), C.c_int64()
            rc = self._lib.er_meto_encode(self._backend_id, self.discrete_bins, p(v, C.c_float), nv, p(f, C.c_int32), nf, p(tok, C.c_int32), tcap,
                                          p(order, C.c_int32), p(ftype, C.c_int32), fcap, C.byref(nt), C.byref(ne))
            if rc != -4:                                                # ER_ERR_CAPACITY: the needed sizes are in nt / ne
                break
            tcap, fcap = nt.value, ne.value
        _lib.check(rc)
        return tok[:nt.value].astype(np.int64), order[:ne.value].astype(np.int64), ftype[:ne.value].astype(np.int64)


def normalize_mesh(vertices, bound=0.95):
    lo, hi = vertices.min(0), vertices.max(0)
    return (vertices - (hi + lo) / 2) * (2 * bound / np.max(hi - lo))
