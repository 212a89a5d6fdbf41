"""Minimal triangle-mesh holder used when ``trimesh`` is not installed (it is absent from the target image).

Covers what ``save_mesh`` / ``infer.py`` touch: ``vertices`` / ``faces``, ``merge_vertices``, ``unique_faces`` +
``update_faces``, ``fix_normals`` (winding propagated over shared edges inside each connected component, then every component
with negative signed volume is inverted — the two steps of trimesh's ``repair.fix_normals``) and ``export`` to .ply/.obj.
Host-side post-processing only; the reference delegates this to third-party trimesh (parity unpinned).
"""

import numpy as np


class SimpleMesh:
    def __init__(self, vertices=None, faces=None, **_):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)

    def merge_vertices(self):
        """Vertices equal to 8 decimals become one; the survivors keep the order of their first occurrence (as in trimesh)."""
        if len(self.vertices) == 0:
            return
        _, first, inv = np.unique(np.round(self.vertices, 8), axis=0, return_index=True, return_inverse=True)
        rank = np.argsort(np.argsort(first))                 # unique id (sorted order) -> position by first occurrence
        self.vertices = self.vertices[np.sort(first)]
        self.faces = rank[inv.reshape(-1)][self.faces]

    def clean_up(self):
        """merge_vertices + update_faces(unique_faces()) + fix_normals in one native call (``er_mesh_clean``)."""
        import ctypes as C
        from edgerunner_b200 import _lib
        lib = _lib.load()
        v = np.ascontiguousarray(self.vertices, dtype=np.float64)
        f = np.ascontiguousarray(self.faces, dtype=np.int32)
        vo, fo = np.empty_like(v), np.empty_like(f)
        nv, nf = C.c_int64(), C.c_int64()
        P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        _lib.check(lib.er_mesh_clean(P(v, C.c_double), len(v), P(f, C.c_int32), len(f), 8, P(vo, C.c_double), P(fo, C.c_int32), C.byref(nv), C.byref(nf)))
        self.vertices, self.faces = vo[:nv.value].copy(), fo[:nf.value].astype(np.int64)

    def unique_faces(self):
        if len(self.faces) == 0:
            return np.zeros(0, dtype=bool)
        key = np.sort(self.faces, axis=1)
        _, first = np.unique(key, axis=0, return_index=True)
        mask = np.zeros(len(self.faces), dtype=bool)
        mask[first] = True
        return mask

    def update_faces(self, mask):
        self.faces = self.faces[mask]

    def fix_normals(self):
        """Make winding consistent inside each edge-connected component (BFS over shared edges)."""
        f = self.faces
        if len(f) == 0:
            return
        edges = {}
        for i, (a, b, c) in enumerate(f):
            for u, v in ((a, b), (b, c), (c, a)):
                edges.setdefault((min(u, v), max(u, v)), []).append((i, u, v))
        seen = np.zeros(len(f), dtype=bool)
        comp = np.zeros(len(f), dtype=np.int64)
        ncomp = 0
        for start in range(len(f)):
            if seen[start]:
                continue
            seen[start] = True
            comp[start] = ncomp
            stack = [start]
            while stack:
                i = stack.pop()
                a, b, c = self.faces[i]
                for u, v in ((a, b), (b, c), (c, a)):
                    for j, uu, vv in edges[(min(u, v), max(u, v))]:
                        if j == i or seen[j]:
                            continue
                        # consistent neighbours traverse the shared edge in opposite directions
                        ja, jb, jc = self.faces[j]
                        dir_j = [(ja, jb), (jb, jc), (jc, ja)]
                        if (u, v) in dir_j:
                            self.faces[j] = self.faces[j][::-1]
                        seen[j] = True
                        comp[j] = comp[i]
                        stack.append(j)
            ncomp += 1
        # outward orientation: signed volume of each component (sum of v0 . (v1 x v2) / 6); negative -> invert that component
        tri = self.vertices[self.faces]
        vol6 = np.einsum('ij,ij->i', tri[:, 0], np.cross(tri[:, 1], tri[:, 2]))
        flip = np.bincount(comp, weights=vol6, minlength=ncomp) < 0
        sel = flip[comp]
        self.faces[sel] = self.faces[sel][:, ::-1]

    @property
    def volume(self):
        tri = self.vertices[self.faces]
        return float(np.einsum('ij,ij->i', tri[:, 0], np.cross(tri[:, 1], tri[:, 2])).sum() / 6.0)

    def export(self, path):
        ext = str(path).rsplit('.', 1)[-1].lower()
        with open(path, 'w') as fh:
            if ext == 'ply':
                fh.write('ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n'
                         'element face %d\nproperty list uchar int vertex_indices\nend_header\n' % (len(self.vertices), len(self.faces)))
                for v in self.vertices:
                    fh.write('%.8g %.8g %.8g\n' % tuple(v))
                for t in self.faces:
                    fh.write('3 %d %d %d\n' % tuple(t))
            else:
                for v in self.vertices:
                    fh.write('v %.8g %.8g %.8g\n' % tuple(v))
                for t in self.faces:
                    fh.write('f %d %d %d\n' % tuple(t + 1))
