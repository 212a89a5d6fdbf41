"""TEST STUB of `trimesh` (absent from the image): the handful of calls the reference's infer.py / core.utils.load_mesh make on the
point-cloud path — load a triangle .obj, area-weighted surface sampling, point-cloud / mesh export.  Mesh clean-up (merge_vertices,
unique_faces, fix_normals) is inherited from edgerunner_b200.mesh.SimpleMesh, the holder the product uses when trimesh is missing."""
import numpy as np

from edgerunner_b200.mesh import SimpleMesh


class Trimesh(SimpleMesh):
    def sample(self, count):
        v, f = self.vertices, self.faces
        a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
        area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
        idx = np.random.choice(len(f), size=count, p=area / area.sum())
        r1, r2 = np.sqrt(np.random.rand(count, 1)), np.random.rand(count, 1)
        return (1 - r1) * a[idx] + r1 * (1 - r2) * b[idx] + r1 * r2 * c[idx]


class Scene:
    pass


class PointCloud:
    def __init__(self, vertices):
        self.vertices = np.asarray(vertices, dtype=np.float64)

    def export(self, path):
        with open(path, 'w') as fh:
            for p in self.vertices:
                fh.write('v %.17g %.17g %.17g\n' % tuple(p))


def load(path, **k):
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if t[0] == 'v':
                vs.append([float(x) for x in t[1:4]])
            elif t[0] == 'f':
                fs.append([int(x.split('/')[0]) - 1 for x in t[1:4]])
    return Trimesh(vertices=np.asarray(vs), faces=np.asarray(fs))


class util:
    @staticmethod
    def concatenate(parts):
        raise NotImplementedError

    @staticmethod
    def wrap_as_stream(b):
        raise NotImplementedError
