"""Token <-> mesh helpers at the tail of ``LMM.generate`` and the training tensor contract.

Mirrors ``/root/reference/core/provider.py``: ``save_mesh`` (:39-66), ``tokenize_mesh`` (:69-110),
``detokenize_mesh`` (:112-147) and ``collate_fn`` (:469-541).  The S3-backed datasets of the reference are hard-wired
to private buckets and are out of scope.
"""

import numpy as np
import torch

from core.options import Options

try:
    import trimesh
    _Mesh = trimesh.Trimesh
except Exception:   # trimesh is absent from the target image
    from edgerunner_b200.mesh import SimpleMesh as _Mesh


def detokenize_mesh(tokens, discrete_bins=None, tokenizer=None):
    """tokens [M] (with the +3 special-token offset) -> (vertices, faces); still to be deduplicated."""
    tokens = np.asarray(tokens) - 3
    if tokenizer is not None:
        vertices, faces, _ = tokenizer.decode(tokens)
        return vertices, faces
    # naive 9-tokens-per-face stream (zyx order)
    if len(tokens) % 9 != 0:
        print(f'[WARN] tokens len is {len(tokens)} % 9 != 0, trimming...')
        tokens = tokens[:-(len(tokens) % 9)]
    invalid = (tokens < 0).reshape(-1, 9).any(axis=1)
    coords = tokens.reshape(-1, 3)
    vertices = coords / coords.max() * 2 - 1 if discrete_bins is None else (coords + 0.5) / discrete_bins * 2 - 1
    faces = np.arange(len(vertices)).reshape(-1, 3)[~invalid]
    return vertices[:, [2, 1, 0]], faces


def save_mesh(tokens, opt: Options, path=None, tokenizer=None, clean=True, verbose=False):
    """Single-sequence tokens -> mesh object (or file if ``path``): cut at the first EOS, detokenize, clean."""
    tokens = np.asarray(tokens)
    eos = np.nonzero(tokens == opt.eos_token_id)[0]
    if len(eos) > 0:
        tokens = tokens[:eos[0]]
    vertices, faces = detokenize_mesh(tokens, opt.discrete_bins, tokenizer=tokenizer)
    if verbose:
        print(f'[INFO] vertices: {vertices.shape[0]}, faces: {faces.shape[0]}')
    mesh = _Mesh(vertices=vertices, faces=faces)
    if clean:
        if hasattr(mesh, 'clean_up'):       # SimpleMesh: the same three steps in one native call (er_mesh_clean)
            mesh.clean_up()
        else:
            mesh.merge_vertices()
            mesh.update_faces(mesh.unique_faces())
            mesh.fix_normals()
        if verbose:
            print(f'[INFO] cleaned vertices: {mesh.vertices.shape[0]}, faces: {mesh.faces.shape[0]}')
    if path is None:
        return mesh
    mesh.export(path)


def tokenize_mesh(vertices, faces, discrete_bins, tokenizer=None):
    """(vertices [N,3] in [-1,1], faces [M,3]) -> tokens with the +3 offset."""
    if tokenizer is not None:
        tokens, _, _ = tokenizer.encode(vertices, faces)
        return np.asarray(tokens) + 3
    order = np.lexsort(vertices.T)
    vertices = vertices[order][:, [2, 1, 0]]
    faces = np.argsort(order)[faces]
    start = faces.argmin(axis=1)
    faces = np.take_along_axis(np.concatenate([faces, faces[:, :2]], axis=1), start[:, None] + np.arange(3)[None, :], axis=1)
    faces = np.array(sorted(faces.tolist()))
    coords = ((vertices[faces] + 1) * 0.5 * discrete_bins).clip(0, discrete_bins - 1).astype(np.int32)
    return coords.reshape(-1) + 3


def collate_fn(batch, opt: Options):
    """Dataset items -> the tensor contract of ``LMM.forward`` (reference :469-541).

    Item keys as in the reference datasets: ``cond`` [N,3], ``coords`` (mesh tokens, +3 offset), ``len``, ``num_faces``,
    ``azimuth``, ``path``.  Output: tokens [B,1+M+1] (BOS, body, EOS, pad), labels [B,C+1+M+1] (-100 on cond / BOS / pad),
    masks [B,C+1+M+1] bool, num_tokens = C+1+len+1.  A sequence longer than ``max_seq_length`` is truncated and gets no EOS
    (reference :516-531): a batch of only such rows is one column narrower (no EOS column), exactly as in the reference; in a
    mixed batch the truncated rows are right-padded to the common width (the reference's np.stack would fail on the ragged rows).
    """
    max_len = min(max(item['len'] for item in batch), opt.max_seq_length)
    any_full = any(item['len'] <= max_len for item in batch)
    B, Cn, W = len(batch), opt.num_cond_tokens, max_len + 1 + int(any_full)
    tokens = np.full((B, W), opt.pad_token_id, dtype=np.int64)
    labels = np.full((B, Cn + W), -100, dtype=np.int64)
    masks = np.zeros((B, Cn + W), dtype=bool)
    num_tokens = np.zeros(B, dtype=np.int64)
    for b, item in enumerate(batch):
        body = np.asarray(item['coords'])
        if item['len'] <= max_len:
            seq = np.concatenate([[opt.bos_token_id], body[:item['len']], [opt.eos_token_id]])
        else:
            seq = np.concatenate([[opt.bos_token_id], body[:max_len]])
        tokens[b, :len(seq)] = seq
        labels[b, Cn + 1:Cn + len(seq)] = seq[1:]
        masks[b, :Cn + len(seq)] = True
        num_tokens[b] = Cn + len(seq)
    results = dict(tokens=torch.from_numpy(tokens), labels=torch.from_numpy(labels), masks=torch.from_numpy(masks),
                   num_tokens=torch.from_numpy(num_tokens))
    results['conds'] = torch.from_numpy(np.stack([np.asarray(item['cond']) for item in batch], axis=0)).float()
    results['num_faces'] = torch.from_numpy(np.stack([item['num_faces'] for item in batch], axis=0)).long()
    if 'azimuth' in batch[0]:
        results['azimuths'] = torch.from_numpy(np.stack([item['azimuth'] for item in batch], axis=0)).long()
    results['paths'] = [item.get('path') for item in batch]
    return results
