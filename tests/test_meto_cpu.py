"""meto detokenizer: native C-ABI implementation and the C oracle against the compiled-reference goldens (CPU)."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def oracle_lib():
    path = os.path.join(REPO, 'oracle', 'libmeto_oracle.so')
    if not os.path.exists(path):
        subprocess.check_call(['make', '-C', os.path.join(REPO, 'oracle'), 'oracle'])
    return C.CDLL(path)


def oracle_decode(lib, bins, tok):
    tok = np.ascontiguousarray(tok, dtype=np.int32)
    n = len(tok)
    cap = n // 4 + 3
    v = np.empty((3 * cap, 3), np.float32); f = np.empty((cap, 3), np.int32); t = np.empty(cap, np.int32)
    nv, nf, nt = C.c_int64(), C.c_int64(), C.c_int64()
    P = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    lib.meto_oracle_decode(C.c_int(bins), P(tok, C.c_int32), C.c_int64(n), P(v, C.c_float), P(f, C.c_int32), P(t, C.c_int32),
                           C.byref(nv), C.byref(nf), C.byref(nt))
    return v[:nv.value].astype(np.float64), f[:nf.value], t[:nt.value]


def cases(golden_dir):
    g = np.load(os.path.join(golden_dir, 'meto.npz'))
    for key in g['names']:
        key = str(key)
        yield key, int(key.rsplit('_', 1)[1]), g[key + '_tokens'], g[key + '_dv'], g[key + '_df'], g[key + '_dt']
    for i in range(int(g['n_streams'])):
        yield f'stream{i}', 512, g[f'stream{i}_tokens'], g[f'stream{i}_dv'], g[f'stream{i}_df'], g[f'stream{i}_dt']


def test_oracle_matches_reference(golden_dir, oracle_lib):
    n = 0
    for name, bins, tok, dv, df, dt in cases(golden_dir):
        v, f, t = oracle_decode(oracle_lib, bins, tok)
        np.testing.assert_array_equal(v, dv.reshape(-1, 3), err_msg=name)      # bit-exact (float32 values)
        np.testing.assert_array_equal(f, df.reshape(-1, 3), err_msg=name)
        np.testing.assert_array_equal(t, dt, err_msg=name)
        n += 1
    assert n >= 30


def test_native_matches_reference(golden_dir):
    from meto import Engine
    for name, bins, tok, dv, df, dt in cases(golden_dir):
        v, f, t = Engine(bins).decode(tok)
        assert v.dtype == np.float64
        np.testing.assert_array_equal(v, dv.reshape(-1, 3), err_msg=name)
        np.testing.assert_array_equal(f, df.reshape(-1, 3), err_msg=name)
        np.testing.assert_array_equal(t, dt, err_msg=name)


def test_cube_known_answer():
    """KAT recorded from the compiled reference in SURVEY.md §8c (cube scaled to +-0.95, bins 512)."""
    from meto import Engine
    tok = [2, 502, 15, 502, 502, 15, 15, 15, 15, 15, 0, 15, 15, 502, 0, 15, 502, 15, 0, 502, 502, 15, 0, 502, 15, 15, 1, 502, 502, 502,
           0, 502, 15, 502, 1, 15, 15, 502, 1, 15, 502, 502, 0, 15, 502, 15, 1, 502, 502, 15, 1, 502, 502, 502]
    v, f, t = Engine(512).decode(np.array(tok))   # _meto alphabet (already -3): 2 = BOM, coords +3
    assert v.shape == (14, 3) and f.shape == (12, 3)
    assert t.tolist() == [0, 0, 0, 0, 1, 0, 1, 1, 0, 1, 1, 2]
    assert v.max() == 0.951171875


def test_save_mesh_tail(golden_dir):
    """provider.save_mesh: EOS cut + detokenize (core/provider.py:39-66) on the tiny golden token stream."""
    from core.provider import save_mesh
    from edgerunner_b200 import synth
    from meto import Engine
    g = np.load(os.path.join(golden_dir, 'tiny.npz'))
    opt = synth.tiny_options()
    mesh = save_mesh(g['greedy_tokens'], opt, tokenizer=Engine(opt.discrete_bins), clean=False)
    np.testing.assert_array_equal(np.asarray(mesh.vertices), g['mesh_vertices'])
    np.testing.assert_array_equal(np.asarray(mesh.faces), g['mesh_faces'])
