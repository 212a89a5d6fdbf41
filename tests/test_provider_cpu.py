"""core.provider host logic against outputs of the reference's own provider functions executed in the build container
(oracle/gen_golden.py --only provider -> tests/golden/provider.npz): tokenize_mesh (naive + meto), detokenize_mesh (naive), collate_fn."""

import os

import numpy as np
import pytest

import meshes
from core.provider import collate_fn, detokenize_mesh, tokenize_mesh
from edgerunner_b200 import synth


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'provider.npz'))


def _items(rng, opt, lens):            # same construction as oracle/gen_golden.py::provider_items
    items = []
    for i, n in enumerate(lens):
        items.append(dict(cond=rng.uniform(-0.95, 0.95, (opt.point_num, 3)).astype(np.float32),
                          coords=rng.randint(3, opt.discrete_bins + 3, size=n).astype(np.int64), len=int(n),
                          num_faces=int(rng.randint(10, 9000)), azimuth=int(rng.randint(0, 360)), path=f'item{i}'))
    return items


@pytest.mark.parametrize('name', ['cube', 'torus', 'icosphere', 'random_soup'])
def test_tokenize_and_detokenize_match_reference(gold, name):
    from meto import Engine
    v, f = meshes.all_meshes()[name]
    v = v.astype(np.float64)
    naive = tokenize_mesh(v, f, 512, tokenizer=None)
    np.testing.assert_array_equal(naive, gold[f'tok_naive_{name}'])
    np.testing.assert_array_equal(tokenize_mesh(v, f, 512, tokenizer=Engine(512)), gold[f'tok_meto_{name}'])
    dv, df = detokenize_mesh(naive, 512, tokenizer=None)
    np.testing.assert_array_equal(np.asarray(dv, dtype=np.float64), gold[f'detok_naive_v_{name}'])
    np.testing.assert_array_equal(df, gold[f'detok_naive_f_{name}'])


@pytest.mark.parametrize('tag', ['plain', 'trunc'])
def test_collate_fn_matches_reference(gold, tag):
    opt = synth.tiny_options()
    lens = [int(x) for x in gold[f'collate_{tag}_lens']]
    res = collate_fn(_items(np.random.RandomState(5), opt, lens), opt)
    for k in ('conds', 'num_faces', 'num_tokens', 'azimuths', 'labels', 'masks'):
        np.testing.assert_array_equal(res[k].numpy(), gold[f'collate_{tag}_{k}'], err_msg=k)
    ref_tok = gold[f'collate_{tag}_tokens']
    ours = res['tokens'].numpy()
    # truncated rows carry no EOS: the reference row is one column shorter; ours is right-padded to the common width
    np.testing.assert_array_equal(ours[:, :ref_tok.shape[1]], ref_tok)
    assert (ours[:, ref_tok.shape[1]:] == opt.pad_token_id).all()
    assert res['tokens'].dtype == res['labels'].dtype and str(res['masks'].dtype) == 'torch.bool'
    assert res['paths'] == [f'item{i}' for i in range(len(lens))]


def test_collate_fn_mixed_batch_is_padded():
    """A batch mixing a truncated and an un-truncated sequence makes the reference's np.stack raise; here the short row is padded."""
    opt = synth.tiny_options()
    lens = [opt.max_seq_length + 7, 21]
    res = collate_fn(_items(np.random.RandomState(9), opt, lens), opt)
    C = opt.num_cond_tokens
    assert res['tokens'].shape == (2, opt.max_seq_length + 2) and res['labels'].shape == (2, C + opt.max_seq_length + 2)
    assert int(res['num_tokens'][0]) == C + 1 + opt.max_seq_length and int(res['num_tokens'][1]) == C + 1 + 21 + 1
    assert (res['tokens'][0] != opt.eos_token_id).all() and int(res['tokens'][1, 22]) == opt.eos_token_id
    assert int(res['masks'][0].sum()) == C + 1 + opt.max_seq_length and int(res['masks'][1].sum()) == C + 23
