"""The C-ABI library builds, loads without a GPU, and exports every symbol include/edgerunner_b200.h declares."""

import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(REPO, 'include', 'edgerunner_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(er_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from edgerunner_b200 import _lib, build
    build.build()
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
        assert s in _lib.SIGNATURES, f'{s} has no ctypes signature'
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.er_version() >= 100


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from core.models import LMM
    from edgerunner_b200 import synth
    model = LMM(synth.tiny_options())
    with pytest.raises(RuntimeError):
        model.generate(torch.zeros(1, 256, 3))


def test_state_dict_schema_matches_reference_spec():
    """LMM.state_dict() keys/shapes == synth.state_dict_spec, which gen_golden.py pinned against the reference's LMM."""
    from core.models import LMM
    from core.options import config_defaults
    from edgerunner_b200 import synth
    opt = synth.tiny_options()
    sd = LMM(opt).state_dict()
    spec = synth.state_dict_spec(opt)
    assert [n for n, _, _ in spec] == list(sd.keys())
    for n, shape, _ in spec:
        assert tuple(sd[n].shape) == tuple(shape), n
    model = LMM(opt)
    model.load_state_dict(synth.synth_state_dict(opt), strict=True)
    assert len(synth.state_dict_spec(config_defaults['ArAE'])) == 416


def test_no_product_import_of_oracle():
    """The product path must never import the oracle (tier rule)."""
    bad = []
    for root in ('core', 'meto', 'edgerunner_b200'):
        for dp, _, files in os.walk(os.path.join(REPO, root)):
            for f in files:
                if f.endswith(('.py', '.cu', '.cpp', '.h', '.cuh')):
                    txt = open(os.path.join(dp, f)).read()
                    if re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M) or re.search(r'#include.*oracle', txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_decode_partition_invariants(tmp_path):
    """The decode kernel's work partition (csrc/decode_partition.h, shared by producer, consumers and the host) checked natively:
    row ranges tile [0, R) on pair boundaries; the KV splits of a head tile its keys / blocks exactly once for every cache length;
    one split owns the new key; the score scratch is large enough for every split handicap er_decode accepts."""
    import subprocess
    exe = str(tmp_path / 'partition_check')
    subprocess.check_call(['g++', '-std=c++17', '-O2', '-o', exe, os.path.join(REPO, 'tests', 'partition_check.cpp')])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith('ok '), out.stdout[-500:]


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: include/edgerunner_b200.h compiles as strict C99 (no C++ / torch types) and a C program links against the
    shared library and calls two entry points that need no GPU."""
    import shutil
    import subprocess
    from edgerunner_b200 import _lib
    if shutil.which('gcc') is None or not os.path.exists(_lib.LIB_PATH):
        pytest.skip('gcc or the built library is missing')
    src = tmp_path / 'abi.c'
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "edgerunner_b200.h"\n'
                   'int main(void) {\n'
                   '    er_dit_config c; memset(&c, 0, sizeof c);\n'
                   '    er_dit* e = 0;\n'
                   '    int rc = er_dit_create(&c, &e);            /* bad dimensions: refused before any CUDA call */\n'
                   '    printf("%d %d %s\\n", er_version(), rc, er_last_error());\n'
                   '    return (er_version() > 0 && rc == ER_ERR_INVALID) ? 0 : 1;\n}\n')
    exe = tmp_path / 'abi'
    inc = os.path.join(REPO, 'include')
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', inc, str(src), '-o', str(exe), _lib.LIB_PATH,
                        '-Wl,-rpath,' + os.path.dirname(_lib.LIB_PATH)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.stdout, out.stderr)
    assert 'bad DiT dimensions' in out.stdout
