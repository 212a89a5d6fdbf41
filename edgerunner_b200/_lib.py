"""ctypes binding of include/edgerunner_b200.h.  Fails loudly: no library -> ImportError-like RuntimeError,
no CUDA device -> the first engine call returns an error; there is no CPU fallback."""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libedgerunner_b200.so')

c_i32, c_i64, c_u64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_void_p


class ErConfig(C.Structure):
    _fields_ = [(n, c_i32) for n in (
        'device', 'hidden_dim', 'num_heads', 'num_layers', 'ffn_dim', 'vocab_size', 'max_positions', 'num_cond_tokens',
        'use_num_face_cond', 'bos_token_id', 'eos_token_id', 'pad_token_id', 'has_point_encoder', 'point_hidden_dim',
        'point_num_heads', 'point_latent_size', 'point_latent_dim', 'max_seq_rows', 'max_points', 'max_tf_rows')]


class ErDitConfig(C.Structure):
    _fields_ = [(n, c_i32) for n in ('device', 'hidden_dim', 'num_heads', 'num_layers', 'latent_size', 'latent_dim', 'cond_tokens', 'cond_dim')]


# name -> (restype, argtypes); must list every symbol the header declares (tests/test_abi_cpu.py checks this)
SIGNATURES = {
    'er_last_error': (C.c_char_p, []),
    'er_version': (C.c_int, []),
    'er_create': (C.c_int, [C.POINTER(ErConfig), C.POINTER(c_vp)]),
    'er_destroy': (None, [c_vp]),
    'er_load_weight': (C.c_int, [c_vp, C.c_char_p, c_vp, c_i32, C.POINTER(c_i64), c_i32, c_vp]),
    'er_finalize_weights': (C.c_int, [c_vp, c_vp]),
    'er_encode_cond': (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    'er_prefill': (C.c_int, [c_vp, C.POINTER(c_i32), c_i32, c_vp]),
    'er_decode': (C.c_int, [c_vp, c_i32, c_i32, c_i32, c_u64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'er_generate_host': (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_u64, c_i32, c_vp, c_vp]),
    'er_forward_tf': (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, C.POINTER(c_i32), c_i32, c_i32, c_f32, c_vp, c_vp, c_vp]),
    'er_forward_tf2': (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, C.POINTER(c_i32), c_i32, c_i32, c_f32, c_vp, c_vp, c_vp, c_vp]),
    'er_attention_bnhd': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    'er_weight_bytes_per_token': (c_i64, [c_vp]),
    'er_kv_bytes_per_row': (c_i64, [c_vp]),
    'er_cache_rows': (c_i32, [c_vp]),
    'er_kernel_launches': (c_i64, [c_vp]),
    'er_debug_set': (C.c_int, [c_vp, C.c_char_p, c_i64]),
    'er_debug_phase_timeline': (C.c_int, [c_vp, c_i32, c_i32]),
    'er_debug_read_timeline': (C.c_int, [c_vp, C.POINTER(c_u64), c_i32]),
    'er_dit_create': (C.c_int, [C.POINTER(ErDitConfig), C.POINTER(c_vp)]),
    'er_dit_destroy': (None, [c_vp]),
    'er_dit_load_weight': (C.c_int, [c_vp, C.c_char_p, c_vp, c_i32, c_i64, c_vp]),
    'er_dit_finalize_weights': (C.c_int, [c_vp, c_vp]),
    'er_dit_cond': (C.c_int, [c_vp, c_vp, c_i32, c_vp, c_vp]),
    'er_dit_forward': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp]),
    'er_dit_run': (C.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, C.POINTER(c_f32), C.POINTER(c_f32), c_f32, c_i32, c_i32, c_vp]),
    'er_dit_run_host': (C.c_int, [c_vp, C.POINTER(c_f32), C.POINTER(c_f32), c_i32, c_i32, C.POINTER(c_f32), C.POINTER(c_f32), c_f32, c_i32, c_i32]),
    'er_dit_kernel_launches': (c_i64, [c_vp]),
    'er_dit_flops_per_forward': (C.c_double, [c_vp, c_i32]),
    'er_dit_debug_set': (C.c_int, [c_vp, C.c_char_p, c_i64]),
    'er_train_step': (C.c_int, [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, C.POINTER(c_i32), c_i32, c_i32, c_f32, c_f32, c_u64, c_f32, c_i32, c_vp, c_vp, c_vp]),
    'er_grad_get': (C.c_int, [c_vp, C.c_char_p, c_vp, c_i64, c_vp]),
    'er_grad_has': (c_i32, [c_vp, C.c_char_p]),
    'er_attention_bwd_bnhd': (C.c_int, [c_vp] * 8 + [c_i32] * 6 + [c_vp]),
    'er_grad_norm_clip': (C.c_int, [c_vp, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp]),
    'er_adamw_step': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_i32, c_vp, c_vp]),
    'er_meto_decode': (C.c_int, [c_i32, c_i32, C.POINTER(c_i32), c_i64, C.POINTER(c_f32), C.POINTER(c_i32), C.POINTER(c_i32),
                                 C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_i64)]),
    'er_meto_encode': (C.c_int, [c_i32, c_i32, C.POINTER(c_f32), c_i64, C.POINTER(c_i32), c_i64, C.POINTER(c_i32), c_i64,
                                 C.POINTER(c_i32), C.POINTER(c_i32), c_i64, C.POINTER(c_i64), C.POINTER(c_i64)]),
    'er_mesh_clean': (C.c_int, [C.POINTER(C.c_double), c_i64, C.POINTER(c_i32), c_i64, c_i32, C.POINTER(C.c_double), C.POINTER(c_i32),
                                C.POINTER(c_i64), C.POINTER(c_i64)]),
}

_lib = None


def load():
    """Load (once) the in-tree CUDA library.  Raises if it has not been built: the product never degrades to a CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'edgerunner_b200: {LIB_PATH} is missing - run `python -m edgerunner_b200.build` '
                               '(there is no CPU / PyTorch fallback for this path)')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class ErError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise ErError(f'edgerunner_b200 error {rc}: {load().er_last_error().decode()}')
