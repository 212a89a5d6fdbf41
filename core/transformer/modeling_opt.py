"""Parameter containers of the OPT-style mesh decoder, in the reference's checkpoint schema.

Mirrors the module tree of ``/root/reference/core/transformer/modeling_opt.py`` (``ShapeOPTConfig`` :86,
``OptFlashAttention2`` :137, ``OPTDecoderLayer`` :239, ``ShapeOPTDecoder`` :307, ``ShapeOPT`` :429) so that
``state_dict()`` / ``load_state_dict()`` / ``.half()`` / ``.to(device)`` behave like the reference's (SURVEY.md
Appendix E).  The modules hold weights only: the arithmetic of prefill, decode step and teacher-forced forward
runs in ``edgerunner_b200/csrc`` (post-LN layers, ReLU MLP, learned absolute positions, untied lm_head).
"""

import math

import torch
from torch import nn


class ShapeOPTConfig:
    model_type = "shape_opt"

    def __init__(self, vocab_size=50272, max_position_embeddings=2048, hidden_dim=1024, intermediate_dim=4096,
                 num_hidden_layers=24, dropout=0.1, attention_dropout=0.0, num_attention_heads=16,
                 activation_function="relu", layerdrop=0.0, init_std=0.02, use_cache=True, pad_token_id=0,
                 bos_token_id=1, eos_token_id=2, enable_bias=True, layer_norm_elementwise_affine=True,
                 num_cond_tokens=257, **kwargs):
        self.vocab_size = vocab_size
        self.max_position_embeddings = max_position_embeddings
        self.hidden_dim = self.hidden_size = hidden_dim
        self.intermediate_dim = intermediate_dim
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.dropout, self.attention_dropout, self.layerdrop = dropout, attention_dropout, layerdrop
        self.activation_function = activation_function
        self.init_std = init_std
        self.use_cache = use_cache
        self.pad_token_id, self.bos_token_id, self.eos_token_id = pad_token_id, bos_token_id, eos_token_id
        self.enable_bias = enable_bias
        self.layer_norm_elementwise_affine = layer_norm_elementwise_affine
        self.num_cond_tokens = num_cond_tokens
        if activation_function != 'relu' or not enable_bias or not layer_norm_elementwise_affine:
            raise NotImplementedError('the B200 decode kernel implements the ArAE decoder: ReLU, biases, affine LayerNorm')


class OptFlashAttention2(nn.Module):
    def __init__(self, config: ShapeOPTConfig, **_):
        super().__init__()
        c = config.hidden_dim
        if c % config.num_attention_heads:
            raise ValueError('hidden_dim must be divisible by num_heads')
        self.num_heads, self.head_dim = config.num_attention_heads, c // config.num_attention_heads
        self.k_proj = nn.Linear(c, c)
        self.v_proj = nn.Linear(c, c)
        self.q_proj = nn.Linear(c, c)
        self.out_proj = nn.Linear(c, c)


class OPTDecoderLayer(nn.Module):
    def __init__(self, config: ShapeOPTConfig, layer_id: int = -1):
        super().__init__()
        self.layer_id = layer_id
        self.self_attn = OptFlashAttention2(config)
        self.self_attn_layer_norm = nn.LayerNorm(config.hidden_dim)
        self.fc1 = nn.Linear(config.hidden_dim, config.intermediate_dim)
        self.fc2 = nn.Linear(config.intermediate_dim, config.hidden_dim)
        self.final_layer_norm = nn.LayerNorm(config.hidden_dim)


class ShapeOPTDecoder(nn.Module):
    def __init__(self, config: ShapeOPTConfig):
        super().__init__()
        self.config = config
        self.embd = nn.Embedding(config.vocab_size, config.hidden_dim, config.pad_token_id)
        self.embed_positions = nn.Embedding(config.max_position_embeddings, config.hidden_dim)
        self.layers = nn.ModuleList([OPTDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.gradient_checkpointing = False

    def gradient_checkpointing_enable(self, *_, **__):   # activation checkpointing is a training-memory device; no-op here
        self.gradient_checkpointing = True


class ShapeOPT(nn.Module):
    def __init__(self, config: ShapeOPTConfig):
        super().__init__()
        self.config = config
        self.model = ShapeOPTDecoder(config)
        self.lm_head = nn.Linear(config.hidden_dim, config.vocab_size, bias=False)
        self.apply(self._init_weights)
        # GPT-2 style scaled init of the residual projections (reference :444-446)
        for name, p in self.named_parameters():
            if name.endswith('out_proj.weight'):
                nn.init.normal_(p, mean=0.0, std=0.02 / math.sqrt(2 * config.num_hidden_layers))

    def _init_weights(self, module):
        std = self.config.init_std
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()

    def generate(self, *a, **k):
        raise RuntimeError('ShapeOPT.generate: the token loop runs on-device; call LMM.generate (core/models.py)')
