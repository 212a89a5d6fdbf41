"""Configuration surface of the mesh-token decode path.

Mirrors the reference's CLI contract (``/root/reference/core/options.py:17-213``): the
``Options`` dataclass field names and defaults ARE the command line of ``infer.py`` /
``main.py`` (tyro sub-commands ``default`` / ``ArAE`` / ``DiT``), so they are kept name-for-name
and default-for-default.  Only the fields read on the hot path matter to the B200 engine
(see ``edgerunner_b200.engine.EngineConfig.from_options``); the rest are carried so that a
reference script parses the same flags.
"""

from dataclasses import dataclass, replace
from typing import Dict, Literal, Optional, Tuple

try:  # tyro is only needed for the CLI sub-command type; the engine itself never imports it
    import tyro
except Exception:  # pragma: no cover - tyro is present in the target image
    tyro = None


@dataclass
class Options:
    # ---- tokenizer (meto) -------------------------------------------------------------
    discrete_bins: int = 512                 # coordinate bins == number of coordinate tokens
    use_meto: bool = True                    # EdgeBreaker-style tokenizer on/off
    meto_backend: Literal['LR', 'LR_ABSCO'] = 'LR_ABSCO'
    bos_token_id: int = 1
    eos_token_id: int = 2
    pad_token_id: int = 0

    # ---- point-cloud encoder ------------------------------------------------------------
    point_num: int = 8192                    # points sampled per cloud
    point_hidden_dim: int = 1024
    point_num_heads: int = 16
    point_latent_size: int = 2048            # number of latent (query) tokens
    point_latent_dim: int = 64
    point_num_layers: int = 24
    point_query_num: int = 81920
    point_encoder_mode: Literal['downsample', 'embed'] = 'embed'
    kl_weight: float = 1e-8

    # ---- DiT (image-conditioned latent diffusion; out of scope for the engine) -----------
    dit_hidden_dim: int = 1024
    dit_num_heads: int = 16
    dit_num_layers: int = 24
    snr_gamma: Optional[float] = 5.0
    noise_scheduler_predtype: Literal["epsilon", "v_prediction"] = "v_prediction"

    # ---- auto-regressive mesh decoder ----------------------------------------------------
    freeze_encoder: bool = True
    max_seq_length: int = 10240              # mesh tokens only (no BOS/EOS/COND)
    hidden_dim: int = 1024
    intermediate_dim: Optional[int] = None   # None -> 4 * hidden_dim
    num_layers: int = 24
    num_heads: int = 16
    cond_mode: Literal['none', 'image', 'point', 'point_latent'] = 'image'
    num_cond_tokens: int = 257
    generate_mode: Literal['greedy', 'sample'] = 'sample'
    use_num_face_cond: bool = False
    nof_dropout_ratio: float = 0.2

    # ---- dataset ------------------------------------------------------------------------
    max_face_length: int = 1000
    dataset: Literal['obj', 'objxl'] = 'obj'
    num_workers: int = 64
    testset_size: int = 32
    use_decimate_aug: bool = True
    use_scale_aug: bool = True

    # ---- training -----------------------------------------------------------------------
    workspace: str = './workspace'
    resume: Optional[str] = None
    resume2: Optional[str] = None
    resume_step_ratio: float = 0
    align_posemb: Literal['left', 'right'] = 'right'
    batch_size: int = 4                      # per GPU
    gradient_accumulation_steps: int = 1
    num_epochs: int = 100
    gradient_clip: float = 1.0
    mixed_precision: Literal['no', 'fp8', 'fp16', 'fp32'] = 'bf16'
    lr: float = 1e-4
    checkpointing: bool = True
    seed: int = 0
    eval_mode: Literal['none', 'loss', 'generate'] = 'loss'
    debug_eval: bool = False
    warmup_ratio: float = 0.01
    use_wandb: bool = False

    # ---- testing / inference ----------------------------------------------------------------
    test_path: Optional[str] = None
    test_resume_tokens: Optional[str] = None
    test_repeat: int = 1
    test_num_face: Tuple[int, ...] = (1000,)
    test_max_seq_length: Optional[int] = None


# Shared by the two point-conditioned presets (reference options.py:158-211).
_POINT_COND = dict(
    point_encoder_mode='embed', kl_weight=1e-8, discrete_bins=512, use_num_face_cond=True,
    cond_mode='point', num_cond_tokens=2049, freeze_encoder=False, use_meto=True,
    meto_backend='LR_ABSCO', max_seq_length=40960, hidden_dim=1536, num_heads=16, num_layers=24,
    gradient_accumulation_steps=1, lr=1e-5,
)

config_defaults: Dict[str, Options] = {}
config_doc: Dict[str, str] = {}

config_doc['default'] = 'the default settings'
config_defaults['default'] = Options()

config_doc['ArAE'] = 'ArAE'
config_defaults['ArAE'] = replace(
    Options(), **_POINT_COND, use_decimate_aug=True, max_face_length=4000, align_posemb='right',
    batch_size=4, warmup_ratio=0, num_epochs=100, eval_mode='loss',
)

config_doc['DiT'] = 'DiT'
config_defaults['DiT'] = replace(
    Options(), **_POINT_COND, use_decimate_aug=False, max_face_length=8000, dit_hidden_dim=1024,
    dit_num_heads=16, dit_num_layers=24, snr_gamma=5.0, noise_scheduler_predtype="v_prediction",
    batch_size=8, num_epochs=300, eval_mode='none',
)

if tyro is not None:
    AllConfigs = tyro.extras.subcommand_type_from_defaults(config_defaults, config_doc)
else:  # pragma: no cover
    AllConfigs = Options
