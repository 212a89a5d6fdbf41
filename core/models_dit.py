"""``MDiT`` — image-conditioned latent diffusion facade, drop-in for ``/root/reference/core/models_dit.py:33-229`` on the inference
path (``infer_dit.py``: ``model_dit.run(cond)`` -> latents ``[B, 2048, 64]`` -> ``LMM.generate`` in ``point_latent`` mode).

Same constructor, attributes and checkpoint key schema (``dit.*``, ``proj_cond.*``, ``norm_cond.*``, ``image_encoder.*``,
``point_encoder.*``).  Where the work happens:

* ``image_encoder``: the CLIP ViT-H/14 vision tower — a third-party library model in the reference (transformers) and here.  With no network
  the pretrained weights cannot be fetched: the tower is built from its published configuration and randomly initialised unless a
  checkpoint provides ``image_encoder.*`` (or the files are in the local HF cache).
* ``proj_cond`` / ``norm_cond``, the 24-layer DiT, classifier-free guidance and the scheduler update: the sm_100a library
  (``edgerunner_b200/csrc/dit.cu``) — the whole sampling loop runs on the device, one CUDA graph per step.
* ``DDIMScheduler`` / ``DDPMScheduler``: diffusers (0.30.2 in the reference's lock file) is not installed in this image.  The few things
  ``MDiT`` uses from it — the scaled-linear beta schedule, "leading" timestep spacing with ``steps_offset=1``, ``add_noise`` and the
  eta = 0 DDIM update — are restated below from the published algorithm; the per-step scalars go to the kernel as a table.
* ``forward`` (the training loss) needs the backward pass, which is outside the B200 path: it raises.
"""

from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from core.options import Options
from core.transformer.dit import DiT


class _SchedulerConfig(dict):
    __getattr__ = dict.__getitem__


class DDIMScheduler:
    """The subset of ``diffusers.DDIMScheduler`` (0.30.2, schedulers/scheduling_ddim.py) that ``MDiT`` uses: ``set_timesteps``,
    ``timesteps``, ``scale_model_input`` (identity), ``add_noise``, and the eta = 0 update as per-step scalars (``step_coefficients``)."""

    def __init__(self, prediction_type='v_prediction', num_train_timesteps=1000, beta_schedule='scaled_linear', beta_start=0.00085,
                 beta_end=0.012, clip_sample=False, thresholding=False, timestep_spacing='leading', set_alpha_to_one=False, steps_offset=1):
        if beta_schedule != 'scaled_linear' or clip_sample or thresholding or timestep_spacing != 'leading':
            raise NotImplementedError('only the configuration MDiT uses is implemented (scaled_linear, no clipping, leading spacing)')
        if prediction_type not in ('epsilon', 'v_prediction'):
            raise ValueError(f'Unknown prediction type {prediction_type}')
        self.config = _SchedulerConfig(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                       beta_end=beta_end, steps_offset=steps_offset, set_alpha_to_one=set_alpha_to_one)
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError('num_inference_steps exceeds num_train_timesteps')
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = torch.as_tensor(timesteps).to(original_samples.device)
        a = ac[timesteps] ** 0.5
        b = (1 - ac[timesteps]) ** 0.5
        a = a.flatten()
        b = b.flatten()
        while a.dim() < original_samples.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * original_samples + b * noise

    def step_coefficients(self, timesteps):
        """[len(timesteps), 4] fp32: sqrt(alpha_t), sqrt(1 - alpha_t), sqrt(alpha_prev), sqrt(1 - alpha_prev) — the scalars of
        ``DDIMScheduler.step`` with eta = 0 (sigma_t = 0), evaluated in fp32 like the 0-dim tensors diffusers computes them with."""
        out = torch.empty(len(timesteps), 4, dtype=torch.float32)
        for i, t in enumerate(torch.as_tensor(timesteps).tolist()):
            prev = t - self.config.num_train_timesteps // self.num_inference_steps
            a_t = self.alphas_cumprod[t]
            a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
            out[i, 0] = a_t ** 0.5
            out[i, 1] = (1 - a_t) ** 0.5
            out[i, 2] = a_prev ** 0.5
            out[i, 3] = (1 - a_prev - 0.0 ** 2) ** 0.5
        return out


class DDPMScheduler(DDIMScheduler):
    """Training-side scheduler of the reference (models_dit.py:78-87): same schedule; only ``add_noise`` / ``get_velocity`` are used."""

    def __init__(self, **kw):
        kw.pop('set_alpha_to_one', None), kw.pop('steps_offset', None)
        super().__init__(set_alpha_to_one=True, steps_offset=0, **kw)

    def get_velocity(self, sample, noise, timesteps):
        ac = self.alphas_cumprod.to(device=sample.device, dtype=sample.dtype)
        a = (ac[timesteps] ** 0.5).flatten()
        b = ((1 - ac[timesteps]) ** 0.5).flatten()
        while a.dim() < sample.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * noise - b * sample


CLIP_VIT_H_14 = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16, image_size=224, patch_size=14,
                     hidden_act='gelu', projection_dim=1024, layer_norm_eps=1e-5)     # laion/CLIP-ViT-H-14-laion2B-s32B-b79K, vision_config


def build_image_encoder(name='laion/CLIP-ViT-H-14-laion2B-s32B-b79K', vision_config=None):
    """The reference's ``CLIPVisionModel.from_pretrained(name)`` (models_dit.py:54); offline: local cache if present, else the published
    architecture with random weights (a checkpoint that carries ``image_encoder.*`` then fills them)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    if vision_config is None:
        try:
            return CLIPVisionModel.from_pretrained(name, local_files_only=True)
        except Exception:
            print(f'[WARN] {name} is not in the local cache (no network): CLIP vision tower randomly initialised')
            vision_config = CLIP_VIT_H_14
    return CLIPVisionModel(CLIPVisionConfig(**vision_config))


class MDiT(nn.Module):
    def __init__(self, opt: Options, image_encoder_config=None):
        super().__init__()
        self.opt = opt
        self.image_encoder = build_image_encoder(vision_config=image_encoder_config).eval().half()
        self.image_encoder.requires_grad_(False)
        cond_dim = self.image_encoder.config.hidden_size
        icfg = self.image_encoder.config
        self.cond_tokens = (icfg.image_size // icfg.patch_size) ** 2 + 1
        self.dit = DiT(hidden_dim=opt.dit_hidden_dim, num_heads=opt.dit_num_heads, latent_size=opt.point_latent_size,
                       latent_dim=opt.point_latent_dim, num_layers=opt.dit_num_layers, gradient_checkpointing=opt.checkpointing,
                       cond_tokens=self.cond_tokens, cond_dim=cond_dim)
        self.normalize_mean = (0.48145466, 0.4578275, 0.40821073)
        self.normalize_std = (0.26862954, 0.26130258, 0.27577711)
        # condition adaptor (out of dit)
        self.proj_cond = nn.Linear(cond_dim, opt.dit_hidden_dim)
        self.norm_cond = nn.LayerNorm(opt.dit_hidden_dim)
        # on-the-fly point encoder of the training step (reference :62-75): parameters only, so that checkpoints load
        if opt.point_encoder_mode != 'embed':
            raise NotImplementedError("point_encoder_mode='downsample' needs torch_cluster FPS; no preset uses it")
        from core.transformer.point import PointEncoderEmbed
        self.point_encoder = PointEncoderEmbed(hidden_dim=opt.point_hidden_dim, num_heads=opt.point_num_heads, latent_size=opt.point_latent_size,
                                               latent_dim=opt.point_latent_dim, gradient_checkpointing=False).eval().half()
        self.point_encoder.requires_grad_(False)
        sched = dict(prediction_type=opt.noise_scheduler_predtype, num_train_timesteps=1000, beta_schedule='scaled_linear', beta_start=0.00085,
                     beta_end=0.012, clip_sample=False, thresholding=False, timestep_spacing='leading')
        self.noise_scheduler = DDPMScheduler(**sched)
        self.scheduler = DDIMScheduler(**sched, set_alpha_to_one=False, steps_offset=1)

    # ---- engine ---------------------------------------------------------------------------------------------------------------
    def get_engine(self):
        extra = {'proj_cond.weight': self.proj_cond.weight, 'proj_cond.bias': self.proj_cond.bias, 'norm_cond.weight': self.norm_cond.weight,
                 'norm_cond.bias': self.norm_cond.bias}
        return self.dit.get_engine(extra_state=extra)

    def normalize_image(self, x):
        mean = torch.as_tensor(self.normalize_mean, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
        std = torch.as_tensor(self.normalize_std, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
        return (x - mean) / std

    @torch.no_grad()
    def get_cond(self, inputs):
        """inputs [B, 3, H, W] in [0, 1] -> cond [B, 257, dit_hidden_dim] fp32 (reference :106-118)."""
        images_clip = self.normalize_image(inputs)
        size = self.image_encoder.config.image_size                                   # 224 for ViT-H/14 (reference :112)
        images_clip = F.interpolate(images_clip, (size, size), mode='bilinear', align_corners=False)
        images_clip = images_clip.to(device=self.image_encoder.device)
        hidden = self.image_encoder(images_clip.to(self.image_encoder.dtype)).last_hidden_state          # [B, 257, 1280], library model
        return self.get_engine().cond(hidden)

    def forward(self, data, step_ratio=1):
        raise NotImplementedError('MDiT.forward is the training loss (reference core/models_dit.py:121-181); the backward pass is outside the '
                                  'B200 path — use run() for sampling')

    @torch.no_grad()
    def run(self, inputs, num_inference_steps=100, guidance_scale=7.5, num_repeat=1, latents=None, strength=0.5):
        """Denoise sampling (reference :184-229) -> latents [B * num_repeat, latent_size, latent_dim] fp32."""
        device = next(self.dit.parameters()).device
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        cond = self.get_cond(inputs)
        batch_size = cond.shape[0]
        cond = cond.repeat_interleave(num_repeat, dim=0)
        if latents is None:
            init_step = 0
            latents = torch.randn(batch_size * num_repeat, self.opt.point_latent_size, self.opt.point_latent_dim, device=device, dtype=torch.float32)
        else:
            init_step = int(num_inference_steps * strength)
            latents = self.scheduler.add_noise(latents, torch.randn_like(latents), self.scheduler.timesteps[init_step])
        latents = latents.to(device=device, dtype=torch.float32).contiguous().clone()
        ts = self.scheduler.timesteps[init_step:].cpu()
        coef = self.scheduler.step_coefficients(ts)
        # guidance (cat([zeros, cond]), uncond + s * (cond - uncond)) and scheduler.step happen inside the engine's loop
        self.get_engine().run(cond, latents, ts.numpy().astype(np.float32), coef.numpy(), guidance_scale=guidance_scale, guided=True,
                              prediction_type=self.opt.noise_scheduler_predtype)
        return latents
