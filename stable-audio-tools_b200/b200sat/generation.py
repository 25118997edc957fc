"""Public generation API — mirrors the reference's `generate_diffusion_cond` (inference/generation.py:91-220):
seed -> noise -> conditioning -> sampler -> (optional) pretransform decode.

The text/number conditioners (T5, NumberConditioner) are out of scope of the hot path; the caller
passes `conditioning_tensors` exactly as `model.conditioner(...)` would have produced them
({"cross_attn_cond": [B, L, cond_dim], "global_cond": [B, global_dim]}), host or device resident.
"""
import torch

from . import sampling
from .dit_engine import DiTEngine


class DiffusionCondModel:
    """Holds what `ConditionedDiffusionModelWrapper` holds for the hot path (models/diffusion.py:100-135): the DiT, the optional
    autoencoder pretransform, io_channels and the sample rate bookkeeping."""

    def __init__(self, engine, pretransform=None, io_channels=64, sample_rate=44100, downsampling_ratio=2048):
        self.engine = engine
        self.pretransform = pretransform
        self.io_channels = io_channels
        self.sample_rate = sample_rate
        self.downsampling_ratio = downsampling_ratio
        self._samplers = {}

    @classmethod
    def from_state_dict(cls, dit_state_dict, pretransform=None, device="cuda", **kw):
        return cls(DiTEngine(dit_state_dict, device=device), pretransform, **kw)

    def sampler(self, B, T, L, has_global, cfg_scale, scale_phi, use_graph=True):
        key = (B, T, L, has_global, float(cfg_scale), float(scale_phi), use_graph)
        if key not in self._samplers:
            self._samplers[key] = sampling.GraphSampler(self.engine, B, self.io_channels, T, L, has_global, cfg_scale, scale_phi, use_graph)
        return self._samplers[key]


@torch.no_grad()
def generate_diffusion_cond(model, steps=250, cfg_scale=6.0, conditioning_tensors=None, batch_size=1, sample_size=2097152,
                            seed=-1, device="cuda", sampler_type="dpmpp-3m-sde", sigma_min=0.03, sigma_max=1000.0, rho=1.0,
                            scale_phi=0.0, return_latents=False, noise=None, step_noise=None, use_graph=True):
    """Returns decoded audio [B, channels, sample_size] (or latents [B, io_channels, sample_size // ratio] when the model has
    no pretransform or return_latents=True).  Sampler defaults are the reference UI defaults
    (interface/interfaces/diffusion_cond.py:46-49)."""
    T = sample_size // model.downsampling_ratio if model.pretransform is not None or model.downsampling_ratio else sample_size
    if seed == -1:
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    dev = torch.device(device)
    if noise is None:
        g = torch.Generator(device=dev).manual_seed(seed)  # generation.py:138-142
        noise = torch.randn(batch_size, model.io_channels, T, device=dev, generator=g)
    ct = conditioning_tensors or {}
    cross = ct.get("cross_attn_cond")
    glob = ct.get("global_cond")
    L = 0 if cross is None else cross.shape[1]
    smp = model.sampler(batch_size, T, L, glob is not None, cfg_scale, scale_phi, use_graph)
    if sampler_type == "dpmpp-3m-sde":
        lat = sampling.sample_k_dpmpp_3m_sde(model.engine, noise, steps, sigma_min, sigma_max, rho, cross, glob, cfg_scale,
                                             scale_phi, step_noise=step_noise, sampler=smp)
    elif sampler_type == "v-ddim":
        lat = sampling.sample_v_ddim(model.engine, noise, steps, sigma_max, cross, glob, cfg_scale, scale_phi, sampler=smp)
    else:
        raise NotImplementedError(f"sampler_type {sampler_type!r}: only 'dpmpp-3m-sde' and 'v-ddim' are implemented")
    if return_latents or model.pretransform is None:
        return lat
    return model.pretransform.decode(lat)
