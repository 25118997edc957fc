"""Public generation API — mirrors the reference's `generate_diffusion_cond` (inference/generation.py:91-220) and
`generate_diffusion_cond_inpaint` (inference/generation.py:222-405): seed -> noise -> conditioning -> sampler -> (optional)
pretransform decode.

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


@torch.no_grad()
def generate_diffusion_cond_inpaint(model, steps=250, cfg_scale=6.0, conditioning_tensors=None, batch_size=1, sample_size=2097152,
                                    seed=-1, device="cuda", init_audio=None, init_noise_level=1.0, inpaint_audio=None, inpaint_mask=None,
                                    return_latents=False, sampler_type="dpmpp-3m-sde", sigma_min=0.03, sigma_max=1000.0, rho=1.0,
                                    scale_phi=0.0, noise=None, step_noise=None, use_graph=True):
    """`generate_diffusion_cond_inpaint` (inference/generation.py:222-405) for a DiT built with input_concat_dim = 1 + io_channels
    (input_concat_ids = [inpaint_mask, inpaint_masked_input], models/diffusion.py:180-183):

      inpaint_mask  [B or 1, sample_size]  1 = keep the supplied audio, 0 = generate (generation.py:265-266, :339-345)
      inpaint_audio [B or 1, C, sample_size] at the model's sample rate and channel count (the reference's `prepare_audio` resampling /
                    channel fix-up, generation.py:317-328, is host-side data preparation and stays with the caller), or latents when the
                    model has no pretransform
      init_audio    same convention; starts the sampler from the (encoded) audio at noise level `init_noise_level` (generation.py:288-314, :365-367)

    Masks are resized to the latent length with nearest-neighbour interpolation as the reference does; the mask and the masked latents are
    concatenated to the DiT input on every step (dit.py:160-165), identically for both classifier-free-guidance halves (dit.py:336-337)."""
    import torch.nn.functional as F
    dev = torch.device(device)
    latent = model.pretransform is not None
    T = sample_size // model.downsampling_ratio if latent else sample_size
    if seed == -1:
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    if noise is None:
        g = torch.Generator(device=dev).manual_seed(seed)      # generation.py:269-275
        noise = torch.randn(batch_size, model.io_channels, T, device=dev, generator=g)

    def encode(a):
        a = a.to(dev, torch.float32)
        if a.dim() == 2:
            a = a.unsqueeze(0)
        if latent:
            a = model.pretransform.encode(a)
        if a.shape[0] == 1 and batch_size > 1:
            a = a.repeat(batch_size, 1, 1)                     # generation.py:314, :338
        return a

    init_data = encode(init_audio) if init_audio is not None else None
    mask = None
    if inpaint_mask is not None:
        mask = inpaint_mask.to(dev, torch.float32)
        if mask.dim() == 1:
            mask = mask.unsqueeze(0)
        mask = F.interpolate(mask.unsqueeze(1), size=T, mode="nearest")          # [*, 1, T]  (generation.py:311, :336, :343)
        if mask.shape[0] == 1 and batch_size > 1:
            mask = mask.repeat(batch_size, 1, 1)
    else:
        mask = torch.zeros(batch_size, 1, T, device=dev)                         # generation.py:346-347
    if inpaint_audio is not None:
        inpaint_input = encode(inpaint_audio) * mask                             # generation.py:354-355
    else:
        inpaint_input = torch.zeros(batch_size, model.io_channels, T, device=dev)
    concat = torch.cat([mask, inpaint_input], dim=1)                             # input_concat_ids order (models/diffusion.py:183)
    if concat.shape[1] != model.engine.cfg.input_concat_dim:
        raise ValueError(f"inpainting needs a DiT with input_concat_dim = {concat.shape[1]} (mask + masked input), this model has "
                         f"{model.engine.cfg.input_concat_dim}")
    if init_data is not None:
        sigma_max = init_noise_level                                             # generation.py:365-367
    ct = conditioning_tensors or {}
    cross, glob = ct.get("cross_attn_cond"), ct.get("global_cond")
    L = 0 if cross is None else cross.shape[1]
    smp = model.sampler(batch_size, T, L, glob is not None, cfg_scale, scale_phi, use_graph)
    if sampler_type == "dpmpp-3m-sde":
        lat = sampling.sample_k_dpmpp_3m_sde(model.engine, noise, steps, sigma_min, sigma_max, rho, cross, glob, cfg_scale, scale_phi,
                                             step_noise=step_noise, sampler=smp, input_concat_cond=concat, init_data=init_data)
    elif sampler_type == "v-ddim":
        lat = sampling.sample_v_ddim(model.engine, noise, steps, sigma_max, cross, glob, cfg_scale, scale_phi, sampler=smp,
                                     input_concat_cond=concat, init_data=init_data)
    else:
        raise NotImplementedError(f"sampler_type {sampler_type!r}: only 'dpmpp-3m-sde' and 'v-ddim' are implemented")
    if return_latents or not latent:
        return lat
    return model.pretransform.decode(lat)
