"""Golden vectors for the inpainting row from the UNMODIFIED reference (authoring container; needs baseline/_ref):

    python -m oracle.gen_golden_inpaint

1. `DiffusionTransformer` built with input_concat_dim = 65 (models/dit.py:86-88, :160-165): plain and classifier-free-guidance forwards with a
   concatenated conditioning tensor, once at the input length and once shorter (nearest-neighbour resize, dit.py:162-163).
2. The reference's own `generate_diffusion_cond_inpaint` (inference/generation.py:222-405) on a `create_model_from_config` model with
   input_concat_ids = [inpaint_mask, inpaint_masked_input], sampler_type 'v-ddim' (the in-repo deterministic sampler: no k-diffusion
   stand-in on the path), with and without init_audio (variation start, generation.py:365-367).
Writes tests/golden/dit_inpaint.npz."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_loader, ref_models  # noqa: E402
from oracle import dit as odit  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
KW = dict(embed_dim=128, depth=2, num_heads=2, io_channels=64, cond_token_dim=64, global_cond_dim=128)
DC = 65


def main():
    torch.set_num_threads(8)
    R = ref_loader.load(force_sdpa=True)
    meta = {"torch": torch.__version__, "reference": "stable-audio-tools 0.0.19", "weights_seed": 31, "cfg": KW, "input_concat_dim": DC}
    out = {}
    # ---- 1. module forward
    m = R.dit.DiffusionTransformer(project_cond_tokens=False, transformer_type="continuous_transformer", global_cond_type="prepend",
                                   input_concat_dim=DC, **KW)
    sd = odit.make_state_dict(seed=31, input_concat_dim=DC, **KW)
    m.load_state_dict(sd, strict=True)
    m.eval()
    g = torch.Generator().manual_seed(32)
    x = torch.randn(2, 64, 96, generator=g); t = torch.rand(2, generator=g)
    c = torch.randn(2, 7, 64, generator=g); ge = torch.randn(2, 128, generator=g)
    cc = torch.randn(2, DC, 96, generator=g); cc_short = torch.randn(2, DC, 24, generator=g)
    with torch.no_grad():
        out.update(x=x, t=t, cross=c, glob=ge, concat=cc, concat_short=cc_short,
                   y_plain=m(x, t, cross_attn_cond=c, global_embed=ge, input_concat_cond=cc),
                   y_cfg=m(x, t, cross_attn_cond=c, global_embed=ge, input_concat_cond=cc, cfg_scale=5.0),
                   y_cfg_short=m(x, t, cross_attn_cond=c, global_embed=ge, input_concat_cond=cc_short, cfg_scale=5.0, scale_phi=0.5))
    # ---- 2. the generation driver
    cfg = ref_models.sao_config(depth=2, embed_dim=128, num_heads=2, cond_token_dim=64, global_cond_dim=128, sample_size=80)
    cfg["model"]["diffusion"]["config"]["input_concat_dim"] = DC
    cfg["model"]["diffusion"]["input_concat_ids"] = ["inpaint_mask", "inpaint_masked_input"]
    model = ref_models.build_diffusion_cond(R, cfg, seed=0)
    model.model.model.load_state_dict(sd, strict=True)
    B, T = 2, 80
    ct = ref_models.conditioning_tensors(model, B, seed=33, prompt_tokens=5)
    g = torch.Generator().manual_seed(34)
    audio = torch.randn(64, T, generator=g)                    # no pretransform: "audio" is the latent itself (io_channels = 64)
    mask = (torch.arange(T) < 37).float().unsqueeze(0).repeat(B, 1)     # [batch, sample_size]: keep the first 37 steps, regenerate the rest
    init = torch.randn(64, T, generator=g)
    gi = model.get_conditioning_inputs({**ct, "inpaint_mask": [torch.zeros(B, 1, T)], "inpaint_masked_input": [torch.zeros(B, 64, T)]})
    kw = dict(steps=6, cfg_scale=4.0, batch_size=B, sample_size=T, seed=5, device="cpu", return_latents=True, sampler_type="v-ddim")
    with torch.no_grad():
        lat = R.generation.generate_diffusion_cond_inpaint(model, conditioning_tensors=dict(ct), inpaint_audio=(44100, audio), inpaint_mask=mask, **kw)
        lat_init = R.generation.generate_diffusion_cond_inpaint(model, conditioning_tensors=dict(ct), inpaint_audio=(44100, audio), inpaint_mask=mask,
                                                                init_audio=(44100, init), init_noise_level=0.7, **kw)
        lat_nomask = R.generation.generate_diffusion_cond_inpaint(model, conditioning_tensors=dict(ct), **kw)
    torch.manual_seed(5)
    noise = torch.randn(B, 64, T)                              # what the driver drew (generation.py:269-275)
    out.update(gen_cross=gi["cross_attn_cond"], gen_glob=gi["global_cond"], gen_audio=audio, gen_mask=mask, gen_init=init, gen_noise=noise,
               gen_lat=lat, gen_lat_init=lat_init, gen_lat_nomask=lat_nomask)
    meta.update(gen=dict(steps=6, cfg_scale=4.0, seed=5, init_noise_level=0.7, sampler_type="v-ddim"))
    np.savez_compressed(os.path.join(OUT, "dit_inpaint.npz"), meta=json.dumps(meta),
                        **{k: (v.detach().float().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print("wrote dit_inpaint.npz", {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})


if __name__ == "__main__":
    main()
