"""Builders for the reference-arm models: the UNMODIFIED reference's `create_model_from_config` on the Stable Audio 2.0 /
Stable-Audio-Open-1.0 architecture values (configs/model_configs/txt2audio/stable_audio_2_0.json:5-89) with seeded random
weights.  The text conditioner (CLAP / T5: needs a checkpoint download) is replaced in the CONFIG by an `int` conditioner under
the same id; the benches and tests pass the prompt embedding as a random [B, 128, 768] tensor through `conditioning_tensors`,
exactly what BASELINE.json's configs call "random T5 cond".  Nothing in the product imports this file."""
import copy

import torch

OOBLECK_PRETRANSFORM = {
    "type": "autoencoder", "iterate_batch": True,
    "config": {
        "encoder": {"type": "oobleck", "config": {"in_channels": 2, "channels": 128, "c_mults": [1, 2, 4, 8, 16], "strides": [2, 4, 4, 8, 8],
                                                   "latent_dim": 128, "use_snake": True}},
        "decoder": {"type": "oobleck", "config": {"out_channels": 2, "channels": 128, "c_mults": [1, 2, 4, 8, 16], "strides": [2, 4, 4, 8, 8],
                                                   "latent_dim": 64, "use_snake": True, "final_tanh": False}},
        "bottleneck": {"type": "vae"}, "latent_dim": 64, "downsampling_ratio": 2048, "io_channels": 2},
}


def sao_config(depth=24, embed_dim=1536, num_heads=24, cond_token_dim=768, global_cond_dim=1536, pretransform=False,
               global_cond_type="prepend", sample_size=2097152, ae_channels=128):
    pt = None
    if pretransform:
        pt = copy.deepcopy(OOBLECK_PRETRANSFORM)
        pt["config"]["encoder"]["config"]["channels"] = ae_channels
        pt["config"]["decoder"]["config"]["channels"] = ae_channels
    dit = {"io_channels": 64, "embed_dim": embed_dim, "depth": depth, "num_heads": num_heads, "cond_token_dim": cond_token_dim,
           "global_cond_dim": global_cond_dim, "project_cond_tokens": False, "transformer_type": "continuous_transformer"}
    if global_cond_type != "prepend":
        dit["global_cond_type"] = global_cond_type
    model = {
        "conditioning": {"configs": [{"id": "prompt", "type": "int", "config": {"min_val": 0, "max_val": 1}},
                                     {"id": "seconds_start", "type": "number", "config": {"min_val": 0, "max_val": 512}},
                                     {"id": "seconds_total", "type": "number", "config": {"min_val": 0, "max_val": 512}}],
                         "cond_dim": cond_token_dim},
        "diffusion": {"cross_attention_cond_ids": ["prompt", "seconds_start", "seconds_total"], "global_cond_ids": ["seconds_start", "seconds_total"],
                      "type": "dit", "config": dit},
        "io_channels": 64,
    }
    if pt is not None:
        model["pretransform"] = pt
    return {"model_type": "diffusion_cond", "sample_size": sample_size, "sample_rate": 44100, "audio_channels": 2, "model": model,
            "training": {"use_ema": True, "log_loss_info": False, "pre_encoded": not pretransform, "cfg_dropout_prob": 0.1,
                         "optimizer_configs": {"diffusion": {"optimizer": {"type": "AdamW", "config": {"lr": 5e-5, "betas": [0.9, 0.999],
                                                                                                      "weight_decay": 1e-3}}}}}}


def rerandomize_zero_init(module, std=0.02, seed=1):
    """The reference zero-initialises every branch output (transformer.py:311-314, :366-367; dit.py:121-123): with untouched
    random-init weights each block is the identity and any comparison is vacuous."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            if p.numel() > 1 and float(p.detach().abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g).mul_(std).to(p.dtype))
    return module


def build_diffusion_cond(ref, config, seed=0, device="cpu", dtype=torch.float32):
    """ref = baseline.ref_loader.load().  Returns the reference's ConditionedDiffusionModelWrapper (eval mode)."""
    with torch.device(device):          # parameters are created on the target device (a 1 B-parameter CPU init takes ~30 s)
        torch.manual_seed(seed)
        model = ref.factory.create_model_from_config(config)
    rerandomize_zero_init(model.model, seed=seed + 1)
    if model.pretransform is not None:
        # random-init weight-norm gains make the encoder's output grow ~2x per block: scale the gains so activations stay O(1)
        with torch.no_grad():
            for n, p in model.pretransform.named_parameters():
                if n.endswith("weight_g"):
                    p.mul_(0.5)
    return model.to(device=device, dtype=dtype).eval().requires_grad_(False)


def conditioning_tensors(model, batch, seconds_total=47.0, prompt_tokens=128, device="cpu", seed=0):
    """What `model.conditioner(metadata, device)` returns (dict id -> (tensor, mask)), with the prompt entry replaced by a seeded
    random [B, prompt_tokens, cond_dim] embedding; the number conditioners are the reference's own modules."""
    meta = [{"prompt": 0, "seconds_start": 0.0, "seconds_total": float(seconds_total)} for _ in range(batch)]
    ct = model.conditioner(meta, device)
    g = torch.Generator().manual_seed(seed)
    dim = ct["seconds_total"][0].shape[-1]
    emb = torch.randn(batch, prompt_tokens, dim, generator=g).to(device=device, dtype=ct["seconds_total"][0].dtype)
    ct["prompt"] = (emb, torch.ones(batch, prompt_tokens, device=device, dtype=torch.bool))
    return ct
