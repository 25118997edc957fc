"""Golden vectors for the ASSEMBLED autoencoder training step: four consecutive calls of the
UNMODIFIED reference `AutoencoderTrainingWrapper.training_step` (training/autoencoders.py:367-527; loss weights :162-243, warm-up
switch, generator / discriminator alternation :476-483, AuralossLoss argument swap training/losses/losses.py:107-113) on a small
Oobleck VAE + EncodecDiscriminator + 5-resolution MRSTFT, CPU fp32.

    python -m oracle.gen_golden_training        (authoring container; needs baseline/_ref)

pytorch_lightning / ema_pytorch are not installable here: the wrapper runs on the stand-ins of baseline/ref_loader.py (an nn.Module with
.log_dict / .optimizers() / .manual_backward), which only supply plumbing — every number below comes out of reference code.
Recorded per step: the returned loss, every logged loss term, selected gradients and selected parameters after the AdamW update,
plus the VAE noise of the step (seeded torch.manual_seed(100 + step) right before the call)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402
from oracle import oobleck as oo, discriminator as odisc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FFT = [512, 256, 128, 64, 32]
HOP = [128, 64, 32, 16, 8]
DISC_FFT, DISC_HOP = [256, 128], [64, 32]
B, T = 2, 4096
WATCH = ["encoder.layers.1.layers.0.layers.1.weight_v", "encoder.layers.3.layers.4.weight_g", "decoder.layers.1.layers.1.weight_v",
         "decoder.layers.3.layers.2.layers.0.alpha", "decoder.layers.5.weight_v", "encoder.layers.0.bias"]
WATCH_D = ["discriminators.discriminators.0.convs.0.conv.weight_v", "discriminators.discriminators.1.convs.2.conv.weight_g",
           "discriminators.discriminators.1.conv_post.conv.bias"]


def build(R, T_):
    from stable_audio_tools.models.factory import create_model_from_config
    cfg = {"model_type": "autoencoder", "sample_size": T, "sample_rate": 44100, "audio_channels": 2,
           "model": {"encoder": {"type": "oobleck", "config": {"in_channels": 2, "channels": 64, "c_mults": [1, 2, 4], "strides": [2, 4, 4],
                                                                "latent_dim": 128, "use_snake": True}},
                     "decoder": {"type": "oobleck", "config": {"out_channels": 2, "channels": 64, "c_mults": [1, 2, 4], "strides": [2, 4, 4],
                                                                "latent_dim": 64, "use_snake": True, "final_tanh": False}},
                     "bottleneck": {"type": "vae"}, "latent_dim": 64, "downsampling_ratio": 32, "io_channels": 2}}
    ae = create_model_from_config(cfg)
    sd = oo.make_state_dict(channels=64, c_mults=(1, 2, 4), strides=(2, 4, 4), enc_latent=128, dec_latent=64, seed=61)
    ae.load_state_dict(sd, strict=True)
    loss_config = {
        "discriminator": {"type": "encodec", "config": {"filters": 64, "n_ffts": DISC_FFT, "hop_lengths": DISC_HOP, "win_lengths": DISC_FFT},
                          "weights": {"adversarial": 0.1, "feature_matching": 5.0}},
        "spectral": {"type": "mrstft", "config": {"fft_sizes": FFT, "hop_sizes": HOP, "win_lengths": FFT, "perceptual_weighting": True},
                     "weights": {"mrstft": 1.0}},
        "time": {"type": "l1", "weights": {"l1": 0.0}},
        "bottleneck": {"type": "kl", "weights": {"kl": 1e-4}},
    }
    opt_cfg = {"autoencoder": {"optimizer": {"type": "AdamW", "config": {"betas": [0.8, 0.99], "lr": 1.5e-4, "weight_decay": 1e-3}}},
               "discriminator": {"optimizer": {"type": "AdamW", "config": {"betas": [0.8, 0.99], "lr": 3e-4, "weight_decay": 1e-3}}}}
    wrap = T_.autoencoders.AutoencoderTrainingWrapper(ae, sample_rate=44100, loss_config=loss_config, optimizer_configs=opt_cfg,
                                                      warmup_steps=0, use_ema=False)
    dsd = odisc.make_state_dict(seed=62, n_scales=len(DISC_FFT)) if "n_scales" in odisc.make_state_dict.__code__.co_varnames else None
    return wrap, ae, sd, dsd, loss_config, opt_cfg


def main():
    torch.set_num_threads(8)
    R = ref_loader.load(force_sdpa=True)
    T_ = ref_loader.load_training()
    torch.manual_seed(60)
    wrap, ae, sd, dsd, loss_config, opt_cfg = build(R, T_)
    disc = wrap.discriminator
    if dsd is not None:
        r = disc.load_state_dict(dsd, strict=False)
        assert not r.unexpected_keys
    disc_sd = {k: v.detach().clone() for k, v in disc.state_dict().items() if "window" not in k}
    g = torch.Generator().manual_seed(63)
    reals = (torch.randn(B, 2, T, generator=g).clamp(-1, 1) * 0.5)
    out = {"reals": reals.numpy()}
    for k, v in disc_sd.items():
        out["disc_init." + k] = v.numpy()
    # diagnostics for the generator loss terms on a FIXED decoded signal (the step-0 reconstruction): value and gradient w.r.t. the
    # decoded audio of every term, computed with the wrapper's own loss modules (AuralossLoss order: module(reals, decoded))
    with torch.no_grad():
        torch.manual_seed(100)
        lat0, _ = ae.encode(reals, return_info=True)
        dec0 = ae.decode(lat0)
    out["diag.decoded"] = dec0.numpy().copy()
    terms = {
        "mrstft": lambda d: wrap.sdstft(reals, d),
        "left": lambda d: wrap.lrstft(reals[:, 0:1], d[:, 0:1]),
        "right": lambda d: wrap.lrstft(reals[:, 1:2], d[:, 1:2]),
        "adv": lambda d: disc.loss(reals=reals, fakes=d)[1],
        "fm": lambda d: disc.loss(reals=reals, fakes=d)[2],
    }
    for name, fn in terms.items():
        leaf = dec0.clone().requires_grad_(True)
        val = fn(leaf)
        (g_,) = torch.autograd.grad(val, leaf)
        out[f"diag.value.{name}"] = np.float64(val.detach())
        out[f"diag.grad.{name}"] = g_.numpy().copy()
    for step in range(4):
        wrap.global_step = step
        torch.manual_seed(100 + step)
        vae_noise = torch.randn(B, 64, T // 32)          # what `torch.randn_like(mean)` will draw (bottleneck.py:105-134)
        torch.manual_seed(100 + step)
        loss = wrap.training_step((reals, None), step)
        out[f"s{step}.vae_noise"] = vae_noise.numpy()
        out[f"s{step}.loss"] = np.float64(loss.detach())
        for k, v in wrap.logged.items():
            if k.startswith("train/"):
                out[f"s{step}.log.{k[6:]}"] = np.float64(v)
        is_d = step % 2 == 1
        named = dict(disc.named_parameters()) if is_d else dict(ae.named_parameters())
        for n in (WATCH_D if is_d else WATCH):
            out[f"s{step}.grad.{n}"] = named[n].grad.detach().numpy().copy()
            out[f"s{step}.param.{n}"] = named[n].detach().numpy().copy()
        wrap.logged.clear()
    meta = {"torch": torch.__version__, "reference": "stable-audio-tools 0.0.19 AutoencoderTrainingWrapper.training_step", "ae_weights_seed": 61,
            "fft": FFT, "hop": HOP, "disc_fft": DISC_FFT, "disc_hop": DISC_HOP, "loss_config": loss_config, "optimizer_configs": opt_cfg,
            "steps": "0 generator, 1 discriminator, 2 generator, 3 discriminator (global_step % 2, warmup_steps = 0)"}
    np.savez_compressed(os.path.join(OUT, "ae_training_step.npz"), meta=json.dumps(meta), **out)
    print("wrote ae_training_step.npz:", {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items() if ".log." in k or k.endswith(".loss")})


if __name__ == "__main__":
    main()
