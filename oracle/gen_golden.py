"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the authoring container only):

    python -m oracle.gen_golden

The reference has no tests or golden vectors of its own; these fixtures are outputs of the reference modules
on seeded inputs and pin the oracle (and through it the CUDA path).  Recorded with every file: torch version, seeds.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness, dit as odit, oobleck as oo, discriminator as odisc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FFT = [2048, 1024, 512, 256, 128, 64, 32]
HOP = [512, 256, 128, 64, 32, 16, 8]


def _np(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def main():
    torch.set_num_threads(8)
    R = ref_harness.load()
    os.makedirs(OUT, exist_ok=True)
    meta = {"torch": torch.__version__, "reference": "stable-audio-tools @ 50049e37 (v0.0.19)"}

    # ---- DiT (prepend / adaLN), plain + CFG with std-rescale; plus v-objective loss and two gradients
    for gct in ("prepend", "adaLN"):
        kw = dict(embed_dim=128, depth=2, num_heads=2, io_channels=64, cond_token_dim=64, global_cond_dim=128)
        m = R.dit.DiffusionTransformer(project_cond_tokens=False, transformer_type="continuous_transformer", global_cond_type=gct, **kw)
        sd = odit.make_state_dict(global_cond_type=gct, seed=11, **kw)
        m.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(12)
        x = torch.randn(2, 64, 96, generator=g); t = torch.rand(2, generator=g)
        c = torch.randn(2, 7, 64, generator=g); ge = torch.randn(2, 128, generator=g)
        nz = torch.randn(2, 64, 96, generator=g)
        m.eval()
        with torch.no_grad():
            y_plain = m(x, t, cross_attn_cond=c, global_embed=ge)
            y_cfg = m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=6.0, scale_phi=0.75)
        # v-objective step (training/diffusion.py:405-449) with the reference model in train mode, no cfg dropout
        m.train()
        alpha, sigma = torch.cos(t * torch.pi / 2)[:, None, None], torch.sin(t * torch.pi / 2)[:, None, None]
        noised = x * alpha + nz * sigma
        target = nz * alpha - x * sigma
        out = m(noised, t, cross_attn_cond=c, global_embed=ge, cfg_dropout_prob=0.0)
        loss = torch.nn.functional.mse_loss(out, target)
        loss.backward()
        grads = {"grad.transformer.layers.0.self_attn.to_qkv.weight": m.transformer.layers[0].self_attn.to_qkv.weight.grad,
                 "grad.transformer.layers.1.ff.ff.2.weight": m.transformer.layers[1].ff.ff[2].weight.grad,
                 "grad.transformer.layers.0.pre_norm.gamma": m.transformer.layers[0].pre_norm.gamma.grad,
                 "grad.transformer.project_in.weight": m.transformer.project_in.weight.grad}
        np.savez_compressed(os.path.join(OUT, f"dit_{gct}.npz"), meta=json.dumps({**meta, "weights_seed": 11, "cfg": kw, "gct": gct}),
                            **_np(dict(x=x, t=t, cross=c, glob=ge, noise=nz, y_plain=y_plain, y_cfg=y_cfg, loss=loss.detach(), **grads)))

    # ---- Oobleck encoder / VAE / decoder (channels=64 so every conv is tensor-core eligible; structure as stable_audio_2_0_vae.json)
    from stable_audio_tools.models.factory import create_model_from_config
    cfg = json.load(open("/root/reference/stable_audio_tools/configs/model_configs/autoencoders/stable_audio_2_0_vae.json"))
    cfg["model"]["encoder"]["config"].update(channels=64, c_mults=[1, 2, 4], strides=[2, 4, 4], latent_dim=128)
    cfg["model"]["decoder"]["config"].update(channels=64, c_mults=[1, 2, 4], strides=[2, 4, 4], latent_dim=64)
    cfg["model"]["latent_dim"] = 64
    cfg["model"]["downsampling_ratio"] = 32
    ae = create_model_from_config(cfg).eval()
    sd = oo.make_state_dict(channels=64, c_mults=(1, 2, 4), strides=(2, 4, 4), enc_latent=128, dec_latent=64, seed=21)
    ae.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(22)
    x = torch.randn(2, 2, 2048, generator=g) * 0.5
    with torch.no_grad():
        enc = ae.encoder(x)
        torch.manual_seed(1)
        lat, info = ae.bottleneck.encode(enc, return_info=True)
        torch.manual_seed(1)
        vae_noise = torch.randn_like(enc[:, :64])
        dec = ae.decoder(lat)
    np.savez_compressed(os.path.join(OUT, "oobleck_small.npz"), meta=json.dumps({**meta, "weights_seed": 21}),
                        **_np(dict(x=x, enc=enc, vae_noise=vae_noise, latents=lat, kl=info["kl"], dec=dec)))

    # ---- MRSTFT / sum-and-difference losses (stable_audio_2_0_vae.json:96-99 resolutions, A-weighting on)
    A = R.auraloss
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 2, 8192, generator=g) * 0.3
    y = x + 0.1 * torch.randn(2, 2, 8192, generator=g)
    sdl = A.SumAndDifferenceSTFTLoss(fft_sizes=FFT, hop_sizes=HOP, win_lengths=FFT, perceptual_weighting=True, sample_rate=44100)
    mrl = A.MultiResolutionSTFTLoss(fft_sizes=FFT, hop_sizes=HOP, win_lengths=FFT, perceptual_weighting=True, sample_rate=44100)
    xg = x.clone().requires_grad_(True)
    l_sd = sdl(xg, y)
    l_sd.backward()
    l_l = mrl(x[:, :1], y[:, :1])
    l_plain = A.MultiResolutionSTFTLoss(fft_sizes=FFT, hop_sizes=HOP, win_lengths=FFT)(x, y)
    mag = mrl.stft_losses[3].stft(x.view(-1, x.shape[-1]))[0]  # n_fft = 256 magnitudes
    np.savez_compressed(os.path.join(OUT, "mrstft.npz"), meta=json.dumps(meta),
                        **_np(dict(x=x, y=y, loss_sd=l_sd.detach(), grad_sd=xg.grad, loss_l=l_l, loss_plain=l_plain, mag256=mag,
                                   taps=mrl.stft_losses[0].prefilter.fir.weight.data.view(-1))))

    # ---- Encodec multi-scale STFT discriminator (row G1): losses, the generator-side gradient, one scale's logits
    import stable_audio_tools.models.discriminators as RD
    disc = RD.EncodecDiscriminator(filters=64, in_channels=2, n_ffts=list(odisc.N_FFTS), hop_lengths=list(odisc.HOPS), win_lengths=list(odisc.N_FFTS))
    dsd = odisc.make_state_dict(seed=51)
    r = disc.load_state_dict(dsd, strict=False)
    assert not r.unexpected_keys and all("window" in k for k in r.missing_keys)
    g = torch.Generator().manual_seed(52)
    reals = torch.randn(1, 2, 8192, generator=g) * 0.3
    fakes = (reals + 0.1 * torch.randn(1, 2, 8192, generator=g)).requires_grad_(True)
    dis, adv, fm = disc.loss(reals, fakes)
    (0.1 * adv + 5.0 * fm).backward(retain_graph=True)   # generator weights of stable_audio_2_0_vae.json:88-91
    gfakes = fakes.grad.clone()
    disc.zero_grad()
    dis.backward()                                        # the discriminator step's gradients (three representative parameters)
    pd = dict(disc.discriminators.named_parameters())
    dsel = {k: pd[k].grad.clone() for k in ("discriminators.4.convs.0.conv.weight_v", "discriminators.0.convs.2.conv.weight_g",
                                            "discriminators.2.convs.4.conv.bias", "discriminators.1.conv_post.conv.weight_v")}
    fakes.grad = gfakes
    with torch.no_grad():
        logits, _ = disc(reals)
    np.savez_compressed(os.path.join(OUT, "encodec_disc.npz"), meta=json.dumps({**meta, "weights_seed": 51}),
                        **_np(dict(reals=reals, fakes=fakes.detach(), dis=dis.detach(), adv=adv.detach(), fm=fm.detach(), grad_fakes=fakes.grad,
                                   logits4=logits[4], **{"dgrad." + k: v for k, v in dsel.items()})))

    # ---- in-repo v-DDIM sampler with a closed-form toy model (inference/sampling.py:253-307)
    toy = lambda x_, t_, **kw: torch.tanh(x_ * 0.7) * (0.3 + t_.view(-1, 1, 1)) - 0.1 * x_
    g = torch.Generator().manual_seed(41)
    n0 = torch.randn(2, 4, 16, generator=g)
    out = R.sampling.sample(toy, n0, 25, 0.0)
    np.savez_compressed(os.path.join(OUT, "vddim_toy.npz"), meta=json.dumps(meta), **_np(dict(noise=n0, out=out)))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
