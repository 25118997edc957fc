"""CPU: the oracle restatements reproduce the committed reference outputs (tests/golden, made by oracle/gen_golden.py)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import dit as odit, oobleck as oo, stft_loss as ost, sampling as osamp

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(G, name), allow_pickle=False)
    return {k: (torch.from_numpy(z[k]) if k != "meta" else json.loads(str(z[k]))) for k in z.files}


@pytest.mark.parametrize("gct", ["prepend", "adaLN"])
def test_dit_forward_matches_reference(gct):
    f = _load(f"dit_{gct}.npz")
    cfg = f["meta"]["cfg"]
    sd = odit.make_state_dict(global_cond_type=gct, seed=f["meta"]["weights_seed"], **cfg)
    with torch.no_grad():
        y = odit.dit_forward(f["x"], f["t"], sd, cfg["depth"], f["cross"], f["glob"], global_cond_type=gct)
        yc = odit.dit_forward(f["x"], f["t"], sd, cfg["depth"], f["cross"], f["glob"], cfg_scale=6.0, scale_phi=0.75, global_cond_type=gct)
    assert (y - f["y_plain"]).abs().max() <= 1e-5 * max(1.0, f["y_plain"].abs().max())
    assert (yc - f["y_cfg"]).abs().max() <= 2e-5 * max(1.0, f["y_cfg"].abs().max())


@pytest.mark.parametrize("gct", ["prepend", "adaLN"])
def test_dit_v_objective_loss_and_grads(gct):
    f = _load(f"dit_{gct}.npz")
    cfg = f["meta"]["cfg"]
    sd = odit.make_state_dict(global_cond_type=gct, seed=f["meta"]["weights_seed"], **cfg)
    names = [k[5:] for k in f if k.startswith("grad.")]
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    model = lambda x, t: odit.dit_forward(x, t, sd, cfg["depth"], f["cross"], f["glob"], global_cond_type=gct)
    loss, _, _ = odit.v_objective_loss(model, f["x"], f["noise"], f["t"])
    loss.backward()
    assert abs(loss.item() - f["loss"].item()) <= 1e-5 * max(1.0, abs(f["loss"].item()))
    for n in names:
        ref = f["grad." + n]
        assert (sd[n].grad - ref).abs().max() <= 1e-4 * max(1e-6, ref.abs().max()), n


def test_oobleck_matches_reference():
    f = _load("oobleck_small.npz")
    sd = oo.make_state_dict(channels=64, c_mults=(1, 2, 4), strides=(2, 4, 4), enc_latent=128, dec_latent=64, seed=f["meta"]["weights_seed"])
    with torch.no_grad():
        enc = oo.oobleck_encode(f["x"], sd, (2, 4, 4))
        lat, kl = oo.vae_sample(enc, f["vae_noise"])
        dec = oo.oobleck_decode(f["latents"], sd, (2, 4, 4))
    assert (enc - f["enc"]).abs().max() <= 1e-4 * f["enc"].abs().max()
    assert (lat - f["latents"]).abs().max() <= 1e-4 * f["latents"].abs().max()
    assert abs(kl.item() - f["kl"].item()) <= 1e-4 * abs(f["kl"].item())
    assert (dec - f["dec"]).abs().max() <= 1e-4 * f["dec"].abs().max()


def test_mrstft_matches_reference():
    f = _load("mrstft.npz")
    ff = [2048, 1024, 512, 256, 128, 64, 32]; hs = [n // 4 for n in ff]
    taps = ost.a_weighting_fir()
    assert (taps - f["taps"]).abs().max() <= 1e-7
    x = f["x"].clone().requires_grad_(True)
    l_sd = ost.sum_and_difference_loss(x, f["y"], ff, hs, taps)
    l_sd.backward()
    assert abs(l_sd.item() - f["loss_sd"].item()) <= 1e-5
    assert (x.grad - f["grad_sd"]).abs().max() <= 1e-3 * f["grad_sd"].abs().max()
    assert abs(ost.mrstft_loss(f["x"][:, :1], f["y"][:, :1], ff, hs, taps).item() - f["loss_l"].item()) <= 1e-5
    assert abs(ost.mrstft_loss(f["x"], f["y"], ff, hs, None).item() - f["loss_plain"].item()) <= 1e-5
    mag = ost.stft_mag(f["x"].reshape(-1, f["x"].shape[-1]), 256, 64)
    assert (mag - f["mag256"]).abs().max() <= 1e-4 * f["mag256"].abs().max()


def test_vddim_matches_reference():
    f = _load("vddim_toy.npz")
    toy = lambda x_, t_, **kw: torch.tanh(x_ * 0.7) * (0.3 + t_.view(-1, 1, 1)) - 0.1 * x_
    out = osamp.sample_v_ddim(toy, f["noise"], 25)
    assert (out - f["out"]).abs().max() <= 1e-6


def test_encodec_discriminator_oracle_matches_reference_golden():
    """Row G1: the discriminator oracle - STFT front end, five weight-normed 2-D conv scales, hinge + feature-matching
    losses, and the generator-side gradient w.r.t. the decoded audio - against vectors produced by the reference classes."""
    from oracle import discriminator as od
    z = np.load(os.path.join(G, "encodec_disc.npz"))
    meta = json.loads(str(z["meta"]))
    sd = od.make_state_dict(seed=meta["weights_seed"])
    reals = torch.from_numpy(z["reals"])
    fakes = torch.from_numpy(z["fakes"]).requires_grad_(True)
    dis, adv, fm = od.discriminator_loss(reals, fakes, sd)
    (0.1 * adv + 5.0 * fm).backward()
    for name, got in (("dis", dis), ("adv", adv), ("fm", fm)):
        ref = float(z[name])
        assert abs(float(got) - ref) <= 1e-5 * max(1.0, abs(ref)), (name, float(got), ref)
    gref = torch.from_numpy(z["grad_fakes"])
    assert ((fakes.grad - gref).norm() / gref.norm()).item() <= 1e-4
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    od.discriminator_loss(reals, fakes.detach(), leaves)[0].backward()     # the discriminator step
    for k in z.files:
        if k.startswith("dgrad."):
            ref = torch.from_numpy(z[k])
            got = leaves["discriminators." + k[6:]].grad
            assert ((got - ref).norm() / (ref.norm() + 1e-30)).item() <= 2e-3 or (got - ref).norm().item() <= 1e-7, k   # fp32 summation order (the reference recomputes under checkpoint)
    with torch.no_grad():
        logits, fmaps = od.discriminator_forward(reals, sd)
    l4 = torch.from_numpy(z["logits4"])
    assert logits[4].shape == l4.shape and ((logits[4] - l4).norm() / l4.norm()).item() <= 1e-5
    assert [tuple(f.shape[1:3]) for f in fmaps[0]] == [(64, 13)] * 5 and fmaps[0][0].shape[-1] == 1025


def _inpaint_concat(f):
    mask = f["gen_mask"].unsqueeze(1)
    return torch.cat([mask, f["gen_audio"].unsqueeze(0).repeat(mask.shape[0], 1, 1) * mask], dim=1)


def test_dit_input_concat_and_inpaint_driver_match_reference():
    """Row f4: the oracle's input_concat_cond path against the reference module (dit.py:160-165, CFG duplication :336-337, nearest resize
    :162-163), and the oracle v-DDIM loop driven like `generate_diffusion_cond_inpaint` (inference/generation.py:222-405: mask + masked
    input concatenated, optional init_audio start at init_noise_level) against the reference driver's own latents."""
    f = _load("dit_inpaint.npz")
    cfg, dc = f["meta"]["cfg"], f["meta"]["input_concat_dim"]
    sd = odit.make_state_dict(seed=f["meta"]["weights_seed"], input_concat_dim=dc, **cfg)
    with torch.no_grad():
        y = odit.dit_forward(f["x"], f["t"], sd, cfg["depth"], f["cross"], f["glob"], input_concat_cond=f["concat"])
        yc = odit.dit_forward(f["x"], f["t"], sd, cfg["depth"], f["cross"], f["glob"], cfg_scale=5.0, input_concat_cond=f["concat"])
        ys = odit.dit_forward(f["x"], f["t"], sd, cfg["depth"], f["cross"], f["glob"], cfg_scale=5.0, scale_phi=0.5, input_concat_cond=f["concat_short"])
    for got, key in ((y, "y_plain"), (yc, "y_cfg"), (ys, "y_cfg_short")):
        assert (got - f[key]).abs().max() <= 2e-5 * max(1.0, f[key].abs().max()), key
    gen = f["meta"]["gen"]
    concat = _inpaint_concat(f)
    fn = lambda x, t: odit.dit_forward(x, t, sd, cfg["depth"], f["gen_cross"], f["gen_glob"], cfg_scale=gen["cfg_scale"], input_concat_cond=concat)
    with torch.no_grad():
        lat = osamp.sample_v_ddim(fn, f["gen_noise"], gen["steps"])
        sm = gen["init_noise_level"]
        a0, s0 = math.cos(sm * math.pi / 2), math.sin(sm * math.pi / 2)
        lat_init = osamp.sample_v_ddim(fn, f["gen_init"].unsqueeze(0) * a0 + f["gen_noise"] * s0, gen["steps"], sigma_max=sm)
        zero = torch.zeros_like(concat)
        fn0 = lambda x, t: odit.dit_forward(x, t, sd, cfg["depth"], f["gen_cross"], f["gen_glob"], cfg_scale=gen["cfg_scale"], input_concat_cond=zero)
        lat0 = osamp.sample_v_ddim(fn0, f["gen_noise"], gen["steps"])
    for got, key in ((lat, "gen_lat"), (lat_init, "gen_lat_init"), (lat0, "gen_lat_nomask")):
        assert (got - f[key]).abs().max() <= 5e-5 * max(1.0, f[key].abs().max()), key
