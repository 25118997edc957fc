"""CPU tests of host-side logic that needs no GPU: EMA decay schedule, discriminator tap tables / plane geometry."""
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "stable-audio-tools_b200"))


def test_ema_decay_schedule_matches_published_formula():
    from b200sat.optim import ema_decay_at
    assert ema_decay_at(0) == 0.0 and ema_decay_at(1) == 0.0 and ema_decay_at(2) == 0.0
    for step in (3, 10, 1000, 123456):
        epoch = step - 1 - 1
        assert abs(ema_decay_at(step) - min(1.0 - (1.0 + epoch) ** -0.75, 0.9999)) < 1e-12
    assert ema_decay_at(10 ** 9) == 0.9999
    assert abs(ema_decay_at(50, beta=0.99, inv_gamma=2.0, power=1.0, update_after_step=10) - min(1 - (1 + 39 / 2.0) ** -1.0, 0.99)) < 1e-12


def test_discriminator_tap_tables_are_the_2d_taps_of_the_flattened_plane():
    """Row shift of tap (it, jf) of a (3 x 9, dilation (d, 1)) conv on a plane with row pitch Fp = F + 8 must move exactly `d` frames per
    time tap and one bin per frequency tap, be antisymmetric under tap reversal (the data-gradient conv reuses the table) and come in
    runs of 9 consecutive rows (what the shared-window path detects)."""
    for n_fft in (2048, 128):
        F = n_fft // 2 + 1
        Fp = F + 8
        for d in (1, 2, 4):
            offs = [(k // 9 - 1) * d * Fp + (k % 9 - 4) for k in range(27)]
            for k in range(27):
                it, jf = k // 9 - 1, k % 9 - 4
                assert offs[k] == it * d * Fp + jf
                assert offs[26 - k] == -offs[k]
            for g in range(3):
                assert all(offs[g * 9 + j] == offs[g * 9] + j for j in range(9))
            # a valid bin shifted by any tap stays inside its frame's row group: no wrap into the neighbouring frame
            for f in (0, F - 1):
                for jf in range(-4, 5):
                    assert 0 <= f + 4 + jf < Fp
        offs33 = [(k // 3 - 1) * Fp + (k % 3 - 1) for k in range(9)]
        assert all(offs33[8 - k] == -offs33[k] for k in range(9))


def test_hann_window_power_used_by_the_stft_front_end():
    """csrc/discriminator.cu normalises by 1/sqrt(3 n / 8): the sum of squares of the periodic Hann window."""
    import torch
    for n in (128, 512, 2048):
        w = torch.hann_window(n, periodic=True, dtype=torch.float64)
        assert abs(w.pow(2).sum().item() - 0.375 * n) < 1e-9 * n


def test_inpaint_driver_host_logic_against_the_reference_driver(monkeypatch):
    """Row f4, CPU: what `b200sat.generation.generate_diffusion_cond_inpaint` hands to the sampler (mask resized / repeated, masked latents,
    concatenation order, the init_audio start with sigma_max = init_noise_level) reproduces the latents of the reference's own
    `generate_diffusion_cond_inpaint` (tests/golden/dit_inpaint.npz) when the sampler is the fp32 oracle loop instead of the CUDA-graph one."""
    import json
    import math
    import os
    import types
    import numpy as np
    import torch
    from oracle import dit as odit, sampling as osamp
    from b200sat import generation as gen
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "dit_inpaint.npz"))
    f = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    meta = json.loads(str(z["meta"]))
    cfg, dc, g = meta["cfg"], meta["input_concat_dim"], meta["gen"]
    sd = odit.make_state_dict(seed=meta["weights_seed"], input_concat_dim=dc, **cfg)
    seen = []

    def oracle_v_ddim(engine, noise, steps, sigma_max, cross, glob, cfg_scale, scale_phi, sampler=None, input_concat_cond=None, init_data=None):
        seen.append((tuple(input_concat_cond.shape), init_data is not None, float(sigma_max)))
        fn = lambda x, t: odit.dit_forward(x, t, sd, cfg["depth"], cross, glob, cfg_scale=cfg_scale, scale_phi=scale_phi,
                                           input_concat_cond=input_concat_cond)
        sm = min(float(sigma_max), 1.0)
        x = noise
        if init_data is not None:                                   # inference/sampling.py:394-399
            x = init_data * math.cos(sm * math.pi / 2) + noise * math.sin(sm * math.pi / 2)
        with torch.no_grad():
            return osamp.sample_v_ddim(fn, x, steps, sigma_max=sm)

    monkeypatch.setattr(gen.sampling, "sample_v_ddim", oracle_v_ddim)
    engine = types.SimpleNamespace(cfg=types.SimpleNamespace(input_concat_dim=dc))
    model = types.SimpleNamespace(engine=engine, pretransform=None, io_channels=64, downsampling_ratio=0, sampler=lambda *a, **k: None)
    B, T = f["gen_noise"].shape[0], f["gen_noise"].shape[2]
    common = dict(steps=g["steps"], cfg_scale=g["cfg_scale"], conditioning_tensors={"cross_attn_cond": f["gen_cross"], "global_cond": f["gen_glob"]},
                  batch_size=B, sample_size=T, sampler_type="v-ddim", noise=f["gen_noise"], return_latents=True, device="cpu")
    lat = gen.generate_diffusion_cond_inpaint(model, inpaint_audio=f["gen_audio"], inpaint_mask=f["gen_mask"], **common)
    lat_one_mask = gen.generate_diffusion_cond_inpaint(model, inpaint_audio=f["gen_audio"], inpaint_mask=f["gen_mask"][:1], **common)   # [1, T] mask
    lat_init = gen.generate_diffusion_cond_inpaint(model, inpaint_audio=f["gen_audio"], inpaint_mask=f["gen_mask"], init_audio=f["gen_init"],
                                                   init_noise_level=g["init_noise_level"], **common)
    lat0 = gen.generate_diffusion_cond_inpaint(model, **common)
    for got, key in ((lat, "gen_lat"), (lat_one_mask, "gen_lat"), (lat_init, "gen_lat_init"), (lat0, "gen_lat_nomask")):
        assert (got - f[key]).abs().max() <= 5e-5 * max(1.0, float(f[key].abs().max())), key
    assert seen[0] == ((B, dc, T), False, 1000.0) and seen[2] == ((B, dc, T), True, g["init_noise_level"])
    bad = types.SimpleNamespace(engine=types.SimpleNamespace(cfg=types.SimpleNamespace(input_concat_dim=0)), pretransform=None, io_channels=64,
                                downsampling_ratio=0, sampler=lambda *a, **k: None)
    import pytest
    with pytest.raises(ValueError):
        gen.generate_diffusion_cond_inpaint(bad, **common)
