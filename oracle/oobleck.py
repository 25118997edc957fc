"""Oracle: Oobleck autoencoder + VAE bottleneck (TEST INFRASTRUCTURE — see oracle/__init__.py).

Functional torch-fp32 restatement on a flat state dict with the reference's names
(`encoder.layers.*`, `decoder.layers.*`; weight-normed convs carry `weight_g` / `weight_v` / `bias`).
Citations are into /root/reference/stable_audio_tools/models/.
"""
import math
import torch
import torch.nn.functional as F


def snake_beta(x, alpha_log, beta_log):
    # blocks.py:291-292, :321-329 — x + 1/(exp(beta)+1e-9) * sin(x*exp(alpha))^2, per channel, log-scale params
    a = torch.exp(alpha_log)[None, :, None]
    b = torch.exp(beta_log)[None, :, None]
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a) ** 2


def wn_weight(g, v):
    # autoencoders.py:23-27 — torch.nn.utils.weight_norm(dim=0): w = g * v / ||v||, norm over all dims but 0
    # (for ConvTranspose1d dim 0 is the INPUT channel: weight_g is [Cin,1,1])
    norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return g * v / norm


def _w(sd, p):
    if (p + "weight_g") in sd:
        return wn_weight(sd[p + "weight_g"], sd[p + "weight_v"])
    return sd[p + "weight"]


def residual_unit(x, sd, p, dilation):
    # autoencoders.py:58-83
    h = snake_beta(x, sd[p + "layers.0.alpha"], sd[p + "layers.0.beta"])
    h = F.conv1d(h, _w(sd, p + "layers.1."), sd[p + "layers.1.bias"], dilation=dilation, padding=3 * dilation)
    h = snake_beta(h, sd[p + "layers.2.alpha"], sd[p + "layers.2.beta"])
    h = F.conv1d(h, _w(sd, p + "layers.3."), sd[p + "layers.3.bias"])
    return x + h


def encoder_block(x, sd, p, stride):
    # autoencoders.py:233-250
    for i, d in enumerate((1, 3, 9)):
        x = residual_unit(x, sd, f"{p}layers.{i}.", d)
    x = snake_beta(x, sd[p + "layers.3.alpha"], sd[p + "layers.3.beta"])
    return F.conv1d(x, _w(sd, p + "layers.4."), sd[p + "layers.4.bias"], stride=stride, padding=math.ceil(stride / 2))


def decoder_block(x, sd, p, stride):
    # autoencoders.py:252-283
    x = snake_beta(x, sd[p + "layers.0.alpha"], sd[p + "layers.0.beta"])
    x = F.conv_transpose1d(x, _w(sd, p + "layers.1."), sd[p + "layers.1.bias"], stride=stride, padding=math.ceil(stride / 2))
    for i, d in enumerate((1, 3, 9)):
        x = residual_unit(x, sd, f"{p}layers.{2 + i}.", d)
    return x


def oobleck_encode(x, sd, strides=(2, 4, 4, 8, 8), pre="encoder."):
    # autoencoders.py:285-317
    p = pre + "layers."
    x = F.conv1d(x, _w(sd, p + "0."), sd[p + "0.bias"], padding=3)
    n = len(strides)
    for i, s in enumerate(strides):
        x = encoder_block(x, sd, f"{p}{i + 1}.", s)
    x = snake_beta(x, sd[f"{p}{n + 1}.alpha"], sd[f"{p}{n + 1}.beta"])
    return F.conv1d(x, _w(sd, f"{p}{n + 2}."), sd[f"{p}{n + 2}.bias"], padding=1)


def oobleck_decode(z, sd, strides=(2, 4, 4, 8, 8), pre="decoder.", final_tanh=False):
    # autoencoders.py:320-362
    p = pre + "layers."
    x = F.conv1d(z, _w(sd, p + "0."), sd[p + "0.bias"], padding=3)
    n = len(strides)
    for i in range(n):
        x = decoder_block(x, sd, f"{p}{i + 1}.", strides[n - 1 - i])
    x = snake_beta(x, sd[f"{p}{n + 1}.alpha"], sd[f"{p}{n + 1}.beta"])
    x = F.conv1d(x, _w(sd, f"{p}{n + 2}."), sd.get(f"{p}{n + 2}.bias"), padding=3)
    return torch.tanh(x) if final_tanh else x


def vae_sample(mean_scale, noise):
    # bottleneck.py:105-113 — returns latents and KL (sum over channel, mean over batch AND time)
    mean, scale = mean_scale.chunk(2, dim=1)
    stdev = F.softplus(scale) + 1e-4
    var = stdev * stdev
    logvar = torch.log(var)
    latents = noise * stdev + mean
    kl = (mean * mean + var - logvar - 1).sum(1).mean()
    return latents, kl


def make_state_dict(channels=128, c_mults=(1, 2, 4, 8, 16), strides=(2, 4, 4, 8, 8), enc_latent=128, dec_latent=64,
                    in_channels=2, seed=0):
    """Seeded random Oobleck weights with the reference's names and shapes (autoencoders.py:233-362)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    cm = [1] + list(c_mults)

    def conv(p, cout, cin, k, bias=True, transpose=False):
        shape = (cin, cout, k) if transpose else (cout, cin, k)
        fan = cin * k if not transpose else cin * k / max(1, 1)
        v = torch.randn(*shape, generator=g) / math.sqrt(fan)
        sd[p + "weight_v"] = v
        gshape = (shape[0], 1, 1)
        sd[p + "weight_g"] = v.flatten(1).norm(dim=1).view(gshape) * (0.7 + 0.07 * torch.randn(gshape, generator=g))
        if bias:
            sd[p + "bias"] = 0.05 * torch.randn(cout, generator=g)

    def snake(p, c):
        sd[p + "alpha"] = 0.3 * torch.randn(c, generator=g)
        sd[p + "beta"] = 0.3 * torch.randn(c, generator=g)

    def ru(p, c):
        snake(p + "layers.0.", c); conv(p + "layers.1.", c, c, 7); snake(p + "layers.2.", c); conv(p + "layers.3.", c, c, 1)

    n = len(strides)
    p = "encoder.layers."
    conv(p + "0.", cm[0] * channels, in_channels, 7)
    for i in range(n):
        ci, co = cm[i] * channels, cm[i + 1] * channels
        for j in range(3):
            ru(f"{p}{i + 1}.layers.{j}.", ci)
        snake(f"{p}{i + 1}.layers.3.", ci)
        conv(f"{p}{i + 1}.layers.4.", co, ci, 2 * strides[i])
    snake(f"{p}{n + 1}.", cm[-1] * channels)
    conv(f"{p}{n + 2}.", enc_latent, cm[-1] * channels, 3)
    p = "decoder.layers."
    conv(p + "0.", cm[-1] * channels, dec_latent, 7)
    for idx, i in enumerate(range(n, 0, -1)):
        ci, co = cm[i] * channels, cm[i - 1] * channels
        q = f"{p}{idx + 1}."
        snake(q + "layers.0.", ci)
        conv(q + "layers.1.", co, ci, 2 * strides[i - 1], transpose=True)
        for j in range(3):
            ru(f"{q}layers.{2 + j}.", co)
    snake(f"{p}{n + 1}.", cm[0] * channels)
    conv(f"{p}{n + 2}.", in_channels, cm[0] * channels, 7, bias=False)
    return sd
