"""Oracle: Encodec multi-scale STFT discriminator + hinge / feature-matching losses (TEST INFRASTRUCTURE - see oracle/__init__.py).

Functional torch-fp32 restatement of /root/reference/stable_audio_tools/models/encodec.py:38-138 (DiscriminatorSTFT,
MultiScaleSTFTDiscriminator) and models/discriminators.py:13-58 (get_hinge_losses, EncodecDiscriminator.loss) on a flat state dict
with the reference's names (`discriminators.discriminators.{i}.convs.{j}.conv.{weight_g,weight_v,bias}`, `...conv_post.conv.*`).
Pinned against the reference classes by oracle/gen_golden.py -> tests/golden/encodec_disc.npz.  Row G1 of SURVEY.md section 8a: the
kernels for it are round-2 work; this file and its golden vectors are the parity target they will be built against.
"""
import math
import torch
import torch.nn.functional as F

N_FFTS = (2048, 1024, 512, 256, 128)          # configs/model_configs/autoencoders/stable_audio_2_0_vae.json:80-91
HOPS = (512, 256, 128, 64, 32)
DILATIONS = (1, 2, 4)                          # encodec.py:59 default
KERNEL = (3, 9)                                # (time, frequency)


def spectrogram(x, n_fft, hop, win):
    # torchaudio.transforms.Spectrogram(normalized=True, center=False, power=None, window_fn=hann_window) - encodec.py:72-74
    B, C, T = x.shape
    w = torch.hann_window(win, dtype=x.dtype, device=x.device)
    z = torch.stft(x.reshape(B * C, T), n_fft, hop_length=hop, win_length=win, window=w, center=False, normalized=False, onesided=True,
                   return_complex=True)
    z = z / w.pow(2).sum().sqrt()
    return z.reshape(B, C, z.shape[-2], z.shape[-1])          # [B, C, freq, frames]


def _wn(sd, p):
    v, g = sd[p + "weight_v"], sd[p + "weight_g"]
    return g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1, 1)


def disc_stft(x, sd, pre, n_fft, hop, win):
    """DiscriminatorSTFT.forward - encodec.py:94-106.  Returns (logits [B,1,frames,freq], [5 feature maps])."""
    z = spectrogram(x, n_fft, hop, win)
    z = torch.cat([z.real, z.imag], dim=1)                    # [B, 2C, freq, frames]
    z = z.transpose(2, 3)                                     # 'b c w t -> b c t w'
    fmap = []
    pads = [(1, 4)] + [(d, 4) for d in DILATIONS] + [(1, 1)]
    dils = [(1, 1)] + [(d, 1) for d in DILATIONS] + [(1, 1)]
    for j in range(5):
        p = f"{pre}convs.{j}.conv."
        z = F.conv2d(z, _wn(sd, p), sd[p + "bias"], stride=(1, 1), padding=pads[j], dilation=dils[j])
        z = F.leaky_relu(z, 0.2)
        fmap.append(z)
    p = pre + "conv_post.conv."
    return F.conv2d(z, _wn(sd, p), sd[p + "bias"], padding=(1, 1)), fmap


def discriminator_forward(x, sd, pre="discriminators.discriminators.", n_ffts=N_FFTS, hops=HOPS):
    logits, fmaps = [], []
    for i, (n, h) in enumerate(zip(n_ffts, hops)):
        lg, fm = disc_stft(x, sd, f"{pre}{i}.", n, h, n)
        logits.append(lg); fmaps.append(fm)
    return logits, fmaps


def discriminator_loss(reals, fakes, sd, pre="discriminators.discriminators.", n_ffts=N_FFTS, hops=HOPS):
    """EncodecDiscriminator.loss (hinge, normalize_losses=False) - discriminators.py:31-58: (dis_loss, adv_loss, feature_matching)."""
    lt, ft = discriminator_forward(reals, sd, pre, n_ffts, hops)
    lf, ff = discriminator_forward(fakes, sd, pre, n_ffts, hops)
    dis = adv = fm = 0.0
    for i in range(len(lt)):
        fm = fm + sum((a - b).abs().mean() for a, b in zip(ft[i], ff[i])) / len(ft[i])
        dis = dis + torch.relu(1 - lt[i]).mean() + torch.relu(1 + lf[i]).mean()
        adv = adv - lf[i].mean()
    n = len(lt)
    return dis / n, adv / n, fm / n


def make_state_dict(filters=64, in_channels=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(p, cout, cin, kh, kw):
        v = torch.randn(cout, cin, kh, kw, generator=g) / math.sqrt(cin * kh * kw)
        sd[p + "weight_v"] = v
        sd[p + "weight_g"] = v.flatten(1).norm(dim=1).view(-1, 1, 1, 1) * (0.9 + 0.1 * torch.rand(cout, 1, 1, 1, generator=g))
        sd[p + "bias"] = 0.05 * torch.randn(cout, generator=g)

    for i in range(len(N_FFTS)):
        pre = f"discriminators.discriminators.{i}."
        conv(pre + "convs.0.conv.", filters, 2 * in_channels, *KERNEL)
        for j in range(1, 4):
            conv(f"{pre}convs.{j}.conv.", filters, filters, *KERNEL)
        conv(pre + "convs.4.conv.", filters, filters, 3, 3)
        conv(pre + "conv_post.conv.", 1, filters, 3, 3)
    return sd
