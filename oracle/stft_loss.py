"""Oracle: multi-resolution STFT loss (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates /root/reference/stable_audio_tools/training/losses/auraloss.py with explicit framing + rFFT instead of
torch.stft, so it is an independent statement of the same arithmetic.
"""
import numpy as np
import scipy.signal
import torch
import torch.nn.functional as F


def a_weighting_fir(fs=44100, ntaps=101):
    # auraloss.py:118-149 — analog A-weighting -> bilinear -> freqz(512) -> firls(101)
    f1, f2, f3, f4, A1000 = 20.598997, 107.65265, 737.86223, 12194.217, 1.9997
    nums = [(2 * np.pi * f4) ** 2 * (10 ** (A1000 / 20)), 0, 0, 0, 0]
    dens = np.polymul([1, 4 * np.pi * f4, (2 * np.pi * f4) ** 2], [1, 4 * np.pi * f1, (2 * np.pi * f1) ** 2])
    dens = np.polymul(np.polymul(dens, [1, 2 * np.pi * f3]), [1, 2 * np.pi * f2])
    b, a = scipy.signal.bilinear(nums, dens, fs=fs)
    w_iir, h_iir = scipy.signal.freqz(b, a, worN=512, fs=fs)
    taps = scipy.signal.firls(ntaps, w_iir, abs(h_iir), fs=fs)
    return torch.tensor(taps.astype("float32"))


def fir_filter(x, taps):
    # auraloss.py:155-169 — F.conv1d(x[B*ch,1,T], taps, padding=ntaps//2) (cross-correlation, as torch)
    b, c, t = x.shape
    y = F.conv1d(x.reshape(b * c, 1, t), taps.view(1, 1, -1), padding=taps.numel() // 2)
    return y.view(b, c, -1)


def stft_mag(x, n_fft, hop, eps=1e-8):
    # auraloss.py:377-387 — torch.stft(center=True, reflect pad, periodic hann, onesided) -> sqrt(clamp(re^2+im^2, eps))
    # x [R, T] -> [R, n_fft//2+1, frames]
    win = torch.hann_window(n_fft, periodic=True, dtype=x.dtype)
    xp = F.pad(x[:, None, :], (n_fft // 2, n_fft // 2), mode="reflect")[:, 0, :]
    frames = xp.unfold(-1, n_fft, hop)                      # [R, frames, n_fft]
    spec = torch.fft.rfft(frames * win, dim=-1)             # [R, frames, bins]
    mag = torch.sqrt(torch.clamp(spec.real ** 2 + spec.imag ** 2, min=eps))
    return mag.transpose(1, 2)


def stft_loss(x, y, n_fft, hop, taps=None, w_sc=1.0, w_log=1.0):
    """STFTLoss.forward(input=x, target=y) — auraloss.py:397-449 with reduction='mean', output='loss'."""
    if taps is not None:
        x, y = fir_filter(x, taps), fir_filter(y, taps)
    xm = stft_mag(x.reshape(-1, x.shape[-1]), n_fft, hop)
    ym = stft_mag(y.reshape(-1, y.shape[-1]), n_fft, hop)
    # SpectralConvergenceLoss :181 — per-row Frobenius ratio, shape [R,1,1]
    sc = (torch.linalg.norm((ym - xm).flatten(1), dim=1) / torch.linalg.norm(ym.flatten(1), dim=1)).view(-1, 1, 1)
    # STFTMagnitudeLoss log, L1 mean :219-223
    lm = (torch.log(xm) - torch.log(ym)).abs().mean()
    loss = w_sc * sc + w_log * lm
    return loss.mean()


def mrstft_loss(x, y, fft_sizes, hop_sizes, taps=None):
    # MultiResolutionSTFTLoss.forward :517-539 — mean over resolutions
    tot = 0.0
    for n, h in zip(fft_sizes, hop_sizes):
        tot = tot + stft_loss(x, y, n, h, taps)
    return tot / len(fft_sizes)


def sum_and_difference_loss(x, y, fft_sizes, hop_sizes, taps=None, w_sum=1.0, w_diff=1.0):
    # SumAndDifferenceSTFTLoss.forward :585-615 — stereo -> (L+R, L-R), each through MRSTFT, averaged
    xs, xd = (x[:, 0] + x[:, 1]).unsqueeze(1), (x[:, 0] - x[:, 1]).unsqueeze(1)
    ys, yd = (y[:, 0] + y[:, 1]).unsqueeze(1), (y[:, 0] - y[:, 1]).unsqueeze(1)
    return (w_sum * mrstft_loss(xs, ys, fft_sizes, hop_sizes, taps) + w_diff * mrstft_loss(xd, yd, fft_sizes, hop_sizes, taps)) / 2
